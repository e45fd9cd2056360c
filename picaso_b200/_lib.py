"""ctypes binding of libpicaso_b200.so (declared in include/picaso_b200.h).

There is deliberately no CPU fallback: if the library is missing or no CUDA device is
present, every entry point raises.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PICASO_B200_LIB") or os.path.join(_HERE, "_build", "libpicaso_b200.so")

PB_HOST, PB_DEVICE = 0, 1
_dp = ctypes.POINTER(ctypes.c_double)
c_int, c_i64, c_dbl, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class PicasoB200Error(RuntimeError):
    pass


class ReflectedArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("nlayer", "nwno", "numg", "numt", "nbatch")] + [("ld", c_i64)] +
        [(n, c_vp) for n in ("dtau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og",
                             "w0_og", "cosb_og", "tau", "tau_og", "surf_reflect", "F0PI", "b_top",
                             "ubar0", "ubar1", "gweight", "tweight")] +
        [("cos_theta", c_dbl), ("single_phase", c_int), ("multi_phase", c_int),
         ("toon_coefficients", c_int)] +
        [(n, c_dbl) for n in ("frac_a", "frac_b", "frac_c", "constant_back", "constant_forward")] +
        [("get_toa_intensity", c_int), ("get_lvl_flux", c_int)] +
        [(n, c_vp) for n in ("xint_at_top", "albedo", "flux_minus", "flux_plus", "flux_minus_mdpt",
                             "flux_plus_mdpt")] + [("variant", c_int), ("gather", c_vp)])


class PeerGather(ctypes.Structure):
    _fields_ = [("nranks", c_int), ("rank", c_int), ("albedo", c_vp), ("flags", c_vp),
                ("step", ctypes.c_uint64), ("wait_step", ctypes.c_uint64), ("done_counter", c_vp),
                ("push", c_int), ("slot", c_int), ("albedo_prev", c_vp)]


class ShArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("nlayer", "nwno", "numg", "numt", "nbatch")] + [("ld", c_i64)] +
        [(n, c_vp) for n in ("dtau", "w0", "ftau_cld", "ftau_ray", "f_deltaM", "dtau_og", "w0_og",
                             "cosb_og", "tau", "tau_og", "surf_reflect", "F0PI", "b_top", "ubar0",
                             "ubar1", "gweight", "tweight")] +
        [("cos_theta", c_dbl)] +
        [(n, c_int) for n in ("w_single_form", "w_multi_form", "psingle_form", "w_single_rayleigh",
                              "w_multi_rayleigh", "psingle_rayleigh")] +
        [(n, c_dbl) for n in ("frac_a", "frac_b", "frac_c", "constant_back", "constant_forward")] +
        [("stream", c_int), ("flx", c_int), ("single_form", c_int)] +
        [(n, c_vp) for n in ("xint_at_top", "albedo", "f_deltaM_out", "flux")])


class ThermalShArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("nlayer", "nwno", "numg", "numt", "nbatch")] + [("ld", c_i64)] +
        [(n, c_vp) for n in ("dtau", "w0", "cosb", "cosb_og", "wno", "surf_reflect", "tlevel", "plevel",
                             "ubar1", "gweight", "tweight")] +
        [("stream", c_int), ("hard_surface", c_int), ("flx", c_int)] +
        [(n, c_vp) for n in ("xint_at_top", "thermal")])


class OpacityArgs(ctypes.Structure):
    _fields_ = (
        [("nlayer", c_int), ("query", c_int), ("pt_index", c_vp), ("weights", c_vp), ("mol_scale", c_vp),
         ("cont_index", c_vp), ("cont_scale", c_vp), ("ray_scale", c_vp), ("raman", c_int), ("jfrac", c_vp),
         ("raman_pollack", c_vp), ("cloud_opd", c_vp), ("cloud_w0", c_vp), ("cloud_g0", c_vp),
         ("cloud_ld", c_i64), ("fthin_cld", c_dbl), ("do_holes", c_int), ("stream", c_int),
         ("delta_eddington", c_int)] +
        [(n, c_vp) for n in ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "DTAU_OG", "TAU_OG",
                             "W0_OG", "COSB_OG", "W0_no_raman", "f_deltaM")] +
        [("ngauss", c_int), ("ck_index", c_vp), ("ck_weights", c_vp), ("ck_scale", c_vp), ("cont_mode", c_int),
         ("cont_index_hi", c_vp), ("cont_t", c_vp), ("ck_direct", c_vp), ("TAUGAS", c_vp), ("TAURAY", c_vp),
         ("TAUCLD", c_vp), ("test_mode", c_int)])


class SpectrumArgs(ctypes.Structure):
    _fields_ = ([("opacity", OpacityArgs), ("nwno", c_int), ("numg", c_int), ("numt", c_int)] +
                [(n, c_vp) for n in ("ubar0", "ubar1", "gweight", "tweight")] + [("cos_theta", c_dbl)] +
                [(n, c_vp) for n in ("surf_reflect", "F0PI", "b_top")] +
                [(n, c_int) for n in ("single_phase", "multi_phase", "toon_coefficients")] +
                [(n, c_dbl) for n in ("frac_a", "frac_b", "frac_c", "constant_back", "constant_forward")] +
                [("albedo", c_vp), ("xint_at_top", c_vp)])


class SpectrumThermalArgs(ctypes.Structure):
    _fields_ = ([("opacity", OpacityArgs), ("nwno", c_int), ("numg", c_int), ("numt", c_int)] +
                [(n, c_vp) for n in ("tlevel", "plevel", "ubar1", "gweight", "tweight", "wno", "surf_reflect")] +
                [("hard_surface", c_int), ("thermal", c_vp), ("flux_at_top", c_vp)])


class SpectrumTransitArgs(ctypes.Structure):
    _fields_ = ([("opacity", OpacityArgs), ("nwno", c_int)] +
                [(n, c_vp) for n in ("z", "dz", "player", "tlayer", "mmw", "colden")] +
                [(n, c_dbl) for n in ("rstar", "k_b", "amu")] + [("F", c_vp)])


class CkMixArgs(ctypes.Structure):
    _fields_ = ([(n, c_int) for n in ("nlayer", "nwno", "ngauss", "ngas", "np", "nt")] +
                [(n, c_vp) for n in ("kappas", "mixes", "indices", "t_interp", "p_interp", "gauss_pts", "gauss_wts",
                                     "molecular_opa", "ln_mixed")])


class ThermalArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("nlayer", "nwno", "numg", "numt", "nbatch")] + [("ld", c_i64)] +
        [(n, c_vp) for n in ("dtau", "w0", "cosb", "wno", "dwno", "surf_reflect", "tlevel",
                             "plevel", "ubar1", "gweight", "tweight")] +
        [("hard_surface", c_int), ("calc_type", c_int)] +
        [(n, c_vp) for n in ("flux_at_top", "thermal", "flux_minus", "flux_plus",
                             "flux_minus_mdpt", "flux_plus_mdpt")] + [("variant", c_int), ("opacity_period", c_int)])


class ClimateArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("nlayer", "nwno", "ngauss", "numg", "numt", "reflected", "thermal")] +
        [(n, c_vp) for n in ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "W0_no_raman",
                             "DTAU_OG", "TAU_OG", "W0_OG", "COSB_OG", "gauss_wts", "wno", "dwno",
                             "surf_reflect", "F0PI", "tlevel", "plevel", "ubar1", "gweight", "tweight")] +
        [("cos_theta", c_dbl), ("single_phase", c_int), ("multi_phase", c_int)] +
        [(n, c_dbl) for n in ("frac_a", "frac_b", "frac_c", "constant_back", "constant_forward")] +
        [(n, c_vp) for n in ("flux_net_v_layer", "flux_net_v", "flux_plus_v", "flux_minus_v",
                             "flux_net_ir_layer", "flux_net_ir", "flux_plus_ir", "flux_minus_ir", "packed")] +
        [("packed_full", c_int), ("nprofiles", c_int), ("tlevels", c_vp), ("jac_out", c_vp)])


class TransitArgs(ctypes.Structure):
    _fields_ = (
        [(n, c_int) for n in ("nlevel", "nwno", "nbatch")] + [("ld", c_i64)] +
        [(n, c_vp) for n in ("DTAU", "z", "dz", "player", "tlayer", "mmw", "colden")] +
        [("rstar", c_dbl), ("k_b", c_dbl), ("amu", c_dbl), ("F", c_vp)])


# every symbol include/picaso_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pb_version": (c_int, []),
    "pb_device_count": (c_int, [ctypes.POINTER(c_int)]),
    "pb_create": (c_int, [c_int, ctypes.POINTER(c_vp)]),
    "pb_destroy": (None, [c_vp]),
    "pb_last_error": (ctypes.c_char_p, [c_vp]),
    "pb_device_name": (c_int, [c_vp, ctypes.c_char_p, ctypes.c_size_t]),
    "pb_sm_count": (c_int, [c_vp, ctypes.POINTER(c_int)]),
    "pb_dev_alloc": (c_int, [c_vp, ctypes.c_size_t, ctypes.POINTER(c_vp)]),
    "pb_dev_free": (c_int, [c_vp, c_vp]),
    "pb_host_alloc": (c_int, [c_vp, ctypes.c_size_t, ctypes.POINTER(c_vp)]),
    "pb_host_free": (c_int, [c_vp, c_vp]),
    "pb_memcpy_h2d": (c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t]),
    "pb_memcpy_d2h": (c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t]),
    "pb_memset": (c_int, [c_vp, c_vp, c_int, ctypes.c_size_t]),
    "pb_sync": (c_int, [c_vp]),
    "pb_set_stream": (c_int, [c_vp, c_vp]),
    "pb_timer_start": (c_int, [c_vp]),
    "pb_timer_stop": (c_int, [c_vp, ctypes.POINTER(ctypes.c_float)]),
    "pb_launch_count": (ctypes.c_uint64, [c_vp]),
    "pb_reflected_toon_1d": (c_int, [c_vp, ctypes.POINTER(ReflectedArgs), c_int]),
    "pb_reflected_sh": (c_int, [c_vp, ctypes.POINTER(ShArgs), c_int]),
    "pb_thermal_sh": (c_int, [c_vp, ctypes.POINTER(ThermalShArgs), c_int]),
    "pb_thermal_toon_1d": (c_int, [c_vp, ctypes.POINTER(ThermalArgs), c_int]),
    "pb_transit_1d": (c_int, [c_vp, ctypes.POINTER(TransitArgs), c_int]),
    "pb_compress_disco": (c_int, [c_vp, c_int, c_dbl, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp,
                                  c_int]),
    "pb_compress_thermal": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int]),
    "pb_selftest_math": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp]),
    "pb_microbench": (c_int, [c_vp, c_int, c_int, c_vp]),
    "pb_climate_bind": (c_int, [c_vp, c_vp, c_int, ctypes.POINTER(c_int)]),
    "pb_climate_run_bound": (c_int, [ctypes.c_ulonglong, c_int]),
    "pb_climate_unbind": (c_int, [c_vp, c_int]),
    "pb_peer_signal": (c_int, [c_vp, c_vp, c_int, c_int, c_int, ctypes.c_ulonglong]),
    "pb_peer_flush": (c_int, [c_vp, c_vp, c_int]),
    "pb_host_plan_bilinear": (c_int, [c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp,
                                      c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pb_spectrum_reflected": (c_int, [c_vp, c_vp, c_vp]),
    "pb_spectrum_thermal": (c_int, [c_vp, c_vp, c_vp]),
    "pb_spectrum_transit": (c_int, [c_vp, c_vp, c_vp]),
    "pb_selftest_exp_tab": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "pb_optab_create": (c_int, [c_vp, c_int, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "pb_optab_destroy": (c_int, [c_vp, c_vp]),
    "pb_optab_set_molecular": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_int]),
    "pb_optab_set_continuum": (c_int, [c_vp, c_vp, c_int, c_vp, c_int]),
    "pb_optab_set_rayleigh": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "pb_optab_set_raman": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "pb_optab_set_ck": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int]),
    "pb_optab_bytes": (c_int, [c_vp, ctypes.POINTER(ctypes.c_size_t)]),
    "pb_compute_opacity": (c_int, [c_vp, c_vp, ctypes.POINTER(OpacityArgs), c_int]),
    "pb_ck_mix": (c_int, [c_vp, ctypes.POINTER(CkMixArgs), c_int]),
    "pb_climate_get_fluxes": (c_int, [c_vp, ctypes.POINTER(ClimateArgs), c_int]),
    "pb_ipc_export": (c_int, [c_vp, c_vp, c_vp]),
    "pb_ipc_open": (c_int, [c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "pb_ipc_close": (c_int, [c_vp, c_vp]),
    "pb_gather_wait": (c_int, [c_vp, c_vp, c_int, ctypes.c_uint64, c_vp]),
    "pb_regrid_plan_create": (c_int, [c_vp, c_int, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "pb_regrid_plan_destroy": (c_int, [c_vp, c_vp]),
    "pb_mean_regrid": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp, c_dbl, c_vp, c_int]),
}

_lib = None
_lock = threading.Lock()


def load_library():
    """dlopen the in-tree library and declare all prototypes; raises if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise PicasoB200Error(
                    "libpicaso_b200.so is not built (%s missing). Run `python -m picaso_b200.build`; "
                    "there is no CPU fallback." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


class Context:
    """One CUDA context/stream/staging arena per process and device (lazy: created on the
    first call, i.e. after any fork/spawn of joblib or MPI workers)."""

    def __init__(self, device=None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("PICASO_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            n = c_int(0)
            self.lib.pb_device_count(ctypes.byref(n))
            if n.value > 0:
                device %= n.value
        h = c_vp()
        rc = self.lib.pb_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise PicasoB200Error("pb_create(device=%d) failed: %s" %
                                  (device, self.lib.pb_last_error(None).decode()))
        self.h = h
        self.device = device
        self._pid = os.getpid()

    def check(self, rc):
        if rc != 0:
            raise PicasoB200Error(self.lib.pb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) is not None and self._pid == os.getpid():
            for ent in self.__dict__.pop("_ws", {}).values():  # context-owned workspaces
                self.lib.pb_dev_free(self.h, ent[0])
            self.lib.pb_destroy(self.h)
        self.h = None

    def microbench(self, which, iters=4096):
        """machine numbers of the fp64 pipe (pb_microbench): dict with ms, ops, Gop/s, cycles per iteration"""
        out = (ctypes.c_double * 4)()
        self.check(self.lib.pb_microbench(self.h, int(which), int(iters), out))
        ms, per_thread, threads, cyc = out[0], out[1], out[2], out[3]
        return {"ms": ms, "ops": per_thread * threads, "gops": per_thread * threads / (ms * 1e-3) / 1e9,
                "cycles_per_iter": cyc}

    def fp64_peak_tflops(self):
        """measured DFMA peak of this GPU (2 flop per FMA), TFLOP/s"""
        return 2.0 * self.microbench(0, 8192)["gops"] / 1e3

    # ---- helpers used by bench/tests ----
    def device_name(self):
        buf = ctypes.create_string_buffer(256)
        self.check(self.lib.pb_device_name(self.h, buf, 256))
        return buf.value.decode()

    def sm_count(self):
        n = c_int(0)
        self.check(self.lib.pb_sm_count(self.h, ctypes.byref(n)))
        return n.value

    def sync(self):
        self.check(self.lib.pb_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        self.check(self.lib.pb_set_stream(self.h, cuda_stream_ptr))

    def timer_start(self):
        self.check(self.lib.pb_timer_start(self.h))

    def timer_stop(self):
        ms = ctypes.c_float(0)
        self.check(self.lib.pb_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.pb_launch_count(self.h))

    def dev_alloc(self, nbytes):
        p = c_vp()
        self.check(self.lib.pb_dev_alloc(self.h, nbytes, ctypes.byref(p)))
        return p.value

    def dev_free(self, ptr):
        self.check(self.lib.pb_dev_free(self.h, ptr))

    def workspace(self, name, nbytes):
        """grow-only device buffer owned by the context and reused by the next call that asks for `name`
        (no cudaMalloc / cudaFree on the per-call path)"""
        ws = self.__dict__.setdefault("_ws", {})
        ent = ws.get(name)
        if ent is None or ent[1] < nbytes:
            if ent is not None:
                self.dev_free(ent[0])
            cap = max(int(nbytes), 8)
            ent = ws[name] = (self.dev_alloc(cap), cap)
        return ent[0]

    def to_device(self, arr):
        """copy a numpy float64 array to a fresh device buffer; returns the device address."""
        a = np.ascontiguousarray(arr, dtype=np.float64)
        d = self.dev_alloc(a.nbytes)
        self.check(self.lib.pb_memcpy_h2d(self.h, d, a.ctypes.data, a.nbytes))
        self.sync()
        return d

    def from_device(self, ptr, shape):
        out = np.empty(shape, dtype=np.float64)
        self.check(self.lib.pb_memcpy_d2h(self.h, out.ctypes.data, ptr, out.nbytes))
        self.sync()
        return out

    def pinned_empty(self, shape):
        """numpy float64 array backed by page-locked host memory (fast async H2D/D2H)."""
        n = int(np.prod(shape)) * 8
        p = c_vp()
        self.check(self.lib.pb_host_alloc(self.h, max(n, 8), ctypes.byref(p)))
        buf = (ctypes.c_char * max(n, 8)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float64, count=int(np.prod(shape))).reshape(shape)
        _PINNED[arr.ctypes.data] = (self, p.value)
        return arr

    def pinned_free(self, arr):
        ent = _PINNED.pop(arr.ctypes.data, None)
        if ent is not None:
            self.check(self.lib.pb_host_free(self.h, ent[1]))


_PINNED = {}
_default = None


def default_context():
    global _default
    if _default is None or _default._pid != os.getpid():
        _default = Context()
    return _default


def addr(a):
    """address of a numpy array (or pass through an int device pointer / None)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data
