"""Host-side mirror of the opacity half of picaso/optics.py for the GPU path.

* ``DeviceOpacities`` plays the role of ``optics.RetrieveOpacities`` (optics.py:1877-2368): it
  owns the cross-section tables - but uploaded ONCE into HBM instead of being re-read from
  sqlite per call - and ``get_opacities(atmosphere)`` only records which table rows / weights
  each layer needs (O(nlayer) host work; no data moves).
* ``compute_opacity(atmosphere, opacityclass, ...)`` has the reference's signature and 13-tuple
  return (optics.py:26-27, :423-431) and runs interpolation + mixing + Raman + delta-Eddington
  in one kernel.  With ``device_outputs=True`` the 13 arrays stay in HBM as ``DeviceArray``
  handles that ``picaso_b200.fluxes`` accepts directly, so a whole spectrum needs no
  O(nlayer x nwno) PCIe traffic.

The ``atmosphere`` argument is duck-typed on the attributes the reference reads from ATMSETUP:
``c.{nlayer,pconv,rgas,amu,k_b}``, ``level['temperature'|'pressure']``, ``layer['temperature'|
'pressure'|'colden'|'mmw'|'mixingratios'|'electrons'|'cloud']``, ``planet.gravity``, ``molecules``,
``continuum_molecules``, ``rayleigh_molecules``.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_DEVICE, PB_HOST, OpacityArgs, addr

__all__ = ["DeviceArray", "DeviceOpacities", "compute_opacity", "j_fraction"]

OUTPUT_NAMES = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "DTAU_OG", "TAU_OG", "W0_OG",
                "COSB_OG", "W0_no_raman", "f_deltaM")
_LEVEL = {"TAU", "TAU_OG"}


def grid_is_monotonic(t_inv_grid, p_log_grid):
    """(1/T grid strictly descending, log10 P grid strictly ascending): what every table of the reference looks like"""
    return (bool(t_inv_grid.size > 1 and np.all(t_inv_grid[1:] < t_inv_grid[:-1])),
            bool(p_log_grid.size > 1 and np.all(p_log_grid[1:] > p_log_grid[:-1])))


def find_needed_pts_grid(t_inv_grid, p_log_grid, nc_p, tlayer, player, monotonic=None):
    """RetrieveOpacities.find_needed_pts (picaso/optics.py:2048-2123) for all layers at once: the bilinear
    neighbours in (1/T, log10 P) of every layer on a (T-major, P-minor, possibly ragged) table grid.
    Returns t_interp[:, None], p_interp[:, None] and the four 0-based row indices (ll, hl, lh, hh).
    `monotonic` = grid_is_monotonic(...) if the caller has it cached."""
    t_inv = 1 / np.asarray(tlayer, dtype=np.float64)
    p_log = np.log10(np.asarray(player, dtype=np.float64))
    nT = t_inv_grid.size

    def last_true(mask):
        """per row: index of the last True (np.where(row)[0][-1]), 0 if none - any grid ordering"""
        n = mask.shape[1]
        return np.where(mask.any(axis=1), n - 1 - np.argmax(mask[:, ::-1], axis=1), 0)

    # last grid temperature strictly below T (1/T grid entry > 1/T), last grid pressure <= P.  On monotonic grids (every
    # table of the reference: temperatures ascending = 1/T descending, pressures ascending) a binary search counts the
    # entries that satisfy the comparison - the same comparisons, hence the same indices, as the masks
    t_mono, p_mono = grid_is_monotonic(t_inv_grid, p_log_grid) if monotonic is None else monotonic
    if t_mono:
        t_low = np.maximum(np.searchsorted(-t_inv_grid, -t_inv, side="left") - 1, 0)
    else:
        t_low = last_true(t_inv_grid[None, :] > t_inv[:, None])
    t_low = np.where(t_low == nT - 1, nT - 2, t_low)
    t_hi = t_low + 1
    if p_mono:
        p_low = np.maximum(np.searchsorted(p_log_grid, p_log, side="right") - 1, 0)
    else:
        p_low = last_true(p_log_grid[None, :] <= p_log[:, None])
    p_low = np.minimum(p_low, nc_p[t_hi] - 3)
    p_hi = p_low + 1
    off = np.concatenate([[0], np.cumsum(nc_p)])
    t_interp = ((t_inv - t_inv_grid[t_low]) / (t_inv_grid[t_hi] - t_inv_grid[t_low]))[:, np.newaxis]
    p_interp = ((p_log - p_log_grid[p_low]) / (p_log_grid[p_hi] - p_log_grid[p_low]))[:, np.newaxis]
    return (t_interp, p_interp, off[t_low] + p_low, off[t_hi] + p_low, off[t_low] + p_hi, off[t_hi] + p_hi)


class DeviceArray:
    """A float64 C-order array living in HBM (owned unless `owner` is given)."""

    def __init__(self, ctx, shape, ptr=None, owner=None):
        self.ctx, self.shape = ctx, tuple(int(s) for s in shape)
        self.nbytes = int(np.prod(self.shape)) * 8
        self._own = ptr is None
        self.ptr = ctx.dev_alloc(max(self.nbytes, 8)) if ptr is None else ptr
        self._owner = owner

    @classmethod
    def from_numpy(cls, ctx, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        d = cls(ctx, a.shape)
        ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, d.ptr, a.ctypes.data, a.nbytes))
        ctx.sync()
        return d

    @property
    def ndim(self):
        return len(self.shape)

    def numpy(self):
        return self.ctx.from_device(self.ptr, self.shape)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, key):
        """supports the X[:, :, ig] slices picaso() takes (justdoit.py:275-283) for ngauss = 1"""
        if (isinstance(key, tuple) and len(key) == 3 and key[0] == slice(None) and key[1] == slice(None)
                and key[2] in (0, -1) and (self.ndim == 2 or self.shape[2] == 1)):
            return DeviceArray(self.ctx, self.shape[:2], ptr=self.ptr, owner=self)
        return self.numpy()[key]

    def free(self):
        if self._own and self.ptr is not None:
            self.ctx.dev_free(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            if self._own and self.ptr is not None and self.ctx.h is not None:
                self.ctx.dev_free(self.ptr)
        except Exception:
            pass


def _partition_function(j, T):
    # optics.py:541-549, as coded (b_energy already carries j(j+1))
    k, b, c, h = 1.38064852e-16, 60.853, 29979245800, 6.62607004e-27
    b_energy = (b * (h) * (c) * j * (j + 1) / k)
    g = (2.0 * j + 1.0) if j % 2 == 0 else 3.0 * (2.0 * j + 1.0)
    return g * np.exp(-0.5 * b_energy * j * (j + 1) / T)


def j_fraction(j, T):
    """fraction of H2 in rotational level J at temperature(s) T (optics.py:552-581)."""
    T = np.asarray(T, dtype=np.float64)
    Z = np.zeros(T.shape)
    for jj in range(20):
        Z += _partition_function(jj, T)
    return _partition_function(j, T) / Z


class DeviceOpacities:
    """GPU-resident stand-in for optics.RetrieveOpacities (monochromatic opacities, ngauss = 1).

    Parameters mirror what the reference reads from its sqlite DB (optics.py:1998-2046):
    wno [W]; pt_pairs = [(ptid, pressure_bar, temperature), ...] in ptid order (T-major);
    tables {molecule: [npt, W]}; cia_temps [nTc] + continuum {pair: [nTc, W]}; rayleigh_opa
    {molecule: [W]}; raman_db = (c, ji, deltanu); query_method 'nearest' (reference default)
    or 'linear'."""

    ngauss = 1

    def __init__(self, wno, pt_pairs, tables, cia_temps, continuum, rayleigh_opa, raman_db=None,
                 query_method="nearest", ctx=None):
        if query_method not in ("nearest", "linear"):
            raise Exception(f"Do not recognize query method for opacities: {query_method}. Options are nearest or linear")
        self.ctx = ctx or _lib.default_context()
        self.query_method = query_method
        self.wno = np.ascontiguousarray(wno, dtype=np.float64)
        self.wave = 1e4 / self.wno
        self.nwno = self.wno.size
        self.gauss_wts = np.array([1])
        self.pt_pairs = [(int(p[0]), float(p[1]), float(p[2])) for p in pt_pairs]
        self._ptid = np.array([p[0] for p in self.pt_pairs])
        P = np.array([p[1] for p in self.pt_pairs])
        T = np.array([p[2] for p in self.pt_pairs])
        self._lnP, self._T = np.log(P), T
        # grid description as get_available_data builds it (optics.py:2019-2025)
        self.temps = T[np.sort(np.unique(T, return_index=True)[1])]
        self.pressures = P[np.sort(np.unique(P, return_index=True)[1])]
        self.nc_p = np.array([np.sum(T == t) for t in np.unique(T)])
        self.t_inv_grid = 1 / self.temps
        self.p_log_grid = np.log10(self.pressures)
        self.molecules = np.array(list(tables.keys()))
        self._mol_index = {m: i for i, m in enumerate(tables)}
        self.cia_temps = np.asarray(cia_temps, dtype=np.float64)
        self._cia_unique = np.unique(self.cia_temps)
        self.avail_continuum = list(continuum.keys())
        self._cont_index = {k: i for i, k in enumerate(continuum)}
        self.rayleigh_molecules = list(rayleigh_opa.keys())
        self.rayleigh_opa = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in rayleigh_opa.items()}
        self._ray_index = {k: i for i, k in enumerate(rayleigh_opa)}
        self.raman_db = raman_db
        self._shifts = None
        self.preload = True
        lib, h = self.ctx.lib, self.ctx.h
        tab = ctypes.c_void_p()
        self.ctx.check(lib.pb_optab_create(h, self.nwno, len(tables), len(continuum), len(rayleigh_opa),
                                           ctypes.byref(tab)))
        self._tab = tab
        store = 3 if query_method == "linear" else 1
        for m, t in tables.items():
            t = np.ascontiguousarray(t, dtype=np.float64)
            if t.shape != (len(self.pt_pairs), self.nwno):
                raise ValueError(f"table of {m} has shape {t.shape}, expected {(len(self.pt_pairs), self.nwno)}")
            self.ctx.check(lib.pb_optab_set_molecular(h, tab, self._mol_index[m], addr(t), t.shape[0], store))
        for k, t in continuum.items():
            t = np.ascontiguousarray(t, dtype=np.float64)
            order = np.argsort(self.cia_temps)  # rows in ascending unique-temperature order
            t = np.ascontiguousarray(t[order])
            self.ctx.check(lib.pb_optab_set_continuum(h, tab, self._cont_index[k], addr(t), t.shape[0]))
        for k, s in self.rayleigh_opa.items():
            self.ctx.check(lib.pb_optab_set_rayleigh(h, tab, self._ray_index[k], addr(s)))
        self._plan = None
        self._ws = {}   # pooled device buffers for compute_opacity(device_outputs=True)

    def _buffer(self, name, shape):
        """pooled DeviceArray: reused by the next call on this connection (no cudaMalloc churn)"""
        d = self._ws.get(name)
        if d is None or d.shape != tuple(shape) or d.ptr is None:
            if d is not None:
                d.free()
            d = DeviceArray(self.ctx, shape)
            self._ws[name] = d
        return d

    # ---- Raman stellar shifts (star(), justdoit.py:1756-1913 sets opa.raman_stellar_shifts) ----
    @property
    def raman_stellar_shifts(self):
        return self._shifts

    @raman_stellar_shifts.setter
    def raman_stellar_shifts(self, shifts):
        self._shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        if self.raman_db is None:
            raise ValueError("raman_db = (c, ji, deltanu) is required before setting raman_stellar_shifts")
        c, ji, dnu = (np.asarray(self.raman_db[k] if isinstance(self.raman_db, dict) else self.raman_db[i])
                      for i, k in enumerate(("c", "ji", "deltanu")))
        c = np.ascontiguousarray(c, dtype=np.float64)
        dnu = np.ascontiguousarray(dnu, dtype=np.float64)
        ji = np.ascontiguousarray(ji, dtype=np.int32)
        if self._shifts.shape != (self.nwno, c.size):
            raise ValueError("raman_stellar_shifts must be [nwno, ntransitions]")
        self._raman = (c, ji, dnu)
        self.ctx.check(self.ctx.lib.pb_optab_set_raman(self.ctx.h, self._tab, addr(self.wno), c.size, addr(c),
                                                       addr(ji), addr(dnu), addr(self._shifts)))

    def device_bytes(self):
        n = ctypes.c_size_t(0)
        self.ctx.lib.pb_optab_bytes(self._tab, ctypes.byref(n))
        return n.value

    # ---- which rows does an atmosphere need (optics.py:2048-2123, :2330-2332, :2298) ----------
    def find_needed_pts(self, tlayer, player):
        """bilinear neighbours in (1/T, log10 P); same return convention as the reference:
        t_interp[:,None], p_interp[:,None], and the four 0-based row indices."""
        mono = self.__dict__.get("_grid_mono")
        if mono is None:
            mono = self._grid_mono = grid_is_monotonic(self.t_inv_grid, self.p_log_grid)
        return find_needed_pts_grid(self.t_inv_grid, self.p_log_grid, self.nc_p, tlayer, player, monotonic=mono)

    def _plan_numpy(self, tlayer, pbar):
        """the bilinear plan in numpy statements (find_needed_pts_grid + the bookkeeping of optics.py:2265-2298)"""
        t, p, ill, ihl, ilh, ihh = self.find_needed_pts(tlayer, pbar)
        t, p = t[:, 0], p[:, 0]
        idx = np.stack([ill, ihl, ihh, ilh], axis=1).astype(np.int32)
        t1, p1 = 1 - t, 1 - p
        wts = np.stack([t1 * p1, t * p1, t * p, t1 * p], axis=1)
        cia = np.abs(self._cia_unique[None, :] - tlayer[:, None]).argmin(axis=1).astype(np.int32)
        return idx, wts, cia, 1 + np.unique(idx).astype(np.int64)

    def get_opacities(self, atmosphere, exclude_mol=1):
        """Record the table rows / weights for this atmosphere; nothing is fetched or copied.
        Sets atmosphere.layer['pt_opa_index'] like the reference (optics.py:2265, :2333)."""
        tlayer = np.asarray(atmosphere.layer["temperature"], dtype=np.float64)
        pbar = np.asarray(atmosphere.layer["pressure"], dtype=np.float64) / atmosphere.c.pconv
        L = tlayer.size
        cia = None
        if self.query_method == "linear":
            # one host-side C call (csrc/host_plan.cu) instead of ~40 small numpy calls: same comparisons and IEEE
            # operations as find_needed_pts_grid, bit-identical plan (tests/test_host_plan_cpu.py)
            g = self.__dict__.get("_plan_grid")
            if g is None or g["L"] != L:
                tg = np.ascontiguousarray(self.t_inv_grid, dtype=np.float64)
                pg = np.ascontiguousarray(self.p_log_grid, dtype=np.float64)
                ncp = np.ascontiguousarray(self.nc_p, dtype=np.int64)
                off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(ncp)]), dtype=np.int64)
                cu = np.ascontiguousarray(self._cia_unique, dtype=np.float64)
                t_mono, p_mono = grid_is_monotonic(tg, pg)
                # output buffers of the plan are owned by the connection and reused by the next get_opacities call (the
                # plan always describes the LAST atmosphere); their addresses are looked up once (ndarray.ctypes is slow)
                out = (np.empty((L, 4), dtype=np.int32), np.empty((L, 4)), np.empty(L, dtype=np.int32),
                       np.empty(4 * L, dtype=np.int64), ctypes.c_int(0))
                g = self._plan_grid = dict(
                    L=L, keep=(tg, pg, ncp, off, cu), out=out, fn=_lib.load_library().pb_host_plan_bilinear,
                    static=(tg.size, tg.ctypes.data, pg.size, pg.ctypes.data, ncp.ctypes.data, off.ctypes.data, int(t_mono),
                            int(p_mono), cu.size, cu.ctypes.data, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data,
                            out[3].ctypes.data, ctypes.addressof(out[4])))
                g["checked"] = False
            if g["fn"] is not None:
                tl = np.ascontiguousarray(tlayer)
                t_inv = 1 / tl
                p_log = np.log10(pbar)
                idx, wts, cia, rows, nrows = g["out"]
                rc = g["fn"](L, t_inv.ctypes.data, p_log.ctypes.data, tl.ctypes.data, *g["static"])
                if rc != 0:
                    raise _lib.PicasoB200Error("pb_host_plan_bilinear: bad table grid (needs >= 2 temperatures and pressures)")
                used = rows[:nrows.value].copy()
                if not g["checked"]:
                    # once per connection: the C planner against the numpy statements it replaces, on this very profile
                    g["checked"] = True
                    i2, w2, c2, u2 = self._plan_numpy(tlayer, pbar)
                    if not (np.array_equal(idx, i2) and np.array_equal(wts, w2, equal_nan=True) and np.array_equal(cia, c2)
                            and np.array_equal(used, u2)):
                        import warnings
                        warnings.warn("picaso_b200: pb_host_plan_bilinear disagrees with the numpy planner on this table grid; "
                                      "using the numpy planner for this connection")
                        g["fn"] = None
            if g["fn"] is None:
                idx, wts, cia, used = self._plan_numpy(tlayer, pbar)
            atmosphere.layer["pt_opa_index"] = used
        else:
            idx = np.zeros((L, 4), dtype=np.int32)
            wts = np.zeros((L, 4))
            rows = np.argmin(np.hypot(self._lnP[None, :] - np.log(pbar)[:, None], self._T[None, :] - tlayer[:, None]), axis=1)
            idx[:, 0] = rows
            atmosphere.layer["pt_opa_index"] = [int(self._ptid[r]) for r in rows]
        if cia is None:
            cia = np.abs(self._cia_unique[None, :] - tlayer[:, None]).argmin(axis=1).astype(np.int32)
        if np.isscalar(exclude_mol) and exclude_mol == 1:
            fac = dict.fromkeys(atmosphere.molecules, 1)
        else:
            fac = {m: exclude_mol[m] for m in atmosphere.molecules}
        self._plan = dict(idx=idx, wts=wts, cia=cia, fac=fac, nlayer=L)
        # the reference exposes dicts of [nlayer, nwno] arrays here; on this path they never exist
        self.molecular_opa = None
        self.continuum_opa = None

    def close(self):
        for d in getattr(self, "_ws", {}).values():
            d.free()
        self._ws = {}
        if getattr(self, "_tab", None) is not None and self.ctx.h is not None:
            self.ctx.lib.pb_optab_destroy(self.ctx.h, self._tab)
        self._tab = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _layer_scalars(atm, opa):
    """per-layer multipliers exactly as compute_opacity parenthesises them (optics.py:147-271).  The species of a
    group (molecules, Rayleigh scatterers, CIA pairs) are stacked so that each group costs one numpy expression -
    elementwise the same operations in the same order as the per-species statements of the reference
    (tests/test_host_plan_cpu.py holds the result to the statement-per-species version bit for bit)."""
    L = atm.c.nlayer
    tlevel = np.asarray(atm.level["temperature"], dtype=np.float64)
    plevel = np.asarray(atm.level["pressure"], dtype=np.float64) / atm.c.pconv
    tlayer = np.asarray(atm.layer["temperature"], dtype=np.float64)
    gravity = atm.planet.gravity / 100.0
    mmw = np.asarray(atm.layer["mmw"], dtype=np.float64)
    colden = np.asarray(atm.layer["colden"], dtype=np.float64)
    player = np.asarray(atm.layer["pressure"], dtype=np.float64)
    mix = atm.layer["mixingratios"]
    if isinstance(mix, dict):
        x = mix.__getitem__                      # numpy vectors (np.array below converts anything else)
    else:
        x = lambda s: mix[s].values              # pandas DataFrame, as ATMSETUP builds it
    t0, t1, p0, p1 = tlevel[:-1], tlevel[1:], plevel[:-1], plevel[1:]
    dp = p1 - p0
    tt = tlayer / (t0 * t1)
    ACOEF = tt * (t1 * p1 - t0 * p0) / dp
    BCOEF = tt * (t0 - t1) / dp
    COEF1 = atm.c.rgas * 273.15 ** 2 * .5E5 * (
        ACOEF * (p1 ** 2 - p0 ** 2) + BCOEF * (2. / 3.) * (p1 ** 3 - p0 ** 3)) / (
        1.01325 ** 2 * gravity * tlayer * mmw)
    cont = np.zeros((len(opa._cont_index), L))
    cia_rows, cia_a, cia_b = [], [], []
    for m in atm.continuum_molecules:
        key = m[0] + m[1]
        if key not in opa._cont_index:
            raise KeyError(f"continuum pair {key} is not in the uploaded tables")
        if m[0] == "H-" and m[1] == "bf":
            cont[opa._cont_index[key]] = (np.asarray(x("H-"), dtype=np.float64) * colden / (mmw * atm.c.amu))
        elif m[0] == "H-" and m[1] == "ff":
            cont[opa._cont_index[key]] = (player * np.asarray(x("H"), dtype=np.float64) * np.asarray(atm.layer["electrons"]) *
                                          colden / (tlayer * mmw * atm.c.amu * atm.c.k_b))
        elif m[0] == "H2-" and m[1] == "":
            cont[opa._cont_index[key]] = (player * np.asarray(x("H2"), dtype=np.float64) * np.asarray(atm.layer["electrons"]) *
                                          colden / (mmw * atm.c.amu))
        else:
            cia_rows.append(opa._cont_index[key])
            cia_a.append(x(m[0]))
            cia_b.append(x(m[1]))
    if cia_rows:
        cont[cia_rows] = (COEF1 * np.array(cia_a, dtype=np.float64) * np.array(cia_b, dtype=np.float64))
    mol = np.zeros((len(opa._mol_index), L))
    if len(atm.molecules):
        try:
            rows = [opa._mol_index[m] for m in atm.molecules]
        except KeyError as e:
            raise KeyError(f"molecule {e.args[0]} is not in the uploaded tables") from None
        fac = np.array([opa._plan["fac"][m] for m in atm.molecules])[:, None]
        mol[rows] = fac * (colden * np.array([x(m) for m in atm.molecules], dtype=np.float64) / mmw)
    ray = np.zeros((len(opa._ray_index), L))
    if len(atm.rayleigh_molecules):
        ray[[opa._ray_index[m] for m in atm.rayleigh_molecules]] = (
            colden * np.array([x(m) for m in atm.rayleigh_molecules], dtype=np.float64) / mmw)
    return mol, cont, ray


def pollack_factor(opa):
    """raman_pollack(nlayer, 1e4 / wno)[0] of the reference (optics.py:652-660): np.interp of the two-column
    table raman_fortran.txt (wavelength, factor) onto the connection's wavelength grid; cached on `opa`."""
    cached = getattr(opa, "_pollack", None)
    if cached is not None:
        return cached
    tab = getattr(opa, "raman_pollack_table", None)
    if tab is None:
        import os
        root = os.environ.get("picaso_refdata")
        path = os.path.join(root, "opacities", "raman_fortran.txt") if root else None
        if not path or not os.path.isfile(path):
            raise FileNotFoundError("raman='pollack' reads $picaso_refdata/opacities/raman_fortran.txt (as the reference does, "
                                    "optics.py:652); set picaso_refdata or opacityclass.raman_pollack_table = (w, f)")
        dat = np.loadtxt(path)
        tab = (dat[:, 0], dat[:, 1])
    w, f = (np.asarray(x, dtype=np.float64) for x in tab)
    opa._pollack = np.ascontiguousarray(np.interp(1e4 / np.asarray(opa.wno, dtype=np.float64), w, f))
    return opa._pollack


def _fill_opacity_args(atm, opa, stream, delta_eddington, raman, fthin_cld, do_holes, device_outputs, into=None,
                       test_mode=None):
    """Everything pb_compute_opacity reads except the output pointers: table rows / weights of the atmosphere (from
    opacityclass.get_opacities), per-layer multipliers, Raman inputs, clouds.  Returns (args, memspace, keep): `keep`
    holds the host arrays whose addresses were taken and must outlive the call; `into` fills an existing struct (the
    `opacity` member of a SpectrumArgs)."""
    ctx = opa.ctx
    if opa._plan is None or opa._plan["nlayer"] != atm.c.nlayer:
        raise RuntimeError("call opacityclass.get_opacities(atmosphere) first (justdoit.py:236)")
    L, W = atm.c.nlayer, opa.nwno
    mol, cont, ray = _layer_scalars(atm, opa)
    a = OpacityArgs() if into is None else into
    a.nlayer = L
    a.query = 1 if opa.query_method == "linear" else 0
    idx, wts, cia = opa._plan["idx"], opa._plan["wts"], opa._plan["cia"]
    a.pt_index, a.weights, a.cont_index = addr(idx), addr(wts), addr(cia)
    a.mol_scale, a.cont_scale, a.ray_scale = addr(mol), addr(cont), addr(ray)
    a.raman = int(raman)
    keep = [idx, wts, cia, mol, cont, ray]
    if raman == 0:
        if opa.raman_stellar_shifts is None:
            raise RuntimeError("raman=0 needs opacityclass.raman_stellar_shifts (set by star())")
        jf = np.ascontiguousarray([j_fraction(j, np.asarray(atm.layer["temperature"])) for j in range(10)])
        a.jfrac = addr(jf)
        keep.append(jf)
    elif raman == 1:
        pol = pollack_factor(opa)
        if device_outputs:
            # raman_pollack follows memspace (include/picaso_b200.h): the [nwno] factor goes to HBM once per connection
            dpol = getattr(opa, "_pollack_dev", None)
            if dpol is None:
                dpol = opa._buffer("raman_pollack", (W,))
                ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, dpol.ptr, pol.ctypes.data, pol.nbytes))
                opa._pollack_dev = dpol
            a.raman_pollack = dpol.ptr
        else:
            a.raman_pollack = addr(pol)
        keep.append(pol)
    cloud = atm.layer.get("cloud") if isinstance(atm.layer, dict) else atm.layer["cloud"]
    memspace = PB_DEVICE if device_outputs else PB_HOST
    tmp_dev = []
    a.test_mode = 0
    if test_mode not in (None, False):
        # optics.py:372-399: 'rayleigh', or anything else = cloud-only.  The reference replaces non-positive cloud
        # single-scattering albedos by 1e-10 in the atmosphere's own array (:393); so does this mirror.
        a.test_mode = 1 if test_mode == "rayleigh" else 2
        if cloud is None:
            raise TypeError("compute_opacity test modes read atmosphere.layer['cloud'] (optics.py:386-395)")
        w0c = cloud["w0"]
        if isinstance(w0c, np.ndarray) and w0c.flags.writeable:
            w0c[w0c <= 0] = 1e-10
        else:
            cloud = dict(cloud, w0=np.where(np.asarray(w0c, dtype=np.float64) <= 0, 1e-10, w0c))
    if cloud is not None and (a.test_mode or np.any(np.asarray(cloud["opd"]) != 0)):
        cl = [np.asarray(cloud[k], dtype=np.float64) for k in ("opd", "w0", "g0")]
        cl = [c if (c.shape == (L, W) and c.flags.c_contiguous) else
              np.ascontiguousarray(np.broadcast_to(c, (L, W))) for c in cl]
        if device_outputs:
            dcl = [opa._buffer("cloud_" + k, (L, W)) for k in ("opd", "w0", "g0")]
            for d, c in zip(dcl, cl):
                ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, d.ptr, c.ctypes.data, c.nbytes))
            keep += cl
            a.cloud_opd, a.cloud_w0, a.cloud_g0 = [d.ptr for d in dcl]
        else:
            a.cloud_opd, a.cloud_w0, a.cloud_g0 = [addr(c) for c in cl]
            keep += cl
        a.cloud_ld = W
    a.fthin_cld = float(fthin_cld) if fthin_cld is not None else 0.0
    a.do_holes = int(bool(do_holes))
    a.stream, a.delta_eddington = int(stream), int(bool(delta_eddington))
    return a, memspace, keep


def compute_opacity(atmosphere, opacityclass, ngauss=1, stream=2, delta_eddington=True, test_mode=False,
                    raman=0, plot_opacity=False, full_output=False, return_mode=False, fthin_cld=None,
                    do_holes=False, *, device_outputs=False, outputs=None):
    """CUDA replacement of optics.compute_opacity (optics.py:26-431) for a ``DeviceOpacities``
    connection: returns the reference's 13-tuple (DTAU, TAU, W0, COSB, ftau_cld, ftau_ray, GCOS2,
    DTAU_OG, TAU_OG, W0_OG, COSB_OG, W0_no_raman, f_deltaM), each [nlayer|nlevel, nwno, 1] like the
    reference's ngauss axis.  ``device_outputs=True`` returns ``DeviceArray`` handles instead
    (2-D, sliceable with [:, :, 0]) that live in buffers pooled on the connection - they are
    overwritten by the next compute_opacity call on the same ``DeviceOpacities``; ``outputs``
    restricts the computed set to the given names (others are returned as None).

    ``raman=1`` ('pollack', the reference's config default): the per-wavelength factor is the reference's
    np.interp(1e4 / wno, w, f) of ``$picaso_refdata/opacities/raman_fortran.txt`` (optics.py:584-660), read once
    per connection, or of ``opacityclass.raman_pollack_table = (w, f)`` if that is set.
    ``full_output=True`` sets ``atmosphere.taugas / tauray / taucld`` ([nlayer, nwno, 1] numpy arrays) like the
    reference (optics.py:322-325).
    ``test_mode='rayleigh'`` / any other string: the reference's test modes (optics.py:372-399) inside the same kernel;
    like the reference they replace non-positive ``layer['cloud']['w0']`` entries by 1e-10 in the caller's array.
    ``test_mode`` None/False both mean "normal run": the reference's own default False would enter its test
    branch (optics.py:372), real callers pass None.
    Not supported on this path: plot_opacity, return_mode (host-side plotting)."""
    from .optics_ck import DeviceCKs, compute_opacity_ck
    if not isinstance(opacityclass, (DeviceOpacities, DeviceCKs)):
        raise TypeError("picaso_b200.compute_opacity needs a DeviceOpacities or DeviceCKs connection")
    if isinstance(opacityclass, DeviceOpacities) and ngauss != 1:
        raise ValueError("monochromatic DeviceOpacities have ngauss = 1")
    if isinstance(opacityclass, DeviceCKs) and ngauss != opacityclass.ngauss:
        raise ValueError("ngauss must equal the number of gauss points of the DeviceCKs table")
    if plot_opacity or return_mode:
        raise NotImplementedError("plot_opacity / return_mode are host-side diagnostics")
    opa, atm = opacityclass, atmosphere
    if isinstance(opa, DeviceCKs):
        if test_mode not in (None, False):
            raise NotImplementedError("compute_opacity test modes are implemented for monochromatic opacities (ngauss = 1)")
        import copy
        atm_ck = copy.copy(atm)
        atm_ck.molecules = []   # pre-mixed tables already contain every molecule (optics.py:257-262)
        return compute_opacity_ck(atm_ck, opa, stream, delta_eddington, raman, fthin_cld, do_holes,
                                  device_outputs, outputs)
    a, memspace, keep = _fill_opacity_args(atm, opa, stream, delta_eddington, raman, fthin_cld, do_holes, device_outputs,
                                           test_mode=test_mode)
    ctx = opa.ctx
    L, W = atm.c.nlayer, opa.nwno
    want = set(OUTPUT_NAMES if outputs is None else outputs)
    res = {}
    for n in OUTPUT_NAMES:
        if n not in want:
            res[n] = None
            continue
        shape = (L + 1, W) if n in _LEVEL else (L, W)
        if device_outputs:
            res[n] = opa._buffer(n, shape)
            setattr(a, n, res[n].ptr)
        else:
            res[n] = np.zeros(shape)
            setattr(a, n, addr(res[n]))
    extra = None
    if full_output:
        # atmosphere.taugas / tauray / taucld (optics.py:322-325): three more [nlayer, nwno] outputs of the same launch
        if device_outputs:
            extra = [opa._buffer(n, (L, W)) for n in ("TAUGAS", "TAURAY", "TAUCLD")]
            a.TAUGAS, a.TAURAY, a.TAUCLD = [x.ptr for x in extra]
        else:
            extra = [np.zeros((L, W)) for _ in range(3)]
            a.TAUGAS, a.TAURAY, a.TAUCLD = [addr(x) for x in extra]
    ctx.check(ctx.lib.pb_compute_opacity(ctx.h, opa._tab, ctypes.byref(a), memspace))
    if device_outputs:
        ctx.sync()   # the host staging arrays in `keep` may be released after this point
    if extra is not None:
        host3 = [x.numpy() if device_outputs else x for x in extra]
        atm.taugas, atm.tauray, atm.taucld = [x[:, :, np.newaxis] for x in host3]
    if device_outputs:
        return tuple(res[n] for n in OUTPUT_NAMES)
    return tuple(None if res[n] is None else res[n][:, :, np.newaxis] for n in OUTPUT_NAMES)


def reflected_spectrum(atmosphere, opacityclass, ubar0, ubar1, cos_theta, gweight, tweight, *, F0PI=None, surf_reflect=None,
                       b_top=None, single_phase=3, multi_phase=0, frac_a=1.0, frac_b=-1.0, frac_c=2.0, constant_back=-0.5,
                       constant_forward=1.0, toon_coefficients=0, stream=2, delta_eddington=True, raman=2, fthin_cld=None,
                       do_holes=False, return_xint=False):
    """One call per reflected-light spectrum (pb_spectrum_reflected): what picaso() does between
    ``opacityclass.get_opacities(atm)`` and ``returns['albedo']`` for a Toon run - compute_opacity, get_reflected_1d,
    compress_disco (justdoit.py:243-310, :530) - with the 11 opacity arrays living only in HBM.  Call
    ``opacityclass.get_opacities(atmosphere)`` first, exactly as picaso() does.  ``F0PI`` / ``surf_reflect`` / ``b_top``:
    None (= 1 / 0 / 0), a scalar, or a [nwno] vector (uploaded once per distinct array object and cached on the
    connection).  Defaults are the reference's config.json (TTHG_ray, N=2, quadrature, delta-Eddington, no Raman).
    Returns albedo[nwno] (and xint_at_top[ng, nt, nwno] with ``return_xint=True``)."""
    from ._lib import SpectrumArgs
    opa, atm = opacityclass, atmosphere
    if not isinstance(opa, DeviceOpacities):
        raise TypeError("reflected_spectrum needs a DeviceOpacities connection")
    sa = SpectrumArgs()
    _, _, keep = _fill_opacity_args(atm, opa, stream, delta_eddington, raman, fthin_cld, do_holes, True, into=sa.opacity)
    ctx, W = opa.ctx, opa.nwno
    u0 = np.ascontiguousarray(ubar0, dtype=np.float64)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64)
    ng, nt = u0.shape if u0.ndim == 2 else (u0.size, 1)
    u0, u1 = u0.reshape(-1), u1.reshape(-1)
    gw = np.ascontiguousarray(gweight, dtype=np.float64)
    tw = np.ascontiguousarray(tweight, dtype=np.float64)
    sa.nwno, sa.numg, sa.numt = W, ng, nt
    sa.ubar0, sa.ubar1, sa.gweight, sa.tweight = addr(u0), addr(u1), addr(gw), addr(tw)
    sa.cos_theta = float(cos_theta)

    def wave_vector(name, value, default):
        """device pointer of a per-wavelength vector, or None when it equals the kernel's default"""
        if value is None or (np.isscalar(value) and value == default):
            return None
        cache = opa.__dict__.setdefault("_wave_vectors", {})
        hit = cache.get(name)
        if hit is not None and hit[0] is value:
            return hit[1].ptr
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (W,)))
        d = opa._buffer("spectrum_" + name, (W,))
        ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, d.ptr, v.ctypes.data, v.nbytes))
        ctx.sync()
        cache[name] = (value, d)
        return d.ptr

    sa.surf_reflect = wave_vector("surf_reflect", surf_reflect, 0.0)
    sa.F0PI = wave_vector("F0PI", F0PI, 1.0)
    sa.b_top = wave_vector("b_top", b_top, 0.0)
    sa.single_phase, sa.multi_phase, sa.toon_coefficients = int(single_phase), int(multi_phase), int(toon_coefficients)
    sa.frac_a, sa.frac_b, sa.frac_c = float(frac_a), float(frac_b), float(frac_c)
    sa.constant_back, sa.constant_forward = float(constant_back), float(constant_forward)
    alb = np.empty(W)
    xint = np.empty((ng, nt, W)) if return_xint else None
    sa.albedo, sa.xint_at_top = addr(alb), addr(xint)
    ctx.check(ctx.lib.pb_spectrum_reflected(ctx.h, opa._tab, ctypes.byref(sa)))
    return (alb, xint) if return_xint else alb


def _wno_on_device(opa):
    """the connection's wavenumber grid in HBM (uploaded once): pb_thermal_toon_1d reads it with PB_DEVICE"""
    d = getattr(opa, "_wno_dev", None)
    if d is None:
        d = DeviceArray.from_numpy(opa.ctx, opa.wno)
        opa._wno_dev = d
    return d.ptr


def thermal_spectrum(atmosphere, opacityclass, ubar1, gweight, tweight, *, surf_reflect=None, hard_surface=0, stream=2,
                     delta_eddington=True, raman=2, fthin_cld=None, do_holes=False, return_flux=False):
    """One call per thermal-emission spectrum (pb_spectrum_thermal): compute_opacity -> get_thermal_1d on DTAU_OG,
    W0_no_raman, COSB_OG -> compress_thermal, as picaso() chains them (justdoit.py:243, :337-342, :567); temperatures and
    pressures are ``atmosphere.level``'s.  Call ``opacityclass.get_opacities(atmosphere)`` first.  Returns thermal[nwno]
    (and flux_at_top[ng, nt, nwno] with ``return_flux=True``)."""
    from ._lib import SpectrumThermalArgs
    opa, atm = opacityclass, atmosphere
    if not isinstance(opa, DeviceOpacities):
        raise TypeError("thermal_spectrum needs a DeviceOpacities connection")
    sa = SpectrumThermalArgs()
    _, _, keep = _fill_opacity_args(atm, opa, stream, delta_eddington, raman, fthin_cld, do_holes, True, into=sa.opacity)
    ctx, W = opa.ctx, opa.nwno
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64)
    ng, nt = u1.shape if u1.ndim == 2 else (u1.size, 1)
    u1 = u1.reshape(-1)
    gw = np.ascontiguousarray(gweight, dtype=np.float64)
    tw = np.ascontiguousarray(tweight, dtype=np.float64)
    tl = np.ascontiguousarray(atm.level["temperature"], dtype=np.float64)
    pl = np.ascontiguousarray(atm.level["pressure"], dtype=np.float64)
    sa.nwno, sa.numg, sa.numt = W, ng, nt
    sa.tlevel, sa.plevel, sa.ubar1, sa.gweight, sa.tweight = addr(tl), addr(pl), addr(u1), addr(gw), addr(tw)
    sa.wno = _wno_on_device(opa)
    sa.surf_reflect = None
    if surf_reflect is not None and not (np.isscalar(surf_reflect) and surf_reflect == 0):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(surf_reflect, dtype=np.float64), (W,)))
        d = opa._buffer("spectrum_thermal_surf", (W,))
        ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, d.ptr, v.ctypes.data, v.nbytes))
        ctx.sync()
        sa.surf_reflect = d.ptr
    sa.hard_surface = int(hard_surface)
    th = np.empty(W)
    ft = np.empty((ng, nt, W)) if return_flux else None
    sa.thermal, sa.flux_at_top = addr(th), addr(ft)
    ctx.check(ctx.lib.pb_spectrum_thermal(ctx.h, opa._tab, ctypes.byref(sa)))
    return (th, ft) if return_flux else th


def transit_spectrum(atmosphere, opacityclass, rstar, *, z=None, dz=None, stream=2, delta_eddington=True, raman=2,
                     fthin_cld=None, do_holes=False):
    """One call per transmission spectrum (pb_spectrum_transit): compute_opacity -> get_transit_1d on DTAU_OG
    (justdoit.py:243, :388-396).  z / dz default to ``atmosphere.level['z']`` / ``['dz']``; like picaso() the LEVEL
    pressures and temperatures are what get_transit_1d receives.  Returns (rp/rs)^2 [nwno]."""
    from ._lib import SpectrumTransitArgs
    opa, atm = opacityclass, atmosphere
    if not isinstance(opa, DeviceOpacities):
        raise TypeError("transit_spectrum needs a DeviceOpacities connection")
    sa = SpectrumTransitArgs()
    _, _, keep = _fill_opacity_args(atm, opa, stream, delta_eddington, raman, fthin_cld, do_holes, True, into=sa.opacity)
    ctx, W = opa.ctx, opa.nwno
    vec = [np.ascontiguousarray(x, dtype=np.float64) for x in (
        atm.level["z"] if z is None else z, atm.level["dz"] if dz is None else dz, atm.level["pressure"],
        atm.level["temperature"], atm.layer["mmw"], atm.layer["colden"])]
    sa.nwno = W
    sa.z, sa.dz, sa.player, sa.tlayer, sa.mmw, sa.colden = [addr(v) for v in vec]
    sa.rstar, sa.k_b, sa.amu = float(rstar), float(atm.c.k_b), float(atm.c.amu)
    F = np.empty(W)
    sa.F = addr(F)
    ctx.check(ctx.lib.pb_spectrum_transit(ctx.h, opa._tab, ctypes.byref(sa)))
    return F
