"""Host mirror of the climate solver's radiative-transfer call.

`get_fluxes` has the signature and return value of the reference's
``picaso.climate.get_fluxes`` (climate.py:1686-1952): it takes the solver's namedtuples
(``Atmosphere``, ``OpacityWEd``, ``OpacityNoEd``, ``ScatteringPhase``, ``Disco``, ``Opagrid``) - anything
with those attributes works - and returns

    flux_net_v_layer [ng,nt,nlevel], flux_net_v [ng,nt,nlevel], flux_plus_v [ng,nt,nlevel,nwno],
    flux_minus_v [ng,nt,nlevel,nwno], flux_net_ir_layer [nlevel], flux_net_ir [nlevel],
    flux_plus_ir [nlevel,nwno], flux_minus_ir [nlevel,nwno]

All correlated-k gauss points go to the device in one `pb_climate_get_fluxes` call
(csrc/climate.cu); the opacity arrays may be numpy ``[nlayer, nwno, ngauss]`` arrays or the
`DeviceArray`s that ``picaso_b200.compute_opacity(..., device_outputs=True)`` returns.

The reference function is numba-nopython and is bound inside `t_start` at compile time
(climate caveat), so it is replaced by driving the solver loop from Python
(`picaso_b200.patch(picaso.climate)` rebinds the module-level name for callers that are not jitted).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_DEVICE, PB_HOST, ClimateArgs, addr

__all__ = ["get_fluxes", "get_fluxes_jacobian", "BoundFluxes"]

_WED = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "W0_no_raman")
_NOED = ("DTAU", "TAU", "W0", "COSB")


def _is_dev(a):
    return hasattr(a, "ptr") and hasattr(a, "ctx")


def _as3(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        a = a[:, :, None]
    return np.ascontiguousarray(a)


def _build_args(ctx, Atmosphere, wed, noed, ScatteringPhase, Disco, Opagrid, F0PI, reflected, thermal, full_arrays=True,
                tlevels=None):
    """fill a pb_climate_args; returns (args, memspace, outputs dict, keep-alive list).  tlevels [P, nlevel] selects
    the Jacobian-batch mode (thermal half for P temperature profiles over one set of opacities)."""
    nlevel = int(Atmosphere.nlevel)
    nlayer = nlevel - 1
    nwno = int(Opagrid.nwno)
    ngauss = int(Opagrid.ngauss)
    ng, nt = int(Disco.ng), int(Disco.nt)
    arrays = {k: getattr(wed, k) for k in _WED}
    arrays.update({k + "_OG": getattr(noed, k) for k in _NOED})
    dev = [_is_dev(v) for v in arrays.values()]
    keep = []
    if any(dev):
        if not all(dev):
            raise _lib.PicasoB200Error("get_fluxes: opacity arrays must be all numpy or all DeviceArray")
        memspace = PB_DEVICE
        ptrs = {k: int(v.ptr) for k, v in arrays.items()}
        shapes = {k: tuple(v.shape) for k, v in arrays.items()}
    else:
        memspace = PB_HOST
        host = {k: _as3(v) for k, v in arrays.items()}
        keep.extend(host.values())
        ptrs = {k: addr(v) for k, v in host.items()}
        shapes = {k: v.shape for k, v in host.items()}
    for k, shp in shapes.items():
        rows = nlevel if k.startswith("TAU") else nlayer
        shp3 = tuple(shp) + (1,) * (3 - len(shp))
        if shp3 != (rows, nwno, ngauss):
            raise _lib.PicasoB200Error("get_fluxes: %s has shape %s, expected %s" % (k, shp, (rows, nwno, ngauss)))

    def vec(x, n, fill=None):
        if x is None:
            return None
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (n,)))
        keep.append(v)
        return v

    a = ClimateArgs()
    a.nlayer, a.nwno, a.ngauss, a.numg, a.numt = nlayer, nwno, ngauss, ng, nt
    a.reflected, a.thermal = int(bool(reflected)), int(bool(thermal))
    for k, p in ptrs.items():
        setattr(a, k, p)
    a.gauss_wts = addr(vec(Opagrid.gauss_wts, ngauss))
    a.wno = addr(vec(Opagrid.wno, nwno))
    a.dwno = addr(vec(Opagrid.delta_wno, nwno))
    a.surf_reflect = addr(vec(ScatteringPhase.surf_reflect, nwno))
    a.F0PI = addr(vec(F0PI, nwno))
    a.tlevel = addr(vec(Atmosphere.t_level, nlevel))
    a.plevel = addr(vec(Atmosphere.p_level, nlevel))
    ubar1 = np.ascontiguousarray(np.asarray(Disco.ubar1, dtype=np.float64).reshape(ng * nt))
    keep.append(ubar1)
    a.ubar1 = addr(ubar1)
    a.gweight = addr(vec(Disco.gweight, ng))
    a.tweight = addr(vec(Disco.tweight, nt))
    a.cos_theta = float(Disco.cos_theta)
    a.single_phase, a.multi_phase = int(ScatteringPhase.single_phase), int(ScatteringPhase.multi_phase)
    a.frac_a, a.frac_b, a.frac_c = (float(ScatteringPhase.frac_a), float(ScatteringPhase.frac_b),
                                    float(ScatteringPhase.frac_c))
    a.constant_back, a.constant_forward = float(ScatteringPhase.constant_back), float(ScatteringPhase.constant_forward)
    # all outputs arrive in one host block with one device-to-host copy (pb_climate_args.packed); the arrays
    # handed back are non-overlapping views of it
    nvw = nlevel * nwno
    blk = np.empty(4 * nlevel + (4 * nvw if full_arrays else 0))
    a.packed, a.packed_full = addr(blk), int(bool(full_arrays))
    big = (lambda i: blk[4 * nlevel + i * nvw:4 * nlevel + (i + 1) * nvw].reshape(nlevel, nwno)) if full_arrays \
        else (lambda i: None)
    vecn = lambda i: blk[i * nlevel:(i + 1) * nlevel]
    out = dict(flux_net_v_layer=vecn(0), flux_net_v=vecn(1), flux_plus_v=big(0), flux_minus_v=big(1),
               flux_net_ir_layer=vecn(2), flux_net_ir=vecn(3), flux_plus_ir=big(2), flux_minus_ir=big(3))
    if tlevels is not None:
        tl = np.ascontiguousarray(tlevels, dtype=np.float64)
        if tl.ndim != 2 or tl.shape[1] != nlevel:
            raise _lib.PicasoB200Error("tlevels must be [nprofiles, nlevel]")
        jac = np.empty((tl.shape[0], 2, nlevel))
        keep += [tl, jac]
        a.nprofiles, a.tlevels, a.jac_out = tl.shape[0], addr(tl), addr(jac)
        out = dict(tlevels=tl, jac=jac)
    keep.append(blk)
    return a, memspace, out, keep


def _one_column(ctx, Atmosphere, wed, noed, ScatteringPhase, Disco, Opagrid, F0PI, reflected, thermal, full_arrays=True):
    a, memspace, out, keep = _build_args(ctx, Atmosphere, wed, noed, ScatteringPhase, Disco, Opagrid, F0PI, reflected,
                                         thermal, full_arrays)
    ctx.check(ctx.lib.pb_climate_get_fluxes(ctx.h, ctypes.byref(a), memspace))
    del keep
    return out


def get_fluxes_jacobian(Atmosphere, OpacityWEd, OpacityNoEd, ScatteringPhase, Disco, Opagrid, tlevels, *, ctx=None):
    """The thermal half of get_fluxes for every row of tlevels [nprofiles, nlevel] in ONE device call - what the
    Jacobian loop of t_start asks for one perturbed level at a time (climate.py:1108-1180: temp[jm] += deltaT,
    get_fluxes(thermal only), restore).  The opacities (which do not change inside that loop) are staged once;
    only the Planck terms differ between profiles.  Returns (flux_net_ir_layer, flux_net_ir), each
    [nprofiles, nlevel]; row p equals get_fluxes(..., reflected=False, thermal=True)[4:6] with t_level = tlevels[p]."""
    ctx = ctx or _lib.default_context()
    a, memspace, out, keep = _build_args(ctx, Atmosphere, OpacityWEd, OpacityNoEd, ScatteringPhase, Disco, Opagrid, None,
                                         False, True, False, tlevels=tlevels)
    ctx.check(ctx.lib.pb_climate_get_fluxes(ctx.h, ctypes.byref(a), memspace))
    jac = out["jac"]
    del keep
    return jac[:, 0, :].copy(), jac[:, 1, :].copy()


class BoundFluxes:
    """A get_fluxes call bound once and re-run from numba nopython code (the reference's t_start / get_fluxes are jitted,
    so rebinding Python names cannot reach them - SURVEY.md section 8b).

        bf = BoundFluxes(Atmosphere, OpacityWEd, ..., reflected=False, thermal=True)      # or tlevels=[P, nlevel]
        run, addr, h = bf.run, bf.ctx_address, bf.handle        # ctypes function + two integers
        @numba.njit
        def solver_step(t_level, net_ir):                       # t_level, net_ir: bf.t_level, bf.flux_net_ir
            t_level[3] += 1.0                                    # the solver writes temperatures in place ...
            rc = run(addr, h)                                    # ... and calls the device
            return net_ir[0]

    `run` takes two integers, so numba's ctypes support can call it; inputs are read from / outputs written to the
    numpy arrays exposed as attributes (t_level or tlevels, flux_net_v_layer, flux_net_v, flux_net_ir_layer,
    flux_net_ir, jac)."""

    def __init__(self, Atmosphere, OpacityWEd, OpacityNoEd, ScatteringPhase, Disco, Opagrid, F0PI=None, reflected=False,
                 thermal=True, tlevels=None, *, ctx=None):
        self.ctx = ctx or _lib.default_context()
        nlevel = int(Atmosphere.nlevel)
        self.t_level = np.ascontiguousarray(Atmosphere.t_level, dtype=np.float64).copy()
        atm = type("Atm", (), dict(nlevel=nlevel, t_level=self.t_level, p_level=Atmosphere.p_level))()
        a, memspace, out, keep = _build_args(self.ctx, atm, OpacityWEd, OpacityNoEd, ScatteringPhase, Disco, Opagrid, F0PI,
                                             reflected, thermal, False, tlevels=tlevels)
        if tlevels is None:
            # _build_args copied t_level into a fresh vector: point the struct at OUR array so in-place writes count
            a.tlevel = addr(self.t_level)
            self.flux_net_v_layer, self.flux_net_v = out["flux_net_v_layer"], out["flux_net_v"]
            self.flux_net_ir_layer, self.flux_net_ir = out["flux_net_ir_layer"], out["flux_net_ir"]
        else:
            self.tlevels, self.jac = out["tlevels"], out["jac"]
        self._keep, self._args = keep, a
        h = ctypes.c_int(-1)
        self.ctx.check(self.ctx.lib.pb_climate_bind(self.ctx.h, ctypes.byref(a), memspace, ctypes.byref(h)))
        self.handle = int(h.value)
        self.ctx_address = int(self.ctx.h.value)
        self.run = self.ctx.lib.pb_climate_run_bound

    def __call__(self):
        self.ctx.check(self.run(self.ctx_address, self.handle))

    def close(self):
        if self.handle >= 0:
            self.ctx.lib.pb_climate_unbind(self.ctx.h, self.handle)
            self.handle = -1


def get_fluxes(Atmosphere, OpacityWEd, OpacityNoEd, ScatteringPhase, Disco, Opagrid, F0PI, reflected, thermal,
               do_holes=False, fhole=0.0, hole_OpacityWEd=None, hole_OpacityNoEd=None, *, ctx=None, full_arrays=True):
    """picaso.climate.get_fluxes (climate.py:1686-1952) on the device.  Same arguments, same 8-tuple.
    full_arrays=False (extension): the four [.., nlevel, nwno] arrays are not copied back (None in the tuple) -
    the solver's Newton / Jacobian iterations only consume the four net-flux vectors."""
    ctx = ctx or _lib.default_context()
    out = _one_column(ctx, Atmosphere, OpacityWEd, OpacityNoEd, ScatteringPhase, Disco, Opagrid, F0PI,
                      reflected, thermal, full_arrays)
    if do_holes:
        # climate.py:1822-1837, :1894-1909: the level arrays of the clear column are mixed in with weight
        # fhole before every (linear) reduction, i.e. the outputs mix with the same weights
        clr = _one_column(ctx, Atmosphere, hole_OpacityWEd, hole_OpacityNoEd, ScatteringPhase, Disco, Opagrid,
                          F0PI, reflected, thermal, full_arrays)
        for k in out:
            if out[k] is not None:
                out[k] = (1.0 - fhole) * out[k] + fhole * clr[k]
    ng, nt = int(Disco.ng), int(Disco.nt)
    nlevel, nwno = int(Atmosphere.nlevel), int(Opagrid.nwno)
    # the visible arrays are [ng, nt, ...] in the reference; the single mu = 0.5 stream is broadcast (climate.py:1757-1761)
    def bc(x, shape):
        if x is None:
            return None
        if ng * nt == 1:
            return x.reshape(shape)  # no copy
        return np.ascontiguousarray(np.broadcast_to(x, shape))
    return (bc(out["flux_net_v_layer"], (ng, nt, nlevel)), bc(out["flux_net_v"], (ng, nt, nlevel)),
            bc(out["flux_plus_v"], (ng, nt, nlevel, nwno)), bc(out["flux_minus_v"], (ng, nt, nlevel, nwno)),
            out["flux_net_ir_layer"], out["flux_net_ir"], out["flux_plus_ir"], out["flux_minus_ir"])
