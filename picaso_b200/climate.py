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

__all__ = ["get_fluxes"]

_WED = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "W0_no_raman")
_NOED = ("DTAU", "TAU", "W0", "COSB")


def _is_dev(a):
    return hasattr(a, "ptr") and hasattr(a, "ctx")


def _as3(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        a = a[:, :, None]
    return np.ascontiguousarray(a)


def _one_column(ctx, Atmosphere, wed, noed, ScatteringPhase, Disco, Opagrid, F0PI, reflected, thermal, full_arrays=True):
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
    ctx.check(ctx.lib.pb_climate_get_fluxes(ctx.h, ctypes.byref(a), memspace))
    del keep
    return out


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
