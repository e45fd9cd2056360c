"""get_reflected_3d / get_thermal_3d mirrors (picaso/fluxes.py:355-660, :2148-2352).

The reference keeps per-facet opacities as [nlayer|nlevel, nwno, ng, nt] with the angle axes
FASTEST; the kernels want wavelength fastest, so the facets become the batch axis of the 1-D
kernels ([ng*nt][rows][nwno], one geometry value per batch entry, `variant = 1`)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_HOST, ReflectedArgs, ThermalArgs, addr


def _facet_major(a, rows, nwno, G):
    a = np.asarray(a, dtype=np.float64)
    if a.shape[:2] != (rows, nwno) or a.size != rows * nwno * G:
        raise ValueError("expected array of shape (%d, %d, ng, nt), got %s" % (rows, nwno, a.shape))
    return np.ascontiguousarray(np.moveaxis(a.reshape(rows, nwno, G), 2, 0))


def _wvec(x, n):
    a = np.asarray(x, dtype=np.float64)
    return np.full(n, float(a)) if a.ndim == 0 else np.ascontiguousarray(a)


def get_reflected_3d(nlevel, wno, nwno, numg, numt, dtau_3d, tau_3d, w0_3d, cosb_3d, gcos2_3d, ftau_cld_3d,
                     ftau_ray_3d, dtau_og_3d, tau_og_3d, w0_og_3d, cosb_og_3d, surf_reflect, ubar0, ubar1,
                     cos_theta, F0PI, single_phase, multi_phase, frac_a, frac_b, frac_c, constant_back,
                     constant_forward, *, ctx=None):
    """CUDA replacement of fluxes.get_reflected_3d: returns xint_at_top [numg, numt, nwno]."""
    ctx = ctx or _lib.default_context()
    nlayer, G = nlevel - 1, numg * numt
    lay = [_facet_major(x, nlayer, nwno, G) for x in (dtau_3d, w0_3d, cosb_3d, gcos2_3d, ftau_cld_3d, ftau_ray_3d,
                                                     dtau_og_3d, w0_og_3d, cosb_og_3d)]
    lev = [_facet_major(x, nlevel, nwno, G) for x in (tau_3d, tau_og_3d)]
    sr, f0 = _wvec(surf_reflect, nwno), _wvec(F0PI, nwno)
    u0 = np.ascontiguousarray(ubar0, dtype=np.float64).reshape(-1)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    xint = np.zeros((numg, numt, nwno))
    a = ReflectedArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlayer, nwno, 1, 1, G, nwno
    (a.dtau, a.w0, a.cosb, a.gcos2, a.ftau_cld, a.ftau_ray, a.dtau_og, a.w0_og, a.cosb_og) = [addr(x) for x in lay]
    a.tau, a.tau_og = addr(lev[0]), addr(lev[1])
    a.surf_reflect, a.F0PI, a.b_top = addr(sr), addr(f0), None
    a.ubar0, a.ubar1 = addr(u0), addr(u1)
    a.cos_theta = float(cos_theta)
    a.single_phase, a.multi_phase, a.toon_coefficients = int(single_phase), int(multi_phase), 0
    a.frac_a, a.frac_b, a.frac_c = float(frac_a), float(frac_b), float(frac_c)
    a.constant_back, a.constant_forward = float(constant_back), float(constant_forward)
    a.get_toa_intensity, a.get_lvl_flux, a.variant = 1, 0, 1
    a.xint_at_top = addr(xint)
    if nwno > 0:
        ctx.check(ctx.lib.pb_reflected_toon_1d(ctx.h, ctypes.byref(a), PB_HOST))
    return xint


def get_thermal_3d(nlevel, wno, nwno, numg, numt, tlevel_3d, dtau_3d, w0_3d, cosb_3d, plevel_3d, ubar1,
                   surf_reflect, hard_surface, *, ctx=None):
    """CUDA replacement of fluxes.get_thermal_3d: returns int_at_top [numg, numt, nwno]."""
    ctx = ctx or _lib.default_context()
    nlayer, G = nlevel - 1, numg * numt
    lay = [_facet_major(x, nlayer, nwno, G) for x in (dtau_3d, w0_3d, cosb_3d)]
    tl = np.ascontiguousarray(np.asarray(tlevel_3d, dtype=np.float64).reshape(nlevel, G).T)
    pl = np.ascontiguousarray(np.asarray(plevel_3d, dtype=np.float64).reshape(nlevel, G).T)
    wn = np.ascontiguousarray(wno, dtype=np.float64)
    sr = _wvec(surf_reflect, nwno)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    out = np.zeros((numg, numt, nwno))
    a = ThermalArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlayer, nwno, 1, 1, G, nwno
    a.dtau, a.w0, a.cosb = [addr(x) for x in lay]
    a.wno, a.dwno, a.surf_reflect = addr(wn), None, addr(sr)
    a.tlevel, a.plevel, a.ubar1 = addr(tl), addr(pl), addr(u1)
    a.hard_surface, a.calc_type, a.variant = int(hard_surface), 0, 1
    a.flux_at_top = addr(out)
    if nwno > 0:
        ctx.check(ctx.lib.pb_thermal_toon_1d(ctx.h, ctypes.byref(a), PB_HOST))
    return out
