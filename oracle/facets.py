"""Oracle wrappers of the 3-D (per-facet) flux functions - TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import _c, _p, _wvec, lib


def _facet_major(a, ng, nt):
    """[rows, nwno, ng, nt] (reference layout, fluxes.py:470-481) -> [ng*nt, rows, nwno]"""
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(np.moveaxis(a.reshape(a.shape[0], a.shape[1], ng * nt), 2, 0))


def get_reflected_3d(nlevel, wno, nwno, numg, numt, dtau_3d, tau_3d, w0_3d, cosb_3d, gcos2_3d, ftau_cld_3d,
                     ftau_ray_3d, dtau_og_3d, tau_og_3d, w0_og_3d, cosb_og_3d, surf_reflect, ubar0, ubar1,
                     cos_theta, F0PI, single_phase, multi_phase, frac_a, frac_b, frac_c, constant_back,
                     constant_forward, nthreads=1):
    """Oracle of fluxes.py:355-660; returns xint_at_top [numg, numt, nwno]."""
    arrs = [_facet_major(a, numg, numt) for a in (dtau_3d, tau_3d, w0_3d, cosb_3d, gcos2_3d, ftau_cld_3d,
                                                  ftau_ray_3d, dtau_og_3d, tau_og_3d, w0_og_3d, cosb_og_3d)]
    sr, f0 = _wvec(surf_reflect, nwno), _wvec(F0PI, nwno)
    u0, u1 = _c(ubar0), _c(ubar1)
    xint = np.zeros((numg, numt, nwno))
    f = lib().orc_get_reflected_3d
    f.restype = None
    ci, cd = ctypes.c_int, ctypes.c_double
    f(ci(nlevel), ci(nwno), ci(numg), ci(numt), *[_p(a) for a in arrs], _p(sr), _p(u0), _p(u1), cd(cos_theta),
      _p(f0), ci(single_phase), ci(multi_phase), cd(frac_a), cd(frac_b), cd(frac_c), cd(constant_back),
      cd(constant_forward), _p(xint), ci(nthreads))
    return xint


def get_thermal_3d(nlevel, wno, nwno, numg, numt, tlevel_3d, dtau_3d, w0_3d, cosb_3d, plevel_3d, ubar1,
                   surf_reflect, hard_surface, nthreads=1):
    """Oracle of fluxes.py:2148-2352; returns int_at_top [numg, numt, nwno]."""
    arrs = [_facet_major(a, numg, numt) for a in (dtau_3d, w0_3d, cosb_3d)]
    G = numg * numt
    tl = np.ascontiguousarray(np.asarray(tlevel_3d, dtype=np.float64).reshape(nlevel, G).T)
    pl = np.ascontiguousarray(np.asarray(plevel_3d, dtype=np.float64).reshape(nlevel, G).T)
    wn, u1, sr = _c(wno), _c(ubar1), _wvec(surf_reflect, nwno)
    out = np.zeros((numg, numt, nwno))
    f = lib().orc_get_thermal_3d
    f.restype = None
    ci = ctypes.c_int
    f(ci(nlevel), _p(wn), ci(nwno), ci(numg), ci(numt), _p(tl), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(pl),
      _p(u1), _p(sr), ci(int(hard_surface)), _p(out), ci(nthreads))
    return out
