"""get_thermal_SH mirror (picaso/fluxes.py:2979-3186) - kept in its own module; re-exported by
picaso_b200.fluxes / picaso_b200."""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_HOST, ThermalShArgs, addr

_ZERO = np.zeros(())


def get_thermal_SH(nlevel, wno, nwno, numg, numt, tlevel, dtau, tau, w0, cosb, dtau_og, tau_og, w0_og,
                   w0_no_raman, cosb_og, plevel, ubar1, surf_reflect, stream, hard_surface, flx=0, *,
                   ctx=None, gweight=None, tweight=None, return_thermal=False):
    """CUDA replacement of fluxes.get_thermal_SH (stream 2 or 4).

    Returns ``(xint_at_top[numg,numt,nwno], flux)`` with ``flux`` the reference's all-zero
    ``[numg,numt,stream*nlevel,nwno]`` array (flx=0; read-only zero view).  ``tau``, ``dtau_og``,
    ``tau_og``, ``w0_og`` and ``w0_no_raman`` are accepted and ignored exactly as the reference
    ignores them; ``cosb`` is only compared with ``cosb_og`` (fluxes.py:3044-3047).  flx=1 raises
    in the reference (``flux_temp`` used before assignment, fluxes.py:3102) and is refused here.
    """
    from .fluxes import _layer_set, _wvec
    if flx != 0:
        raise NotImplementedError("get_thermal_SH(flx=1) is broken in the reference and not implemented")
    ctx = ctx or _lib.default_context()
    nlayer = nlevel - 1
    alias = cosb is cosb_og
    lay, ld = _layer_set([dtau, w0, cosb, cosb_og], nlayer, nwno)
    if ld != nwno:  # the equality test runs on dense device copies
        lay = [np.ascontiguousarray(x) for x in lay]
        ld = nwno
    if alias:
        lay[2] = lay[3]
    wn = np.ascontiguousarray(wno, dtype=np.float64)
    sr = _wvec(surf_reflect, nwno)
    tl = np.ascontiguousarray(tlevel, dtype=np.float64)
    pl = np.ascontiguousarray(plevel, dtype=np.float64)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    xint = np.zeros((numg, numt, nwno))
    th = np.zeros(nwno) if return_thermal else None
    gw = tw = None
    if return_thermal:
        gw = np.ascontiguousarray(gweight, dtype=np.float64)
        tw = np.ascontiguousarray(tweight, dtype=np.float64)
    a = ThermalShArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlayer, nwno, numg, numt, 1, ld
    a.dtau, a.w0, a.cosb, a.cosb_og = [addr(x) for x in lay]
    a.wno, a.surf_reflect = addr(wn), addr(sr)
    a.tlevel, a.plevel, a.ubar1, a.gweight, a.tweight = addr(tl), addr(pl), addr(u1), addr(gw), addr(tw)
    a.stream, a.hard_surface, a.flx = int(stream), int(hard_surface), 0
    a.xint_at_top, a.thermal = addr(xint), addr(th)
    if nwno > 0:
        ctx.check(ctx.lib.pb_thermal_sh(ctx.h, ctypes.byref(a), PB_HOST))
    flux = np.broadcast_to(_ZERO, (numg, numt, stream * nlevel, nwno))
    if return_thermal:
        return xint, flux, th
    return xint, flux
