"""Host-side mirror of picaso/fluxes.py for the hot path: same names, argument order,
defaults and return values as the reference's numba functions, executing on the GPU
through the C ABI (include/picaso_b200.h).  Extra keyword-only arguments (`ctx`, fused
outputs) are additions; positional use is identical to the reference.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_DEVICE, PB_HOST, ReflectedArgs, ShArgs, ThermalArgs, TransitArgs, addr

__all__ = ["get_reflected_1d", "get_reflected_SH", "get_thermal_1d", "get_transit_1d"]


_ZERO = np.zeros(())


def _f64(a):
    """float64 view with unit wavelength stride; copies only when it has to (e.g. the
    X[:, :, ig] slices picaso() passes when ngauss > 1, justdoit.py:275-283)."""
    a = np.asarray(a)
    if a.dtype != np.float64 or a.ndim == 0 or a.strides[-1] != 8 or \
            (a.ndim == 2 and (a.strides[0] % 8 or a.strides[0] < a.shape[1] * 8)):
        a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _layer_set(arrs, nrows, nwno):
    """normalise a group of [rows, nwno] arrays to one common leading dimension."""
    out = [_f64(a) for a in arrs]
    for a in out:
        if a.shape != (nrows, nwno):
            raise ValueError("expected array of shape (%d, %d), got %s" % (nrows, nwno, a.shape))
    lds = {a.strides[0] // 8 if nrows > 1 else nwno for a in out}
    if len(lds) > 1:
        out = [np.ascontiguousarray(a) for a in out]
        lds = {nwno}
    return out, lds.pop()


def _is_dev(a):
    return hasattr(a, "ptr") and hasattr(a, "ctx")


def _resolve(ctx, arrs, nrows, nwno, wave_vecs=()):
    """Normalise a group of [rows, nwno] inputs.  Host (numpy) inputs -> (addresses, ld, PB_HOST,
    wave-vector addresses, keepalive).  DeviceArray inputs (picaso_b200.optics.compute_opacity with
    device_outputs=True) -> PB_DEVICE pointers; the small per-wavelength vectors are uploaded."""
    if any(_is_dev(a) for a in arrs):
        if not all(_is_dev(a) for a in arrs):
            raise TypeError("either all or none of the layer/level arrays may be DeviceArrays")
        for a in arrs:
            if tuple(a.shape[:2]) not in ((nrows, nwno), (nrows + 1, nwno)):
                raise ValueError("DeviceArray of shape %s does not match (%d|%d, %d)" % (a.shape, nrows, nrows + 1, nwno))
        from .optics import DeviceArray
        vecs = [None if v is None else DeviceArray.from_numpy(ctx, v) for v in wave_vecs]
        return [a.ptr for a in arrs], nwno, PB_DEVICE, [None if v is None else v.ptr for v in vecs], vecs
    return None


def _wvec(x, n):
    """scalar or [n] -> contiguous float64 [n] (surf_reflect/b_top arrive either way)."""
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0:
        return np.full(n, float(a))
    if a.shape != (n,):
        raise ValueError("expected scalar or vector of length %d, got shape %s" % (n, a.shape))
    return np.ascontiguousarray(a)


def get_reflected_1d(nlevel, wno, nwno, numg, numt, dtau, tau, w0, cosb, gcos2, ftau_cld, ftau_ray,
                     dtau_og, tau_og, w0_og, cosb_og, surf_reflect, ubar0, ubar1, cos_theta, F0PI,
                     single_phase, multi_phase, frac_a, frac_b, frac_c, constant_back,
                     constant_forward, get_toa_intensity=1, get_lvl_flux=0, toon_coefficients=0,
                     b_top=0, *, ctx=None, gweight=None, tweight=None, return_albedo=False):
    """CUDA replacement of fluxes.get_reflected_1d (picaso/fluxes.py:1010-1413).

    Returns ``(xint_at_top[numg,numt,nwno], (flux_minus, flux_plus, flux_minus_midpt,
    flux_plus_midpt))`` each level array [numg,numt,nlevel,nwno] (zeros unless
    get_lvl_flux) - exactly as the reference.  With ``return_albedo=True`` (needs gweight,
    tweight) the fused compress_disco result is appended as a third element.
    """
    ctx = ctx or _lib.default_context()
    nlayer = nlevel - 1
    if any(_is_dev(x) for x in (dtau, tau, w0, cosb, gcos2, ftau_cld, ftau_ray, dtau_og, tau_og, w0_og, cosb_og)):
        return _reflected_device(ctx, nlevel, nwno, numg, numt, [dtau, w0, cosb, gcos2, ftau_cld, ftau_ray,
                                 dtau_og, w0_og, cosb_og], [tau, tau_og], surf_reflect, ubar0, ubar1, cos_theta,
                                 F0PI, single_phase, multi_phase, frac_a, frac_b, frac_c, constant_back,
                                 constant_forward, get_toa_intensity, get_lvl_flux, toon_coefficients, b_top,
                                 gweight, tweight, return_albedo)
    same = dict(dtau_og=dtau_og is dtau, w0_og=w0_og is w0, cosb_og=cosb_og is cosb,
                tau_og=tau_og is tau)
    lay, ld = _layer_set([dtau, w0, cosb, gcos2, ftau_cld, ftau_ray, dtau_og, w0_og, cosb_og],
                         nlayer, nwno)
    lev, ldv = _layer_set([tau, tau_og], nlevel, nwno)
    if ldv != ld:
        lay = [np.ascontiguousarray(a) for a in lay]
        lev = [np.ascontiguousarray(a) for a in lev]
        ld = nwno
    if same["dtau_og"]: lay[6] = lay[0]
    if same["w0_og"]: lay[7] = lay[1]
    if same["cosb_og"]: lay[8] = lay[2]
    if same["tau_og"]: lev[1] = lev[0]
    sr, f0, bt = _wvec(surf_reflect, nwno), _wvec(F0PI, nwno), _wvec(b_top, nwno)
    u0 = np.ascontiguousarray(ubar0, dtype=np.float64).reshape(-1)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    if u0.size != numg * numt or u1.size != numg * numt:
        raise ValueError("ubar0/ubar1 must have numg*numt entries")
    xint = np.zeros((numg, numt, nwno))
    lv = [np.zeros((numg, numt, nlevel, nwno)) for _ in range(4)] if get_lvl_flux else None
    alb = np.zeros(nwno) if return_albedo else None
    gw = tw = None
    if return_albedo:
        gw = np.ascontiguousarray(gweight, dtype=np.float64)
        tw = np.ascontiguousarray(tweight, dtype=np.float64)
    a = ReflectedArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlayer, nwno, numg, numt, 1, ld
    (a.dtau, a.w0, a.cosb, a.gcos2, a.ftau_cld, a.ftau_ray, a.dtau_og, a.w0_og,
     a.cosb_og) = [addr(x) for x in lay]
    a.tau, a.tau_og = addr(lev[0]), addr(lev[1])
    a.surf_reflect, a.F0PI, a.b_top = addr(sr), addr(f0), addr(bt)
    a.ubar0, a.ubar1, a.gweight, a.tweight = addr(u0), addr(u1), addr(gw), addr(tw)
    a.cos_theta = float(cos_theta)
    a.single_phase, a.multi_phase, a.toon_coefficients = int(single_phase), int(multi_phase), int(toon_coefficients)
    a.frac_a, a.frac_b, a.frac_c = float(frac_a), float(frac_b), float(frac_c)
    a.constant_back, a.constant_forward = float(constant_back), float(constant_forward)
    a.get_toa_intensity, a.get_lvl_flux = int(get_toa_intensity), int(get_lvl_flux)
    a.xint_at_top, a.albedo = addr(xint), addr(alb)
    if lv is not None:
        a.flux_minus, a.flux_plus, a.flux_minus_mdpt, a.flux_plus_mdpt = [addr(x) for x in lv]
    if nwno > 0:
        ctx.check(ctx.lib.pb_reflected_toon_1d(ctx.h, ctypes.byref(a), PB_HOST))
    if lv is None:
        # the reference always returns four zero arrays (fluxes.py:1113-1121); hand back
        # read-only zero-stride views of the same shape instead of touching ~100 MB per call
        z = np.broadcast_to(_ZERO, (numg, numt, nlevel, nwno))
        lv = [z, z, z, z]
    if return_albedo:
        return xint, tuple(lv), alb
    return xint, tuple(lv)


def _zero_or_vec(x, n):
    """None for an all-zero scalar (the C side treats NULL surf_reflect / b_top as 0), else [n]."""
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0 and float(a) == 0.0:
        return None
    return _wvec(x, n)


def _reflected_device(ctx, nlevel, nwno, numg, numt, lay, lev, surf_reflect, ubar0, ubar1, cos_theta, F0PI,
                      single_phase, multi_phase, frac_a, frac_b, frac_c, constant_back, constant_forward,
                      get_toa_intensity, get_lvl_flux, toon_coefficients, b_top, gweight, tweight,
                      return_albedo):
    """get_reflected_1d on DeviceArray inputs: kernels read HBM directly, only [G, W] comes back."""
    from .optics import DeviceArray
    nlayer, G = nlevel - 1, numg * numt
    f0 = np.asarray(F0PI, dtype=np.float64)
    vecs = [_zero_or_vec(surf_reflect, nwno), None if (f0.ndim == 0 and float(f0) == 1.0) else _wvec(F0PI, nwno),
            _zero_or_vec(b_top, nwno)]
    ptrs, ld, memspace, vptr, keep = _resolve(ctx, lay + lev, nlayer, nwno, vecs)
    u0 = np.ascontiguousarray(ubar0, dtype=np.float64).reshape(-1)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    gw = tw = None
    if return_albedo:
        gw = np.ascontiguousarray(gweight, dtype=np.float64)
        tw = np.ascontiguousarray(tweight, dtype=np.float64)
    # TOA outputs share one context-owned workspace block ([G + 1][nwno]: xint rows, then the albedo) so that a
    # call costs no cudaMalloc and one device-to-host copy
    ws = ctx.workspace("refl_toa_out", (G + 1) * nwno * 8)
    d_x = DeviceArray(ctx, (numg, numt, nwno), ptr=ws)
    d_a = DeviceArray(ctx, (nwno,), ptr=ws + G * nwno * 8) if return_albedo else None
    d_lv = [DeviceArray(ctx, (numg, numt, nlevel, nwno)) for _ in range(4)] if get_lvl_flux else None
    a = ReflectedArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlayer, nwno, numg, numt, 1, ld
    (a.dtau, a.w0, a.cosb, a.gcos2, a.ftau_cld, a.ftau_ray, a.dtau_og, a.w0_og, a.cosb_og, a.tau, a.tau_og) = ptrs
    a.surf_reflect, a.F0PI, a.b_top = vptr
    a.ubar0, a.ubar1, a.gweight, a.tweight = addr(u0), addr(u1), addr(gw), addr(tw)
    a.cos_theta = float(cos_theta)
    a.single_phase, a.multi_phase, a.toon_coefficients = int(single_phase), int(multi_phase), int(toon_coefficients)
    a.frac_a, a.frac_b, a.frac_c = float(frac_a), float(frac_b), float(frac_c)
    a.constant_back, a.constant_forward = float(constant_back), float(constant_forward)
    a.get_toa_intensity, a.get_lvl_flux = int(get_toa_intensity), int(get_lvl_flux)
    a.xint_at_top, a.albedo = d_x.ptr, (d_a.ptr if d_a else None)
    if d_lv:
        a.flux_minus, a.flux_plus, a.flux_minus_mdpt, a.flux_plus_mdpt = [x.ptr for x in d_lv]
    ctx.check(ctx.lib.pb_reflected_toon_1d(ctx.h, ctypes.byref(a), memspace))
    both = ctx.from_device(ws, ((G + 1) if return_albedo else G, nwno))
    xint = both[:G].reshape(numg, numt, nwno)
    if d_lv:
        lv = [x.numpy() for x in d_lv]
    else:
        z = np.broadcast_to(_ZERO, (numg, numt, nlevel, nwno))
        lv = [z, z, z, z]
    if return_albedo:
        return xint, tuple(lv), both[G]
    return xint, tuple(lv)


def get_reflected_SH(nlevel, nwno, numg, numt, dtau, tau, w0, cosb, ftau_cld, ftau_ray, f_deltaM,
                     dtau_og, tau_og, w0_og, cosb_og, surf_reflect, ubar0, ubar1, cos_theta, F0PI,
                     w_single_form, w_multi_form, psingle_form, w_single_rayleigh, w_multi_rayleigh,
                     psingle_rayleigh, frac_a, frac_b, frac_c, constant_back, constant_forward, stream,
                     b_top=0, flx=0, single_form=0, *, ctx=None, gweight=None, tweight=None,
                     return_albedo=False, inplace_f_deltaM=True):
    """CUDA replacement of fluxes.get_reflected_SH (picaso/fluxes.py:2675-2976), stream 2 or 4.

    Returns ``(xint_at_top[numg,numt,nwno], flux)`` where ``flux`` is the reference's
    ``[numg,numt,stream*nlevel,nwno]`` array: all zero for flx=0 (a read-only zero view), the layer fluxes
    ``calculate_flux(F, G, X)`` for flx=1 (fluxes.py:2889-2890; ``calculate_fluxes='on'``, justdoit.py:4638).
    ``inplace_f_deltaM=True`` reproduces the reference's side effect on a writeable float64
    ``f_deltaM`` argument (it is scaled once per angle when a TTHG form is active,
    fluxes.py:2823-2824).  ``cosb`` is accepted and ignored, as in the reference.
    """
    if flx not in (0, 1):
        raise ValueError("get_reflected_SH: flx must be 0 or 1")
    ctx = ctx or _lib.default_context()
    nlayer = nlevel - 1
    same = (dtau_og is dtau, w0_og is w0, tau_og is tau)
    lay, ld = _layer_set([dtau, w0, ftau_cld, ftau_ray, f_deltaM, dtau_og, w0_og, cosb_og], nlayer, nwno)
    lev, ldv = _layer_set([tau, tau_og], nlevel, nwno)
    if ldv != ld:
        lay = [np.ascontiguousarray(a) for a in lay]
        lev = [np.ascontiguousarray(a) for a in lev]
        ld = nwno
    if same[0]: lay[5] = lay[0]
    if same[1]: lay[6] = lay[1]
    if same[2]: lev[1] = lev[0]
    sr, f0, bt = _wvec(surf_reflect, nwno), _wvec(F0PI, nwno), _wvec(b_top, nwno)
    u0 = np.ascontiguousarray(ubar0, dtype=np.float64).reshape(-1)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    xint = np.zeros((numg, numt, nwno))
    alb = np.zeros(nwno) if return_albedo else None
    drift = inplace_f_deltaM and (w_single_form == 0 or w_multi_form == 0) and \
        isinstance(f_deltaM, np.ndarray) and f_deltaM.flags.writeable
    fd_out = np.zeros((nlayer, nwno)) if drift else None
    gw = tw = None
    if return_albedo:
        gw = np.ascontiguousarray(gweight, dtype=np.float64)
        tw = np.ascontiguousarray(tweight, dtype=np.float64)
    a = ShArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlayer, nwno, numg, numt, 1, ld
    (a.dtau, a.w0, a.ftau_cld, a.ftau_ray, a.f_deltaM, a.dtau_og, a.w0_og, a.cosb_og) = [addr(x) for x in lay]
    a.tau, a.tau_og = addr(lev[0]), addr(lev[1])
    a.surf_reflect, a.F0PI, a.b_top = addr(sr), addr(f0), addr(bt)
    a.ubar0, a.ubar1, a.gweight, a.tweight = addr(u0), addr(u1), addr(gw), addr(tw)
    a.cos_theta = float(cos_theta)
    a.w_single_form, a.w_multi_form, a.psingle_form = int(w_single_form), int(w_multi_form), int(psingle_form)
    a.w_single_rayleigh, a.w_multi_rayleigh, a.psingle_rayleigh = int(w_single_rayleigh), int(w_multi_rayleigh), int(psingle_rayleigh)
    a.frac_a, a.frac_b, a.frac_c = float(frac_a), float(frac_b), float(frac_c)
    a.constant_back, a.constant_forward = float(constant_back), float(constant_forward)
    a.stream, a.flx, a.single_form = int(stream), int(flx), int(single_form)
    a.xint_at_top, a.albedo, a.f_deltaM_out = addr(xint), addr(alb), addr(fd_out)
    flux = np.zeros((numg, numt, stream * nlevel, nwno)) if flx else None
    a.flux = addr(flux)
    if nwno > 0:
        ctx.check(ctx.lib.pb_reflected_sh(ctx.h, ctypes.byref(a), PB_HOST))
        if drift:
            f_deltaM[...] = fd_out
    if flux is None:
        flux = np.broadcast_to(_ZERO, (numg, numt, stream * nlevel, nwno))
    if return_albedo:
        return xint, flux, alb
    return xint, flux


def get_thermal_1d(nlevel, wno, nwno, numg, numt, tlevel, dtau, w0, cosb, plevel, ubar1,
                   surf_reflect, hard_surface, dwno, calc_type, *, ctx=None, level_fluxes=True,
                   gweight=None, tweight=None, return_thermal=False):
    """CUDA replacement of fluxes.get_thermal_1d (picaso/fluxes.py:1683-1912).

    Returns ``(flux_at_top[numg,numt,nwno], (flux_minus, flux_plus, flux_minus_mdpt,
    flux_plus_mdpt))`` like the reference (which always computes the level arrays).
    ``level_fluxes=False`` skips them (second element None) and runs the single-sweep
    TOA kernel only; ``return_thermal=True`` appends the fused compress_thermal vector.
    """
    ctx = ctx or _lib.default_context()
    nlayer = nlevel - 1
    dev = None
    if any(_is_dev(x) for x in (dtau, w0, cosb)):
        wn_h = np.ascontiguousarray(wno, dtype=np.float64)
        dw_h = _wvec(dwno, nwno) if (calc_type == 1 or np.ndim(dwno) > 0) else None
        dev = _resolve(ctx, [dtau, w0, cosb], nlayer, nwno, [wn_h, dw_h, _zero_or_vec(surf_reflect, nwno)])
    if dev is not None:
        return _thermal_device(ctx, dev, nlevel, nwno, numg, numt, tlevel, plevel, ubar1, hard_surface, calc_type,
                               level_fluxes, gweight, tweight, return_thermal)
    lay, ld = _layer_set([dtau, w0, cosb], nlayer, nwno)
    wn = np.ascontiguousarray(wno, dtype=np.float64)
    sr = _wvec(surf_reflect, nwno)
    dw = _wvec(dwno, nwno) if (calc_type == 1 or np.ndim(dwno) > 0) else None
    tl = np.ascontiguousarray(tlevel, dtype=np.float64)
    pl = np.ascontiguousarray(plevel, dtype=np.float64)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    ftop = np.zeros((numg, numt, nwno))
    lv = [np.zeros((numg, numt, nlevel, nwno)) for _ in range(4)] if level_fluxes else None
    th = np.zeros(nwno) if return_thermal else None
    gw = tw = None
    if return_thermal:
        gw = np.ascontiguousarray(gweight, dtype=np.float64)
        tw = np.ascontiguousarray(tweight, dtype=np.float64)
    a = ThermalArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlayer, nwno, numg, numt, 1, ld
    a.dtau, a.w0, a.cosb = [addr(x) for x in lay]
    a.wno, a.dwno, a.surf_reflect = addr(wn), addr(dw), addr(sr)
    a.tlevel, a.plevel, a.ubar1, a.gweight, a.tweight = addr(tl), addr(pl), addr(u1), addr(gw), addr(tw)
    a.hard_surface, a.calc_type = int(hard_surface), int(calc_type)
    a.flux_at_top, a.thermal = addr(ftop), addr(th)
    if lv is not None:
        a.flux_minus, a.flux_plus, a.flux_minus_mdpt, a.flux_plus_mdpt = [addr(x) for x in lv]
    if nwno > 0:
        ctx.check(ctx.lib.pb_thermal_toon_1d(ctx.h, ctypes.byref(a), PB_HOST))
    res = (ftop, tuple(lv) if lv is not None else None)
    if return_thermal:
        res = res + (th,)
    return res


def _thermal_device(ctx, dev, nlevel, nwno, numg, numt, tlevel, plevel, ubar1, hard_surface, calc_type,
                    level_fluxes, gweight, tweight, return_thermal):
    from .optics import DeviceArray
    ptrs, ld, memspace, vptr, keep = dev
    tl = np.ascontiguousarray(tlevel, dtype=np.float64)
    pl = np.ascontiguousarray(plevel, dtype=np.float64)
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64).reshape(-1)
    gw = tw = None
    if return_thermal:
        gw = np.ascontiguousarray(gweight, dtype=np.float64)
        tw = np.ascontiguousarray(tweight, dtype=np.float64)
    d_f = DeviceArray(ctx, (numg, numt, nwno))
    d_t = DeviceArray(ctx, (nwno,)) if return_thermal else None
    d_lv = [DeviceArray(ctx, (numg, numt, nlevel, nwno)) for _ in range(4)] if level_fluxes else None
    a = ThermalArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = nlevel - 1, nwno, numg, numt, 1, ld
    a.dtau, a.w0, a.cosb = ptrs
    a.wno, a.dwno, a.surf_reflect = vptr
    a.tlevel, a.plevel, a.ubar1, a.gweight, a.tweight = addr(tl), addr(pl), addr(u1), addr(gw), addr(tw)
    a.hard_surface, a.calc_type = int(hard_surface), int(calc_type)
    a.flux_at_top, a.thermal = d_f.ptr, (d_t.ptr if d_t else None)
    if d_lv:
        a.flux_minus, a.flux_plus, a.flux_minus_mdpt, a.flux_plus_mdpt = [x.ptr for x in d_lv]
    ctx.check(ctx.lib.pb_thermal_toon_1d(ctx.h, ctypes.byref(a), memspace))
    res = (d_f.numpy(), tuple(x.numpy() for x in d_lv) if d_lv else None)
    if return_thermal:
        res = res + (d_t.numpy(),)
    return res


def get_transit_1d(z, dz, nlevel, nwno, rstar, mmw, k_b, amu, player, tlayer, colden, DTAU, *,
                   ctx=None):
    """CUDA replacement of fluxes.get_transit_1d (picaso/fluxes.py:2582-2663): returns
    (Rp/Rs)^2 per wavelength.  `player`/`tlayer` take nlevel entries, as picaso() passes
    level pressure/temperature (justdoit.py:392-396); only the first nlevel-1 are read."""
    ctx = ctx or _lib.default_context()
    nlayer = nlevel - 1
    dev_in = _is_dev(DTAU)
    if dev_in:
        dt, ld = DTAU, nwno
    else:
        (dt,), ld = _layer_set([DTAU], nlayer, nwno)
    vec = lambda x, n: np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (n,)))
    z_, dz_ = vec(z, nlevel), vec(dz, nlevel)
    pl = np.zeros(nlevel); tl = np.ones(nlevel)
    p_in, t_in = np.asarray(player, dtype=np.float64), np.asarray(tlayer, dtype=np.float64)
    pl[:min(nlevel, p_in.size)] = p_in[:nlevel]
    tl[:min(nlevel, t_in.size)] = t_in[:nlevel]
    mm, cd = vec(mmw, nlayer), vec(colden, nlayer)
    F = np.zeros(nwno)
    a = TransitArgs()
    a.nlevel, a.nwno, a.nbatch, a.ld = nlevel, nwno, 1, ld
    a.z, a.dz, a.player, a.tlayer, a.mmw, a.colden = [addr(x) for x in (z_, dz_, pl, tl, mm, cd)]
    a.rstar, a.k_b, a.amu = float(rstar), float(k_b), float(amu)
    if dev_in:
        from .optics import DeviceArray
        d_F = DeviceArray(ctx, (nwno,))
        a.DTAU, a.F = dt.ptr, d_F.ptr
        ctx.check(ctx.lib.pb_transit_1d(ctx.h, ctypes.byref(a), PB_DEVICE))
        return d_F.numpy()
    a.DTAU, a.F = addr(dt), addr(F)
    if nwno > 0:
        ctx.check(ctx.lib.pb_transit_1d(ctx.h, ctypes.byref(a), PB_HOST))
    return F
