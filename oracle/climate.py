"""Oracle of the climate solver's RT call - TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).

Restates picaso/climate.py:1686-1952 (get_fluxes) in numpy over the C oracle's get_reflected_1d /
get_thermal_1d: the same per-gauss-point loop, the same order of the gauss-weight accumulation,
np.sum(axis=3) for the visible net fluxes, compress_thermal and the sequential dwni loop for the IR
ones.  Pinned to the reference's own get_fluxes (tests/golden/make_golden_climate.py -> climate.npz).
"""
import numpy as np

from . import compress_thermal, get_reflected_1d, get_thermal_1d


def get_fluxes(Atmosphere, OpacityWEd, OpacityNoEd, ScatteringPhase, Disco, Opagrid, F0PI, reflected, thermal,
               do_holes=False, fhole=0.0, hole_OpacityWEd=None, hole_OpacityNoEd=None, nthreads=1, quad=False):
    """quad=True: the level fluxes come from the binary128 build of the C oracle (the exact-arithmetic
    yardstick of tests/util.py); the reductions stay fp64."""
    pressure, temperature, nlevel = Atmosphere.p_level, Atmosphere.t_level, Atmosphere.nlevel
    W, N, S = OpacityWEd, OpacityNoEd, ScatteringPhase
    ng, nt = Disco.ng, Disco.nt
    nwno, dwni, wno, ngauss, gauss_wts = Opagrid.nwno, Opagrid.delta_wno, Opagrid.wno, Opagrid.ngauss, Opagrid.gauss_wts
    flux_net_v = np.zeros((ng, nt, nlevel))
    flux_net_v_layer = np.zeros((ng, nt, nlevel))
    flux_plus_v = np.zeros((ng, nt, nlevel, nwno))
    flux_minus_v = np.zeros((ng, nt, nlevel, nwno))
    flux_plus_midpt = np.zeros((ng, nt, nlevel, nwno))
    flux_minus_midpt = np.zeros((ng, nt, nlevel, nwno))
    flux_plus = np.zeros((ng, nt, nlevel, nwno))
    flux_minus = np.zeros((ng, nt, nlevel, nwno))
    flux_net_ir = np.zeros(nlevel)
    flux_net_ir_layer = np.zeros(nlevel)
    flux_plus_ir = np.zeros((nlevel, nwno))
    flux_minus_ir = np.zeros((nlevel, nwno))
    sl = lambda a, ig: np.ascontiguousarray(a[:, :, ig])

    def refl(Wt, Nt, ig):  # climate.py:1803-1816
        half = np.full((1, 1), 0.5)
        _, lv = get_reflected_1d(nlevel, wno, nwno, 1, 1, sl(Wt.DTAU, ig), sl(Wt.TAU, ig), sl(Wt.W0, ig),
                                 sl(Wt.COSB, ig), sl(Wt.GCOS2, ig), sl(Wt.ftau_cld, ig), sl(Wt.ftau_ray, ig),
                                 sl(Nt.DTAU, ig), sl(Nt.TAU, ig), sl(Nt.W0, ig), sl(Nt.COSB, ig), S.surf_reflect,
                                 half, half, Disco.cos_theta, F0PI, S.single_phase, S.multi_phase, S.frac_a,
                                 S.frac_b, S.frac_c, S.constant_back, S.constant_forward, get_toa_intensity=0,
                                 get_lvl_flux=1, nthreads=nthreads, quad=quad)
        return lv

    def therm(Wt, Nt, ig):  # climate.py:1887-1892
        _, lv = get_thermal_1d(nlevel, wno, nwno, ng, nt, temperature, sl(Nt.DTAU, ig), sl(Wt.W0_no_raman, ig),
                               sl(Nt.COSB, ig), pressure, Disco.ubar1, S.surf_reflect, 0, dwni, 1,
                               nthreads=nthreads, quad=quad)
        return lv

    def mix(a, b):
        return [(1.0 - fhole) * x + fhole * y for x, y in zip(a, b)]

    if reflected:
        for ig in range(ngauss):
            fm, fp, fmm, fpm = refl(W, N, ig)
            if do_holes:
                fm, fp, fmm, fpm = mix((fm, fp, fmm, fpm), refl(hole_OpacityWEd, hole_OpacityNoEd, ig))
            flux_net_v_layer += (np.sum(fpm, axis=3) - np.sum(fmm, axis=3)) * gauss_wts[ig]
            flux_net_v += (np.sum(fp, axis=3) - np.sum(fm, axis=3)) * gauss_wts[ig]
            flux_plus_v += fp * gauss_wts[ig]
            flux_minus_v += fm * gauss_wts[ig]
    if thermal:
        for ig in range(ngauss):
            fm, fp, fmm, fpm = therm(W, N, ig)
            if do_holes:
                fm, fp, fmm, fpm = mix((fm, fp, fmm, fpm), therm(hole_OpacityWEd, hole_OpacityNoEd, ig))
            flux_plus += fp * gauss_wts[ig]
            flux_minus += fm * gauss_wts[ig]
            flux_plus_midpt += fpm * gauss_wts[ig]
            flux_minus_midpt += fmm * gauss_wts[ig]
        flux_plus = compress_thermal(nwno, flux_plus, Disco.gweight, Disco.tweight)
        flux_minus = compress_thermal(nwno, flux_minus, Disco.gweight, Disco.tweight)
        flux_plus_midpt = compress_thermal(nwno, flux_plus_midpt, Disco.gweight, Disco.tweight)
        flux_minus_midpt = compress_thermal(nwno, flux_minus_midpt, Disco.gweight, Disco.tweight)
        for wvi in range(nwno):
            flux_net_ir_layer += (flux_plus_midpt[:, wvi] - flux_minus_midpt[:, wvi]) * dwni[wvi]
            flux_net_ir += (flux_plus[:, wvi] - flux_minus[:, wvi]) * dwni[wvi]
            flux_plus_ir[:, wvi] += flux_plus[:, wvi] * dwni[wvi]
            flux_minus_ir[:, wvi] += flux_minus[:, wvi] * dwni[wvi]
    return (flux_net_v_layer, flux_net_v, flux_plus_v, flux_minus_v, flux_net_ir_layer, flux_net_ir,
            flux_plus_ir, flux_minus_ir)
