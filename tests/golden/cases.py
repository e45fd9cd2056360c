"""Seeded parity cases shared by make_golden.py (reference run, build container only),
the oracle tests (CPU) and the GPU parity tests.  Inputs are regenerated from seeds by
picaso_b200.synth; only reference OUTPUTS are stored in the .npz fixtures."""
import itertools

import numpy as np

from picaso_b200 import synth


def reflected_cases():
    cases = {}
    # every single_phase x multi_phase x toon_coefficients combination, small
    for sp, mp, tc in itertools.product((0, 1, 2, 3), (0, 1), (0, 1)):
        cases[f"refl_combo_sp{sp}_mp{mp}_tc{tc}"] = dict(
            build=dict(L=20, W=48, seed=100 + 8 * sp + 4 * mp + tc, phase=0.6),
            kw=dict(single_phase=sp, multi_phase=mp, toon_coefficients=tc, get_lvl_flux=0))
    # BASELINE config 1: 60 layers x 300 waves x 5 gauss angles, OTHG, no Raman
    cases["refl_cfg1"] = dict(build=dict(L=60, W=300, seed=1001),
                              kw=dict(single_phase=1, multi_phase=0, toon_coefficients=0,
                                      get_lvl_flux=0))
    cases["refl_cfg1_tthg_ray"] = dict(build=dict(L=60, W=300, seed=1001),
                                       kw=dict(single_phase=3, multi_phase=0, toon_coefficients=0,
                                               get_lvl_flux=0))
    # level fluxes (climate path), reflective surface, non-zero b_top
    cases["refl_lvl"] = dict(build=dict(L=24, W=40, seed=77, phase=0.3),
                             kw=dict(single_phase=3, multi_phase=0, toon_coefficients=0,
                                     get_lvl_flux=1),
                             surf_reflect=0.4)
    cases["refl_lvl_edd"] = dict(build=dict(L=16, W=33, seed=78, delta_eddington=False),
                                 kw=dict(single_phase=1, multi_phase=1, toon_coefficients=1,
                                         get_lvl_flux=1),
                                 surf_reflect=0.1)
    cases["refl_adversarial"] = dict(adversarial=True,
                                     kw=dict(single_phase=3, multi_phase=0, toon_coefficients=0,
                                             get_lvl_flux=0))
    cases["refl_ragged"] = dict(build=dict(L=7, W=37, seed=5, ngauss=7, phase=2.0),
                                kw=dict(single_phase=0, multi_phase=0, toon_coefficients=0,
                                        get_lvl_flux=0), surf_reflect=0.25)
    cases["refl_one_wave_one_layer"] = dict(build=dict(L=1, W=1, seed=6, ngauss=5),
                                            kw=dict(single_phase=3, multi_phase=0,
                                                    toon_coefficients=0, get_lvl_flux=1))
    return cases


def build_reflected(case):
    if case.get("adversarial"):
        d = synth.adversarial_reflected()
    else:
        d = synth.reflected_inputs(**case["build"])
    if "surf_reflect" in case:
        d["surf_reflect"] = np.full(d["nwno"], case["surf_reflect"])
    return d


def reflected_args(d, kw):
    """positional argument list of get_reflected_1d (fluxes.py:1010-1015)."""
    return (d["nlevel"], d["wno"], d["nwno"], d["numg"], d["numt"], d["dtau"], d["tau"], d["w0"],
            d["cosb"], d["gcos2"], d["ftau_cld"], d["ftau_ray"], d["dtau_og"], d["tau_og"],
            d["w0_og"], d["cosb_og"], d["surf_reflect"], d["ubar0"], d["ubar1"], d["cos_theta"],
            d["F0PI"], kw["single_phase"], kw["multi_phase"], d["frac_a"], d["frac_b"],
            d["frac_c"], d["constant_back"], d["constant_forward"], 1, kw["get_lvl_flux"],
            kw["toon_coefficients"], 0.0)


def thermal_cases():
    cases = {}
    for ct, hs in itertools.product((0, 1), (0, 1)):
        cases[f"therm_ct{ct}_hs{hs}"] = dict(build=dict(L=30, W=64, seed=200 + 2 * ct + hs),
                                             calc_type=ct, hard_surface=hs, surf_reflect=0.2 * hs)
    # BASELINE config 2 shape at reduced wave count
    cases["therm_cfg2_small"] = dict(build=dict(L=90, W=200, seed=1002), calc_type=0,
                                     hard_surface=0, surf_reflect=0.0)
    cases["therm_ragged"] = dict(build=dict(L=5, W=35, seed=9, ngauss=8), calc_type=0,
                                 hard_surface=0, surf_reflect=0.0)
    cases["therm_cold"] = dict(build=dict(L=12, W=40, seed=10, t_range=(60.0, 300.0),
                                          wno_range=(50.0, 30000.0)), calc_type=0,
                               hard_surface=0, surf_reflect=0.0)
    return cases


def build_thermal(case):
    d = synth.thermal_inputs(**case["build"])
    d["hard_surface"] = case["hard_surface"]
    d["surf_reflect"] = np.full(d["nwno"], case["surf_reflect"])
    d["calc_type"] = case["calc_type"]
    return d


def thermal_args(d):
    """positional argument list of get_thermal_1d (fluxes.py:1683-1684)."""
    return (d["nlevel"], d["wno"], d["nwno"], d["numg"], d["numt"], d["tlevel"], d["dtau"],
            d["w0"], d["cosb"], d["plevel"], d["ubar1"], d["surf_reflect"], d["hard_surface"],
            d["dwno"], d["calc_type"])


def transit_cases():
    return {
        "transit_cfg4_small": dict(L=80, W=256, seed=1004),
        "transit_ragged": dict(L=9, W=19, seed=12),
        "transit_two_level": dict(L=1, W=5, seed=13),
    }


def transit_args(d):
    """positional argument list of get_transit_1d (fluxes.py:2582-2583)."""
    return (d["z"], d["dz"], d["nlevel"], d["nwno"], d["rstar"], d["mmw"], d["k_b"], d["amu"],
            d["player"], d["tlayer"], d["colden"], d["DTAU"])


def sh_cases():
    """get_reflected_SH (fluxes.py:2675): forms = (w_single_form, w_multi_form, psingle_form,
    w_single_rayleigh, w_multi_rayleigh, psingle_rayleigh); 0 = TTHG / off, 1 = OTHG / on."""
    cases = {}
    forms_list = {"othg": (1, 1, 1, 1, 1, 1), "tthg": (0, 0, 0, 1, 1, 1), "mix1": (1, 0, 0, 0, 0, 0),
                  "mix2": (0, 1, 1, 1, 0, 1)}
    for stream in (2, 4):
        for fname, forms in forms_list.items():
            for sf in (0, 1):
                cases[f"sh{stream}_{fname}_sf{sf}"] = dict(
                    build=dict(L=18, W=24, seed=300 + stream, stream=stream, phase=0.5), forms=forms,
                    stream=stream, single_form=sf, surf_reflect=0.2)
        # BASELINE config 3 shape (60 layers, 5 angles, SH, "Raman on" = w0 with a Raman factor) at
        # reduced wave count; OTHG is drift-free, TTHG is the reference default
        cases[f"sh{stream}_cfg3_othg"] = dict(build=dict(L=60, W=48, seed=1003, stream=stream),
                                              forms=(1, 1, 1, 1, 1, 1), stream=stream, single_form=0,
                                              surf_reflect=0.0)
        cases[f"sh{stream}_cfg3_tthg"] = dict(build=dict(L=60, W=48, seed=1003, stream=stream),
                                              forms=(0, 0, 0, 1, 1, 1), stream=stream, single_form=0,
                                              surf_reflect=0.0)
        cases[f"sh{stream}_no_deltaM"] = dict(build=dict(L=9, W=33, seed=310, stream=stream,
                                                         delta_eddington=False, ngauss=7, phase=1.7),
                                              forms=(0, 0, 0, 1, 1, 1), stream=stream, single_form=0,
                                              surf_reflect=0.5)
        cases[f"sh{stream}_one_layer"] = dict(build=dict(L=1, W=3, seed=311, stream=stream),
                                              forms=(1, 1, 1, 1, 1, 1), stream=stream, single_form=0,
                                              surf_reflect=0.3)
    return cases


def build_sh(case):
    d = synth.reflected_inputs(**case["build"])
    d["surf_reflect"] = np.full(d["nwno"], case["surf_reflect"])
    return d


def sh_flux_cases():
    """get_reflected_SH(flx=1) (fluxes.py:2889-2890): the cases whose layer fluxes F.X + G are pinned
    (tests/golden/sh_flux.npz, make_golden_sh_flux.py)"""
    c = sh_cases()
    names = [f"sh{s}_{n}" for s in (2, 4) for n in ("othg_sf0", "tthg_sf0", "mix2_sf1", "cfg3_othg", "cfg3_tthg",
                                                     "no_deltaM", "one_layer")]
    return {n: c[n] for n in names}


def sh_args(d, case, flx=0):
    """positional argument list of get_reflected_SH (fluxes.py:2675-2679); f_deltaM is copied
    because the reference modifies it in place."""
    f = case["forms"]
    return (d["nlevel"], d["nwno"], d["numg"], d["numt"], d["dtau"], d["tau"], d["w0"], d["cosb"],
            d["ftau_cld"], d["ftau_ray"], d["f_deltaM"].copy(), d["dtau_og"], d["tau_og"], d["w0_og"],
            d["cosb_og"], d["surf_reflect"], d["ubar0"], d["ubar1"], d["cos_theta"], d["F0PI"],
            f[0], f[1], f[2], f[3], f[4], f[5], d["frac_a"], d["frac_b"], d["frac_c"],
            d["constant_back"], d["constant_forward"], case["stream"], 0.0, flx, case["single_form"])


def optics_cases():
    """opacity path (a9-a11): RetrieveOpacities.get_opacities ('linear' = bilinear in 1/T, log10 P;
    'nearest' = the reference default) + compute_opacity + compute_raman."""
    return {
        "opt_linear_raman": dict(db=dict(W=60, nmol=4, seed=2001), atm=dict(L=12, seed=2003), query="linear",
                                 raman=0, stream=2, dedd=True),
        "opt_nearest_noraman": dict(db=dict(W=45, nmol=3, seed=2011), atm=dict(L=9, seed=2013), query="nearest",
                                    raman=2, stream=4, dedd=True),
        "opt_linear_clear_nodedd": dict(db=dict(W=33, nmol=2, seed=2021, ragged=False),
                                        atm=dict(L=7, seed=2023, cloudy=False), query="linear", raman=2,
                                        stream=2, dedd=False),
    }


def thermal_sh_cases():
    """get_thermal_SH (fluxes.py:2979): stream x hard_surface x (cosb == cosb_og ?)"""
    cases = {}
    for stream in (2, 4):
        for hs in (0, 1):
            for same in (0, 1):
                cases[f"thsh{stream}_hs{hs}_same{same}"] = dict(build=dict(L=22, W=26, seed=400 + stream + hs),
                                                               stream=stream, hard_surface=hs, same=same,
                                                               surf_reflect=0.25 * hs)
        cases[f"thsh{stream}_cfg2_small"] = dict(build=dict(L=90, W=40, seed=1002), stream=stream,
                                                 hard_surface=0, same=0, surf_reflect=0.0)
        cases[f"thsh{stream}_one_layer"] = dict(build=dict(L=1, W=4, seed=411), stream=stream, hard_surface=0,
                                                same=1, surf_reflect=0.0)
    return cases


def build_thermal_sh(case):
    d = synth.thermal_inputs(**case["build"])
    d["cosb_og"] = d["cosb"].copy() if case["same"] else np.clip(d["cosb"] * 1.15 + 0.01, 0.0, 0.95)
    d["surf_reflect"] = np.full(d["nwno"], case["surf_reflect"])
    return d


def thermal_sh_args(d, case):
    """positional argument list of get_thermal_SH (fluxes.py:2979-2981)"""
    return (d["nlevel"], d["wno"], d["nwno"], d["numg"], d["numt"], d["tlevel"], d["dtau"], None, d["w0"],
            d["cosb"], None, None, None, d["w0"], d["cosb_og"], d["plevel"], d["ubar1"], d["surf_reflect"],
            case["stream"], case["hard_surface"])


def facet_geometry(ng, nt, phase):
    """get_angles_3d + compute_disco (disco.py:92-115, :36-50) in plain numpy"""
    i = np.linspace(1, nt, nt)
    tangle = np.cos(i * np.pi / (nt + 1))
    tweight = np.pi / (nt + 1) * np.sin(i * np.pi / (nt + 1)) ** 2.0
    gangle, gweight = np.polynomial.legendre.leggauss(ng)
    ct = np.cos(phase)
    lon = np.arcsin((gangle - (ct - 1.0) / (ct + 1.0)) / (2.0 / (ct + 1)))
    f = np.sin(np.arccos(tangle))
    return np.outer(np.cos(lon - phase), f), np.outer(np.cos(lon), f), float(ct), gweight, tweight


REFL_KEYS = ("dtau", "tau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og", "tau_og", "w0_og", "cosb_og")


def facets_cases():
    """get_reflected_3d / get_thermal_3d (fluxes.py:355, :2148): per-facet opacities [rows, nwno, ng, nt]"""
    cases = {}
    for sp in (0, 1, 2, 3):
        cases[f"refl3d_sp{sp}"] = dict(kind="refl", ng=3, nt=2, L=14, W=21, seed=500 + sp, phase=0.7, sp=sp,
                                       mp=sp % 2, surf=0.2)
    cases["refl3d_10x10ish"] = dict(kind="refl", ng=4, nt=3, L=30, W=40, seed=520, phase=1.9, sp=3, mp=0, surf=0.0)
    for hs in (0, 1):
        cases[f"therm3d_hs{hs}"] = dict(kind="therm", ng=3, nt=2, L=16, W=23, seed=540 + hs, phase=0.0, hs=hs,
                                        surf=0.3 * hs)
    return cases


def build_facets(case):
    ng, nt, L, W = case["ng"], case["nt"], case["L"], case["W"]
    ubar0, ubar1, ct, gweight, tweight = facet_geometry(ng, nt, case["phase"])
    out = dict(ng=ng, nt=nt, nlevel=L + 1, nwno=W, ubar0=ubar0, ubar1=ubar1, cos_theta=ct, gweight=gweight,
               tweight=tweight, surf_reflect=np.full(W, case["surf"]))
    if case["kind"] == "refl":
        arrs = {k: np.zeros((L + 1 if k in ("tau", "tau_og") else L, W, ng, nt)) for k in REFL_KEYS}
        for ig in range(ng):
            for it in range(nt):
                d = synth.reflected_inputs(L=L, W=W, seed=case["seed"] + 17 * ig + it)
                for k in REFL_KEYS:
                    arrs[k][:, :, ig, it] = d[k]
        out.update(arrs)
        out.update(wno=d["wno"], F0PI=d["F0PI"])
    else:
        arrs = {k: np.zeros((L, W, ng, nt)) for k in ("dtau", "w0", "cosb")}
        tl = np.zeros((L + 1, ng, nt))
        pl = np.zeros((L + 1, ng, nt))
        for ig in range(ng):
            for it in range(nt):
                d = synth.thermal_inputs(L=L, W=W, seed=case["seed"] + 17 * ig + it)
                for k in arrs:
                    arrs[k][:, :, ig, it] = d[k]
                tl[:, ig, it] = d["tlevel"] * (1 + 0.04 * ig - 0.02 * it)
                pl[:, ig, it] = d["plevel"] * (1 + 0.1 * it)
        out.update(arrs)
        out.update(wno=d["wno"], tlevel=tl, plevel=pl)
    return out


def facets_args(d, case):
    if case["kind"] == "refl":
        return (d["nlevel"], d["wno"], d["nwno"], d["ng"], d["nt"], *[d[k] for k in REFL_KEYS], d["surf_reflect"],
                d["ubar0"], d["ubar1"], d["cos_theta"], d["F0PI"], case["sp"], case["mp"], 1.0, -1.0, 2.0, -0.5, 1.0)
    return (d["nlevel"], d["wno"], d["nwno"], d["ng"], d["nt"], d["tlevel"], d["dtau"], d["w0"], d["cosb"],
            d["plevel"], d["ubar1"], d["surf_reflect"], case["hs"])


def ck_cases():
    """pre-mixed correlated-k: RetrieveCKs.get_pre_mix_ck + get_continuum + compute_opacity(ngauss=K)"""
    return {
        "ck8_dedd": dict(db=dict(W=40, K=8, seed=2101), atm=dict(L=11, seed=2103), stream=2, dedd=True),
        "ck4_clear": dict(db=dict(W=25, K=4, seed=2111), atm=dict(L=8, seed=2113, cloudy=False), stream=4, dedd=False),
    }


def mix_cases():
    """resort-rebin mixing (deq_chem.mix_all_gases_gasesfly + optics.mix_my_opacities_gasesfly)"""
    return {
        "mix_4gas_nk8": dict(W=6, K=8, ngas=4, L=5, seed=2201),
        "mix_2gas_nk8": dict(W=4, K=8, ngas=2, L=3, seed=2211),
        "mix_6gas_nk4": dict(W=5, K=4, ngas=6, L=4, seed=2221),
    }


def build_mix(case):
    """per-gas ln(kappa) tables [nP, nT, W, K] sharing one (P, T) grid + an atmosphere profile"""
    rng = np.random.default_rng(case["seed"])
    dbs = [synth.ck_database(W=case["W"], K=case["K"], seed=case["seed"] + 7 * g, nT=6, nP=6) for g in range(case["ngas"])]
    db = dbs[0]
    gases = synth.MOLECULES[:case["ngas"]]
    kappas = {m: d["kappa"] + 0.5 * g for g, (m, d) in enumerate(zip(gases, dbs))}
    atm = synth.atmosphere_profile(dict(db, molecules=gases), L=case["L"], seed=case["seed"] + 3, cloudy=False)
    # Gauss-Legendre points/weights on (0, 1), as opacity_factory.g_w_2gauss style arrays would provide
    x, w = np.polynomial.legendre.leggauss(case["K"])
    gauss_pts, gauss_wts = 0.5 * (x + 1.0), 0.5 * w
    return db, gases, kappas, atm, gauss_pts, gauss_wts


def climate_cases():
    """picaso.climate.get_fluxes (climate.py:1686-1952): all gauss points, visible + IR, cloud holes"""
    return {
        "clim_k4": dict(build=dict(L=24, W=70, K=4, seed=3001), reflected=True, thermal=True),
        "clim_k8_disk5": dict(build=dict(L=30, W=45, K=8, seed=3011, ng=5), reflected=True, thermal=True),
        "clim_k1_surf": dict(build=dict(L=12, W=33, K=1, seed=3021, surf=0.3), reflected=True, thermal=True),
        "clim_thermal_only_3d": dict(build=dict(L=10, W=20, K=2, seed=3031, ng=3, nt=2), reflected=False, thermal=True),
        "clim_reflected_only": dict(build=dict(L=16, W=50, K=3, seed=3041), reflected=True, thermal=False),
        "clim_holes": dict(build=dict(L=14, W=40, K=2, seed=3051), reflected=True, thermal=True, fhole=0.35),
    }


def build_climate(case):
    d = synth.climate_inputs(**case["build"])
    if "fhole" in case:
        c = synth.climate_inputs(**dict(case["build"], clear=True))
        d["hole_OpacityWEd"], d["hole_OpacityNoEd"] = c["OpacityWEd"], c["OpacityNoEd"]
    return d


def climate_args(d, case):
    args = [d["Atmosphere"], d["OpacityWEd"], d["OpacityNoEd"], d["ScatteringPhase"], d["Disco"], d["Opagrid"],
            d["F0PI"], case["reflected"], case["thermal"]]
    if "fhole" in case:
        args += [True, case["fhole"], d["hole_OpacityWEd"], d["hole_OpacityNoEd"]]
    return args


CLIMATE_OUT = ("flux_net_v_layer", "flux_net_v", "flux_plus_v", "flux_minus_v", "flux_net_ir_layer",
               "flux_net_ir", "flux_plus_ir", "flux_minus_ir")
