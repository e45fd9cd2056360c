"""Pre-mixed correlated-k opacity path: numpy oracle vs reference golden vectors (CPU); GPU vs golden."""
import numpy as np
import pytest

import cases as C
from oracle import optics as oo
from optics_util import OUT_NAMES, duck_atmosphere
from picaso_b200 import synth
from util import assert_close, golden


def _setup(name):
    case = C.ck_cases()[name]
    g = golden("ck")
    db = synth.ck_database(**case["db"])
    atm = synth.atmosphere_profile(dict(db, molecules=[]), **case["atm"])
    atm["cia_pairs"] = {a + b: (a, b) for a, b in db["continuum_molecules"]}
    ray = {m: g[f"{name}/in/rayleigh/{m}"] for m in db["rayleigh_molecules"]}
    return case, g, db, atm, ray


@pytest.mark.parametrize("name", sorted(C.ck_cases()))
def test_oracle_ck(name):
    case, g, db, atm, ray = _setup(name)
    pbar = atm["player"] / atm["pconv"]
    mol = oo.premix_ck(db["kappa"], *oo.ck_find_pts(db["pressures"], db["temps"], db["nc_p"], atm["tlayer"], pbar))
    assert_close(mol, g[f"{name}/molecular_opa"], 1e-12, name + " premixed kappa")
    cont = {}
    for k, tab in db["continuum"].items():
        cont[k] = oo.continuum_loglinear(db["cia_temps"], tab, atm["tlayer"])[0]
        assert_close(cont[k], g[f"{name}/continuum_opa/{k}"], 1e-12, name + " continuum " + k)
    res = oo.compute_opacity_ck(atm, mol, cont, ray, stream=case["stream"], delta_eddington=case["dedd"])
    for n, arr in zip(OUT_NAMES, res):
        assert_close(arr, g[f"{name}/out/{n}"], 1e-11, name + " " + n)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(C.ck_cases()))
def test_gpu_ck(name):
    import picaso_b200 as pb
    case, g, db, atm, ray = _setup(name)
    opa = pb.DeviceCKs(db["wno"], db["pressures"], db["temps"], db["nc_p"], db["kappa"], db["gauss_wts"],
                       db["cia_temps"], db["continuum"], ray)
    a = duck_atmosphere(dict(db, molecules=["H2O", "CH4"]), atm)   # profile molecules are ignored by pre-mixed CK
    opa.get_opacities(a)
    res = pb.compute_opacity(a, opa, ngauss=db["ngauss"], stream=case["stream"], delta_eddington=case["dedd"],
                             test_mode=None, raman=2)
    for n, arr in zip(OUT_NAMES, res):
        want = g[f"{name}/out/{n}"]
        assert arr.shape == want.shape
        assert_close(arr, want, 1e-10, name + " " + n)
    # gauss-point slices feed the flux solvers like picaso() does (justdoit.py:256-283)
    import oracle
    L, W = atm["nlayer"], db["nwno"]
    gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = synth.geometry_1d(5, 0.0)
    ig = db["ngauss"] - 1
    ft, _ = pb.get_thermal_1d(L + 1, db["wno"], W, 5, 1, atm["tlevel"], res[7][:, :, ig], res[11][:, :, ig],
                              res[10][:, :, ig], atm["plevel"], ubar1, 0, 0, 0.0, 0, level_fluxes=False)
    oft, _ = oracle.get_thermal_1d(L + 1, db["wno"], W, 5, 1, atm["tlevel"], g[f"{name}/out/DTAU_OG"][:, :, ig],
                                   g[f"{name}/out/W0_no_raman"][:, :, ig], g[f"{name}/out/COSB_OG"][:, :, ig],
                                   atm["plevel"], ubar1, 0, 0, np.zeros(W), 0, level_fluxes=False)
    assert_close(ft, oft, 1e-6, name + " thermal from CK slice")
    opa.close()
