"""CPU: the numpy restatement of the opacity path (oracle/optics.py) against golden vectors made
by the unmodified reference classes (tests/golden/make_golden_optics.py)."""
import numpy as np
import pytest

import cases as C
from oracle import optics as oo
from optics_util import OUT_NAMES, load_case
from util import assert_close


@pytest.mark.parametrize("name", sorted(C.optics_cases()))
def test_opacity_path(name):
    case, g, db, atm, ins = load_case(name)
    pbar = atm["player"] / atm["pconv"]
    if case["query"] == "linear":
        ti, pi, ill, ihl, ilh, ihh = oo.find_needed_pts(db["temps"], db["pressures"], db["nc_p"],
                                                        atm["tlayer"], pbar)
        want = 1 + np.unique(np.concatenate([ill, ihl, ilh, ihh]))
        assert np.array_equal(want, g[f"{name}/pt_opa_index"])
        mol = {m: oo.interp_molecular(db["tables"][m], ti, pi, ill, ihl, ilh, ihh) for m in db["molecules"]}
    else:
        ind = oo.nearest_pt(db["pt_pairs"], atm["tlayer"], pbar)
        assert np.array_equal(ind + 1, g[f"{name}/pt_opa_index"])
        mol = {m: oo.nearest_molecular(db["tables"][m], ind) for m in db["molecules"]}
    for m in db["molecules"]:
        assert_close(mol[m], g[f"{name}/molecular_opa/{m}"], 1e-12, name + " molecular " + m)
    ic = oo.nearest_cia_temp(db["cia_temps"], atm["tlayer"])
    cont = {k: db["continuum"][k][ic] for k in db["continuum"]}
    for k in cont:
        assert np.array_equal(cont[k], g[f"{name}/continuum_opa/{k}"])
    rf = None
    if case["raman"] == 0:
        rf = oo.compute_raman(db["nwno"], atm["nlayer"], db["wno"], ins["stellar_shifts"], atm["tlayer"],
                              ins["raman_c"], ins["raman_ji"], ins["raman_deltanu"])
        assert_close(rf, g[f"{name}/raman_factor"], 1e-12, name + " raman factor")
    res = oo.compute_opacity(atm, mol, cont, ins["rayleigh"], rf, stream=case["stream"],
                             delta_eddington=case["dedd"])
    for n, arr in zip(OUT_NAMES, res):
        assert_close(arr, g[f"{name}/out/{n}"], 1e-11, name + " " + n)


def test_product_bin_search_matches_oracle():
    """the vectorised (all layers at once) find_needed_pts of the product's host mirror against the oracle's
    restatement of RetrieveOpacities.find_needed_pts, on ragged grids and profiles that leave the grid"""
    from picaso_b200.optics import find_needed_pts_grid
    from picaso_b200 import synth
    rng = np.random.default_rng(12)
    for seed, ragged in ((1, True), (2, False), (3, True)):
        db = synth.opacity_database(W=8, nmol=1, seed=seed, nT=14, nP=11, ragged=ragged)
        tlayer = np.concatenate([rng.uniform(40.0, 5000.0, size=60), db["temps"][:3], [db["temps"][-1]]])
        pbar = np.concatenate([10.0 ** rng.uniform(-7.5, 4.0, size=60), db["pressures"][:3], [db["pressures"][-1]]])
        want = oo.find_needed_pts(db["temps"], db["pressures"], db["nc_p"], tlayer, pbar)
        got = find_needed_pts_grid(1 / db["temps"], np.log10(db["pressures"]), db["nc_p"], tlayer, pbar)
        for a, b in zip(got, want):
            assert np.array_equal(np.asarray(a).ravel(), np.asarray(b).ravel())   # reference shape [:, None] vs flat


def test_pollack_raman_and_full_output():
    """compute_opacity(raman=1 'pollack', full_output=True): oracle against the unmodified reference run on a synthetic
    raman_fortran.txt (tests/golden/make_golden_pollack.py)"""
    from util import golden
    g = golden("pollack")
    case, _, db, atm, ins = load_case(str(g["case"]))
    pbar = atm["player"] / atm["pconv"]
    ti, pi, ill, ihl, ilh, ihh = oo.find_needed_pts(db["temps"], db["pressures"], db["nc_p"], atm["tlayer"], pbar)
    mol = {m: oo.interp_molecular(db["tables"][m], ti, pi, ill, ihl, ilh, ihh) for m in db["molecules"]}
    ic = oo.nearest_cia_temp(db["cia_temps"], atm["tlayer"])
    cont = {k: db["continuum"][k][ic] for k in db["continuum"]}
    rf = oo.raman_pollack(db["wno"], g["table_w"], g["table_f"], atm["nlayer"])
    assert (rf > 0.99999).any() and (rf < 0.99999).any()      # the cap is exercised
    full = {}
    res = oo.compute_opacity(atm, mol, cont, ins["rayleigh"], rf, stream=case["stream"], delta_eddington=case["dedd"],
                             full=full)
    for n, arr in zip(OUT_NAMES, res):
        assert_close(arr, g["out/" + n], 1e-11, "pollack " + n)
    for n in ("taugas", "tauray", "taucld"):
        assert_close(full[n], g["full/" + n], 1e-11, "full_output " + n)


@pytest.mark.parametrize("mode", ["rayleigh", "constant_tau"])
@pytest.mark.parametrize("name", ["opt_linear_raman", "opt_nearest_noraman"])
def test_compute_opacity_test_modes(name, mode):
    """compute_opacity(test_mode=...) (optics.py:372-399): oracle against the unmodified reference
    (tests/golden/make_golden_testmode.py), including the in-place w0 <= 0 -> 1e-10 replacement"""
    from util import golden
    g = golden("testmode")
    case, _, db, atm, ins = load_case(name)
    atm["cloud_w0"][::3, ::5] = 0.0
    pbar = atm["player"] / atm["pconv"]
    if case["query"] == "linear":
        ti, pi, ill, ihl, ilh, ihh = oo.find_needed_pts(db["temps"], db["pressures"], db["nc_p"], atm["tlayer"], pbar)
        mol = {m: oo.interp_molecular(db["tables"][m], ti, pi, ill, ihl, ilh, ihh) for m in db["molecules"]}
    else:
        mol = {m: oo.nearest_molecular(db["tables"][m], oo.nearest_pt(db["pt_pairs"], atm["tlayer"], pbar)) for m in db["molecules"]}
    ic = oo.nearest_cia_temp(db["cia_temps"], atm["tlayer"])
    cont = {k: db["continuum"][k][ic] for k in db["continuum"]}
    res = oo.compute_opacity(atm, mol, cont, ins["rayleigh"], None, stream=case["stream"], delta_eddington=case["dedd"],
                             test_mode=mode)
    for n, arr in zip(OUT_NAMES, res):
        assert_close(arr, g[f"{name}/{mode}/{n}"], 1e-12, "%s test_mode=%s %s" % (name, mode, n))
    assert np.array_equal(atm["cloud_w0"], g[f"{name}/{mode}/cloud_w0_after"])
