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
