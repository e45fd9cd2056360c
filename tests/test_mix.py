"""Resort-rebin k-mixing (deq_chem.mix_all_gases_gasesfly + RetrieveCKs.mix_my_opacities_gasesfly):
numpy oracle vs reference golden vectors (CPU); CUDA path vs golden and oracle (GPU)."""
import numpy as np
import pytest

import cases as C
from oracle import optics as oo
from oracle import resort_rebin as rr
from optics_util import OUT_NAMES, duck_atmosphere
from picaso_b200 import synth
from util import assert_close, golden

RTOL = 1e-10   # fp64; the sort/resample is exact up to the rounding of log10/exp10 and the prefix sum


def _oracle_mix(case):
    db, gases, kappas, atm, gp, gw = C.build_mix(case)
    ti, pi, pl, tl, ph, th = oo.ck_find_pts(db["pressures"], db["temps"], db["nc_p"], atm["tlayer"],
                                            atm["player"] / atm["pconv"])
    lnm = rr.mix_all_gases([kappas[m] for m in gases], [atm["mixingratios"][m] for m in gases], gp, gw, (pl, ph, tl, th))
    return (pl, ph, tl, th), lnm, rr.interpolate_mixed(lnm, ti, pi)


@pytest.mark.parametrize("name", sorted(C.mix_cases()))
def test_oracle_mix(name):
    g = golden("mix")
    ind, lnm, mol = _oracle_mix(C.mix_cases()[name])
    assert np.array_equal(np.array(ind), g[name + "/indices"])
    assert_close(mol, g[name + "/molecular_opa"], 1e-12, name + " molecular_opa")


def test_mix_2_gases_properties():
    """identical gases mix to themselves at the resampled quantiles; a trace second gas leaves the first"""
    x, w = np.polynomial.legendre.leggauss(8)
    gp, gw = 0.5 * (x + 1), 0.5 * w
    k = np.exp(np.linspace(-30, -20, 8))
    out, mt = rr.mix_2_gases(k, k * 1e-30, 1.0, 1e-30, gp, gw)
    assert mt == 1.0
    # the Nk^2 products collapse onto Nk plateaus of weight w_i; resampling returns values inside [k0, k7]
    assert np.all(out >= k[0] * (1 - 1e-12)) and np.all(out <= k[-1] * (1 + 1e-12))
    assert np.all(np.diff(out) >= 0)


def _device(case):
    import picaso_b200 as pb
    db, gases, kappas, atm, gp, gw = C.build_mix(case)
    ray = {m: np.full(db["nwno"], 1e-27 * (i + 1)) for i, m in enumerate(db["rayleigh_molecules"])}
    opa = pb.DeviceGasCKs(db["wno"], db["pressures"], db["temps"], db["nc_p"], kappas, gp, gw, db["cia_temps"],
                          db["continuum"], ray)
    atm["cia_pairs"] = {a + b: (a, b) for a, b in db["continuum_molecules"]}
    a = duck_atmosphere(dict(db, molecules=gases), atm)
    return pb, opa, a, db, atm, ray


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(C.mix_cases()))
def test_gpu_mix(name):
    g = golden("mix")
    case = C.mix_cases()[name]
    pb, opa, a, db, atm, ray = _device(case)
    ind, ti, pi = opa.get_mixing_indices(a)
    assert np.array_equal(ind, g[name + "/indices"])
    mol, lnm = opa.mix_my_opacities_gasesfly(a, return_ln_mixed=True)
    assert_close(mol, g[name + "/molecular_opa"], RTOL, name + " molecular_opa")
    _, o_lnm, _ = _oracle_mix(case)
    assert_close(lnm, o_lnm, RTOL, name + " ln kappa_mixed")
    # device-resident output is the same bits
    dmol = opa.mix_my_opacities_gasesfly(a, device_output=True)
    assert np.array_equal(dmol.numpy(), mol)
    opa.close()


@pytest.mark.gpu
def test_gpu_mix_exclude_and_order():
    """exclude_mol drops a gas (optics.py:1177-1181); the fold order follows atmosphere.molecules"""
    name = "mix_4gas_nk8"
    case = C.mix_cases()[name]
    pb, opa, a, db, atm, ray = _device(case)
    dbm, gases, kappas, atm2, gp, gw = C.build_mix(case)
    ti, pi, pl, tl, ph, th = oo.ck_find_pts(db["pressures"], db["temps"], db["nc_p"], atm["tlayer"],
                                            atm["player"] / atm["pconv"])
    keep = [m for m in gases if m != gases[1]]
    want = rr.interpolate_mixed(rr.mix_all_gases([kappas[m] for m in keep], [atm["mixingratios"][m] for m in keep],
                                                 gp, gw, (pl, ph, tl, th)), ti, pi)
    got = opa.mix_my_opacities_gasesfly(a, exclude_mol={m: (0 if m == gases[1] else 1) for m in gases})
    assert_close(got, want, RTOL, "excluded gas")
    a.molecules = list(reversed(gases))
    rev = list(reversed(gases))
    want = rr.interpolate_mixed(rr.mix_all_gases([kappas[m] for m in rev], [atm["mixingratios"][m] for m in rev],
                                                 gp, gw, (pl, ph, tl, th)), ti, pi)
    assert_close(opa.mix_my_opacities_gasesfly(a), want, RTOL, "reversed order")
    opa.close()


@pytest.mark.gpu
def test_gpu_mix_ties_and_single_gas():
    """equal k values (ties resolved by flat index like numpy's stable mergesort) and the 1-gas pass-through"""
    import picaso_b200 as pb
    case = dict(W=3, K=8, ngas=3, L=3, seed=2301)
    db, gases, kappas, atm, gp, gw = C.build_mix(case)
    flat = {m: np.full_like(k, -50.0 - i) for i, (m, k) in enumerate(kappas.items())}   # every product ties
    flat[gases[2]] = flat[gases[0]].copy()
    ray = {m: np.zeros(db["nwno"]) for m in db["rayleigh_molecules"]}
    opa = pb.DeviceGasCKs(db["wno"], db["pressures"], db["temps"], db["nc_p"], flat, gp, gw, db["cia_temps"],
                          db["continuum"], ray)
    a = duck_atmosphere(dict(db, molecules=gases), atm)
    ti, pi, pl, tl, ph, th = oo.ck_find_pts(db["pressures"], db["temps"], db["nc_p"], atm["tlayer"],
                                            atm["player"] / atm["pconv"])
    want = rr.interpolate_mixed(rr.mix_all_gases([flat[m] for m in gases], [atm["mixingratios"][m] for m in gases],
                                                 gp, gw, (pl, ph, tl, th)), ti, pi)
    assert_close(opa.mix_my_opacities_gasesfly(a), want, RTOL, "ties")
    a.molecules = gases[:1]
    got = opa.mix_my_opacities_gasesfly(a)
    assert_close(got, np.exp(-50.0) * rr.N_A * np.ones_like(got), 1e-13, "single gas")
    opa.close()


@pytest.mark.gpu
@pytest.mark.parametrize("K", [8, 4])
def test_gpu_mix_tables_not_monotone_in_g(K):
    """Real correlated-k tables are non-decreasing in the gauss index and so are the synthetic ones; the
    sorting network must not rely on it (skipping its first stages for pre-sorted rows was tried and was
    slower: 3.07 vs 2.89 ms): tables whose gauss axis is shuffled still match the oracle."""
    import picaso_b200 as pb
    case = dict(W=5, K=K, ngas=4, L=4, seed=2401)
    db, gases, kappas, atm, gp, gw = C.build_mix(case)
    rng = np.random.default_rng(9)
    shuf = {m: (k[..., rng.permutation(K)] if i % 2 else k) for i, (m, k) in enumerate(kappas.items())}
    ray = {m: np.zeros(db["nwno"]) for m in db["rayleigh_molecules"]}
    opa = pb.DeviceGasCKs(db["wno"], db["pressures"], db["temps"], db["nc_p"], shuf, gp, gw, db["cia_temps"],
                          db["continuum"], ray)
    a = duck_atmosphere(dict(db, molecules=gases), atm)
    ti, pi, pl, tl, ph, th = oo.ck_find_pts(db["pressures"], db["temps"], db["nc_p"], atm["tlayer"],
                                            atm["player"] / atm["pconv"])
    want = rr.interpolate_mixed(rr.mix_all_gases([shuf[m] for m in gases], [atm["mixingratios"][m] for m in gases],
                                                 gp, gw, (pl, ph, tl, th)), ti, pi)
    assert_close(opa.mix_my_opacities_gasesfly(a), want, RTOL, "shuffled gauss axis")
    opa.close()


@pytest.mark.gpu
def test_gpu_mix_into_compute_opacity():
    """get_opacities (deq on-the-fly) keeps molecular_opa in HBM and compute_opacity consumes it"""
    name = "mix_4gas_nk8"
    g = golden("mix")
    pb, opa, a, db, atm, ray = _device(C.mix_cases()[name])
    opa.get_opacities(a)
    res = pb.compute_opacity(a, opa, ngauss=8, stream=2, delta_eddington=True, test_mode=None, raman=2)
    cont = {k: oo.continuum_loglinear(db["cia_temps"], tab, atm["tlayer"])[0] for k, tab in db["continuum"].items()}
    want = oo.compute_opacity_ck(atm, g[name + "/molecular_opa"], cont, ray, stream=2, delta_eddington=True)
    for n, got, w in zip(OUT_NAMES, res, want):
        assert_close(got, w, 1e-9, name + " " + n)
    opa.close()


@pytest.mark.gpu
def test_gpu_mix_errors():
    import picaso_b200 as pb
    pb_, opa, a, db, atm, ray = _device(C.mix_cases()["mix_2gas_nk8"])
    a.molecules = ["H2O", "NotAGas"]
    with pytest.raises(KeyError):
        opa.mix_my_opacities_gasesfly(a)
    opa.close()
