"""GPU: device-resident opacity path (pb_optab_* + pb_compute_opacity through
picaso_b200.DeviceOpacities / compute_opacity) against the reference golden vectors, and the
fully device-resident chain opacity -> flux solvers against the CPU oracle."""
import numpy as np
import pytest

import cases as C
import oracle
import picaso_b200 as pb
from optics_util import OUT_NAMES, device_opacities, duck_atmosphere, load_case
from picaso_b200 import synth
from util import assert_close

pytestmark = pytest.mark.gpu
RTOL = 1e-6          # north_star tolerance; the opacity arrays actually agree to ~1e-13


@pytest.mark.parametrize("name", sorted(C.optics_cases()))
def test_compute_opacity_vs_reference(name):
    case, g, db, atm, ins = load_case(name)
    opa = device_opacities(pb, case, db, ins)
    a = duck_atmosphere(db, atm)
    opa.get_opacities(a)
    assert np.array_equal(np.asarray(a.layer["pt_opa_index"]), g[f"{name}/pt_opa_index"])
    res = pb.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"],
                             test_mode=None, raman=case["raman"])
    assert len(res) == 13
    for n, arr in zip(OUT_NAMES, res):
        want = g[f"{name}/out/{n}"]
        assert arr.shape == want.shape + (1,)
        assert_close(arr[:, :, 0], want, 1e-10, name + " " + n)
    # restricted output set (what transit needs) gives the same DTAU_OG
    only = pb.compute_opacity(a, opa, stream=case["stream"], delta_eddington=case["dedd"], test_mode=None,
                              raman=case["raman"], outputs=("DTAU_OG",))
    assert only[0] is None and np.array_equal(only[7], res[7])
    opa.close()


def test_device_resident_chain_matches_oracle():
    """profile -> opacity kernel -> reflected / thermal / transit kernels without any
    O(nlayer x nwno) host transfer, against oracle flux solvers fed the golden opacity arrays."""
    name = "opt_linear_raman"
    case, g, db, atm, ins = load_case(name)
    opa = device_opacities(pb, case, db, ins)
    a = duck_atmosphere(db, atm)
    opa.get_opacities(a)
    dev = pb.compute_opacity(a, opa, stream=2, delta_eddington=True, test_mode=None, raman=0, device_outputs=True)
    (DTAU, TAU, W0, COSB, ftau_cld, ftau_ray, GCOS2, DTAU_OG, TAU_OG, W0_OG, COSB_OG, W0_no_raman, f_deltaM) = dev
    ref = {n: g[f"{name}/out/{n}"] for n in OUT_NAMES}
    L, W = atm["nlayer"], db["nwno"]
    gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = synth.geometry_1d(5, 0.0)
    F0PI = np.ones(W)
    args_tail = (0, ubar0, ubar1, cos_theta, F0PI, 3, 0, 1.0, -1.0, 2.0, -0.5, 1.0)
    xint, _, alb = pb.get_reflected_1d(L + 1, db["wno"], W, 5, 1, DTAU[:, :, 0], TAU[:, :, 0], W0[:, :, 0],
                                       COSB[:, :, 0], GCOS2[:, :, 0], ftau_cld[:, :, 0], ftau_ray[:, :, 0],
                                       DTAU_OG[:, :, 0], TAU_OG[:, :, 0], W0_OG[:, :, 0], COSB_OG[:, :, 0],
                                       *args_tail, gweight=gweight, tweight=tweight, return_albedo=True)
    ox, _ = oracle.get_reflected_1d(L + 1, db["wno"], W, 5, 1, ref["DTAU"], ref["TAU"], ref["W0"], ref["COSB"],
                                    ref["GCOS2"], ref["ftau_cld"], ref["ftau_ray"], ref["DTAU_OG"], ref["TAU_OG"],
                                    ref["W0_OG"], ref["COSB_OG"], *args_tail)
    assert_close(xint, ox, RTOL, "chain reflected")
    assert_close(alb, oracle.compress_disco(W, cos_theta, ox, gweight, tweight, F0PI), RTOL, "chain albedo")
    ft, none, th = pb.get_thermal_1d(L + 1, db["wno"], W, 5, 1, atm["tlevel"], DTAU_OG, W0_no_raman, COSB_OG,
                                     atm["plevel"], ubar1, 0, 0, 0.0, 0, level_fluxes=False, gweight=gweight,
                                     tweight=tweight, return_thermal=True)
    oft, _ = oracle.get_thermal_1d(L + 1, db["wno"], W, 5, 1, atm["tlevel"], ref["DTAU_OG"], ref["W0_no_raman"],
                                   ref["COSB_OG"], atm["plevel"], ubar1, 0, 0, np.zeros(W), 0, level_fluxes=False)
    assert_close(ft, oft, RTOL, "chain thermal")
    z = np.linspace(8.0e9, 7.0e9, L + 1)
    dz = np.full(L + 1, (z[0] - z[1]))
    tr_args = (z, dz, L + 1, W, 6.957e10, atm["mmw"], atm["k_b"], atm["amu"], atm["plevel"], atm["tlevel"],
               atm["colden"])
    F = pb.get_transit_1d(*tr_args, DTAU_OG)
    assert_close(F, oracle.get_transit_1d(*tr_args, ref["DTAU_OG"]), RTOL, "chain transit")
    opa.close()


@pytest.mark.parametrize("name", ["opt_linear_raman", "opt_nearest_noraman", "opt_linear_clear_nodedd"])
def test_reflected_spectrum_one_call(name):
    """pb_spectrum_reflected (opacity -> flux -> disk integration behind one C call) returns the bits of the
    compute_opacity(device_outputs=True) -> get_reflected_1d(return_albedo=True) chain, for scalar / vector / default
    surface reflectivity and stellar flux"""
    case, g, db, atm, ins = load_case(name)
    opa = device_opacities(pb, case, db, ins)
    a = duck_atmosphere(db, atm)
    L, W = atm["nlayer"], db["nwno"]
    gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = synth.geometry_1d(5, 0.0)
    rng = np.random.default_rng(5)
    for F0PI, surf in ((None, None), (1.0 + rng.random(W), 0.3), (None, 0.5 * rng.random(W))):
        opa.get_opacities(a)
        dev = pb.compute_opacity(a, opa, stream=case["stream"], delta_eddington=case["dedd"], test_mode=None,
                                 raman=case["raman"], device_outputs=True)
        sl = [d[:, :, 0] for d in dev[:11]]
        DTAU, TAU, W0, COSB, fcld, fray, GCOS2, DTAU_OG, TAU_OG, W0_OG, COSB_OG = sl
        x0, _, alb0 = pb.get_reflected_1d(L + 1, db["wno"], W, 5, 1, DTAU, TAU, W0, COSB, GCOS2, fcld, fray, DTAU_OG, TAU_OG,
                                          W0_OG, COSB_OG, 0 if surf is None else surf, ubar0, ubar1, cos_theta,
                                          np.ones(W) if F0PI is None else F0PI, 3, 0, 1.0, -1.0, 2.0, -0.5, 1.0,
                                          gweight=gweight, tweight=tweight, return_albedo=True)
        opa.get_opacities(a)
        alb, xint = pb.reflected_spectrum(a, opa, ubar0, ubar1, cos_theta, gweight, tweight, F0PI=F0PI, surf_reflect=surf,
                                          stream=case["stream"], delta_eddington=case["dedd"], raman=case["raman"],
                                          return_xint=True)
        assert np.array_equal(alb, alb0) and np.array_equal(xint, x0)
    opa.close()


@pytest.mark.parametrize("name", ["opt_linear_raman", "opt_nearest_noraman", "opt_linear_clear_nodedd"])
def test_thermal_and_transit_spectrum_one_call(name):
    """pb_spectrum_thermal / pb_spectrum_transit return the bits of compute_opacity(device_outputs=True) ->
    get_thermal_1d(return_thermal=True) / get_transit_1d on the arrays picaso() passes (justdoit.py:337-342, :388-396)"""
    case, g, db, atm, ins = load_case(name)
    opa = device_opacities(pb, case, db, ins)
    a = duck_atmosphere(db, atm)
    L, W = atm["nlayer"], db["nwno"]
    gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = synth.geometry_1d(5, 0.0)
    z = np.linspace(8.0e9, 7.0e9, L + 1)
    dz = np.full(L + 1, (z[0] - z[1]))
    a.level["z"], a.level["dz"] = z, dz
    rng = np.random.default_rng(9)
    for surf, hard in ((None, 0), (0.3 * rng.random(W), 1)):
        opa.get_opacities(a)
        dev = pb.compute_opacity(a, opa, stream=case["stream"], delta_eddington=case["dedd"], test_mode=None,
                                 raman=case["raman"], device_outputs=True)
        DTAU_OG, COSB_OG, W0_no_raman = dev[7][:, :, 0], dev[10][:, :, 0], dev[11][:, :, 0]
        ft0, _, th0 = pb.get_thermal_1d(L + 1, db["wno"], W, 5, 1, atm["tlevel"], DTAU_OG, W0_no_raman, COSB_OG,
                                        atm["plevel"], ubar1, 0 if surf is None else surf, hard, db["wno"] * 0, 0,
                                        level_fluxes=False, gweight=gweight, tweight=tweight, return_thermal=True)
        F0 = pb.get_transit_1d(z, dz, L + 1, W, 6.957e10, atm["mmw"], atm["k_b"], atm["amu"], atm["plevel"], atm["tlevel"],
                               atm["colden"], DTAU_OG)
        opa.get_opacities(a)
        th, ft = pb.thermal_spectrum(a, opa, ubar1, gweight, tweight, surf_reflect=surf, hard_surface=hard,
                                     stream=case["stream"], delta_eddington=case["dedd"], raman=case["raman"], return_flux=True)
        assert np.array_equal(th, th0) and np.array_equal(ft, ft0)
        F = pb.transit_spectrum(a, opa, 6.957e10, stream=case["stream"], delta_eddington=case["dedd"], raman=case["raman"])
        assert np.array_equal(F, F0)
    opa.close()


@pytest.mark.parametrize("dev", [False, True])
@pytest.mark.parametrize("mode", ["rayleigh", "constant_tau"])
@pytest.mark.parametrize("name", ["opt_linear_raman", "opt_nearest_noraman"])
def test_compute_opacity_test_modes(name, mode, dev):
    """compute_opacity(test_mode='rayleigh' | other) (optics.py:372-399) inside the opacity kernel, against the unmodified
    reference (tests/golden/testmode.npz); the caller's cloud w0 array gets the reference's <= 0 -> 1e-10 replacement"""
    from util import golden
    g = golden("testmode")
    case, _, db, atm, ins = load_case(name)
    atm["cloud_w0"][::3, ::5] = 0.0
    opa = device_opacities(pb, dict(case, raman=2), db, ins)
    a = duck_atmosphere(db, atm)
    opa.get_opacities(a)
    res = pb.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"], test_mode=mode,
                             raman=2, device_outputs=dev)
    for n, arr in zip(OUT_NAMES, res):
        got = arr.numpy() if dev else arr[:, :, 0]
        assert_close(got, g[f"{name}/{mode}/{n}"], 1e-12, "%s test_mode=%s %s" % (name, mode, n))
    assert np.array_equal(a.layer["cloud"]["w0"], g[f"{name}/{mode}/cloud_w0_after"])
    opa.close()


def test_pollack_raman_and_full_output(tmp_path, monkeypatch):
    """raman=1 ('pollack', the reference's config default) and full_output=True through the public mirror: the table
    comes from $picaso_refdata/opacities/raman_fortran.txt like the reference's (optics.py:652); golden vectors from
    the unmodified reference (tests/golden/make_golden_pollack.py)"""
    from util import golden
    g = golden("pollack")
    case, _, db, atm, ins = load_case(str(g["case"]))
    (tmp_path / "opacities").mkdir()
    np.savetxt(tmp_path / "opacities" / "raman_fortran.txt", np.column_stack([g["table_w"], g["table_f"]]))
    monkeypatch.setenv("picaso_refdata", str(tmp_path))
    for dev in (False, True):
        opa = device_opacities(pb, dict(case, raman=2), db, ins)
        a = duck_atmosphere(db, atm)
        opa.get_opacities(a)
        res = pb.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"], test_mode=None,
                                 raman=1, full_output=True, device_outputs=dev)
        for n, arr in zip(OUT_NAMES, res):
            got = arr.numpy() if dev else arr[:, :, 0]
            assert_close(got, g["out/" + n], 1e-10, "pollack %s (device_outputs=%s)" % (n, dev))
        for n in ("taugas", "tauray", "taucld"):
            got = getattr(a, n)
            assert got.shape == g["full/" + n].shape + (1,)
            assert_close(got[:, :, 0], g["full/" + n], 1e-10, "full_output " + n)
        opa.close()
    # without the file: the reference's own failure mode
    monkeypatch.setenv("picaso_refdata", str(tmp_path / "nowhere"))
    opa = device_opacities(pb, dict(case, raman=2), db, ins)
    a = duck_atmosphere(db, atm)
    opa.get_opacities(a)
    with pytest.raises(FileNotFoundError):
        pb.compute_opacity(a, opa, stream=2, delta_eddington=True, test_mode=None, raman=1)
    opa.raman_pollack_table = (g["table_w"], g["table_f"])      # or hand the table over directly
    res = pb.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"], test_mode=None, raman=1)
    assert_close(res[2][:, :, 0], g["out/W0"], 1e-10, "pollack via raman_pollack_table")
    opa.close()
