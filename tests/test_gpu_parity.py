"""GPU: the CUDA path (through the C ABI, host buffers) against the CPU oracle and the
reference golden vectors.  Tolerance for all TOA / albedo / thermal / transit outputs:
fp64 rtol 1e-6 (BASELINE.json north_star); observed errors are ~1e-11."""
import numpy as np
import pytest

import cases as C
import oracle
import picaso_b200 as pb
from picaso_b200 import synth
from util import assert_close, assert_level_close, assert_level_close_yardstick, golden

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.mark.parametrize("name", sorted(C.reflected_cases()))
def test_reflected_vs_golden_and_oracle(name):
    g = golden("reflected")
    case = C.reflected_cases()[name]
    d = C.build_reflected(case)
    args = C.reflected_args(d, case["kw"])
    xint, lv, alb = pb.get_reflected_1d(*args, gweight=d["gweight"], tweight=d["tweight"],
                                        return_albedo=True)
    assert_close(xint, g[name + "/xint"], RTOL, name + " xint vs reference")
    assert_close(alb, g[name + "/albedo"], RTOL, name + " fused albedo vs reference")
    ox, olv = oracle.get_reflected_1d(*args)
    assert_close(xint, ox, RTOL, name + " xint vs oracle")
    alb2 = pb.compress_disco(d["nwno"], d["cos_theta"], xint, d["gweight"], d["tweight"], d["F0PI"])
    assert_close(alb2, g[name + "/albedo"], RTOL, name + " compress_disco")
    if case["kw"]["get_lvl_flux"]:
        _, qlv = oracle.get_reflected_1d(*args, quad=True, nthreads=8)
        for k, a, o, q in zip(("fm", "fp", "fmm", "fpm"), lv, olv, qlv):
            assert_level_close_yardstick(a, g[name + "/" + k], q, what=name + " " + k + " vs reference")
            assert_level_close_yardstick(a, o, q, what=name + " " + k + " vs oracle")
    else:
        assert all(not a.any() for a in lv)


@pytest.mark.parametrize("mode", ["scan", "fused", "rec"])
@pytest.mark.parametrize("name", sorted(n for n, c in C.reflected_cases().items() if c["kw"]["get_lvl_flux"]))
def test_reflected_level_flux_paths(name, mode, monkeypatch):
    """the three level-flux implementations (PB_REFL_LEVELS): one fused kernel, precomputed layer records +
    serial sweeps, and the layer-parallel warp scans (one warp per column, lanes = layers)"""
    monkeypatch.setenv("PB_REFL_LEVELS", mode)
    g = golden("reflected")
    case = C.reflected_cases()[name]
    d = C.build_reflected(case)
    args = C.reflected_args(d, case["kw"])
    xint, lv = pb.get_reflected_1d(*args)
    _, olv = oracle.get_reflected_1d(*args)
    _, qlv = oracle.get_reflected_1d(*args, quad=True, nthreads=8)
    for k, a, o, q in zip(("fm", "fp", "fmm", "fpm"), lv, olv, qlv):
        assert np.isfinite(a).all(), name + " " + k
        assert_level_close_yardstick(a, g[name + "/" + k], q, what=name + " " + k + " vs reference (" + mode + ")")
        assert_level_close_yardstick(a, o, q, what=name + " " + k + " vs oracle (" + mode + ")")


@pytest.mark.parametrize("L", [33, 45, 90, 120])
def test_reflected_level_scan_many_layers(L, monkeypatch):
    """more layers than lanes: 2, 3 and 4 layers per lane in the layer-parallel scan kernel"""
    monkeypatch.setenv("PB_REFL_LEVELS", "scan")
    d = synth.reflected_inputs(L=L, W=19, seed=900 + L)
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=1)
    args = C.reflected_args(d, kw)
    _, lv = pb.get_reflected_1d(*args)
    _, olv = oracle.get_reflected_1d(*args, nthreads=8)
    _, qlv = oracle.get_reflected_1d(*args, quad=True, nthreads=8)
    for k, a, o, q in zip(("fm", "fp", "fmm", "fpm"), lv, olv, qlv):
        assert np.isfinite(a).all(), k
        assert_level_close_yardstick(a, o, q, what="L=%d %s scan vs oracle" % (L, k))


@pytest.mark.parametrize("name", sorted(C.thermal_cases()))
def test_thermal_vs_golden_and_oracle(name):
    g = golden("thermal")
    d = C.build_thermal(C.thermal_cases()[name])
    args = C.thermal_args(d)
    ftop, lv = pb.get_thermal_1d(*args)
    assert_close(ftop, g[name + "/ftop"], RTOL, name + " ftop vs reference")
    f2, none, th = pb.get_thermal_1d(*args, level_fluxes=False, gweight=d["gweight"],
                                     tweight=d["tweight"], return_thermal=True)
    assert none is None
    assert_close(f2, g[name + "/ftop"], RTOL, name + " single-sweep ftop vs reference")
    assert_close(th, g[name + "/thermal"], RTOL, name + " fused thermal vs reference")
    th2 = pb.compress_thermal(d["nwno"], ftop, d["gweight"], d["tweight"])
    assert_close(th2, g[name + "/thermal"], RTOL, name + " compress_thermal")
    oftop, olv = oracle.get_thermal_1d(*args)
    assert_close(ftop, oftop, RTOL, name + " vs oracle")
    qftop, qlv = oracle.get_thermal_1d(*args, quad=True, nthreads=8)
    assert_close(ftop, qftop, RTOL, name + " vs binary128 yardstick")
    for k, a, o, q in zip(("fm", "fp", "fmm", "fpm"), lv, olv, qlv):
        assert_level_close_yardstick(a, o, q, what=name + " " + k + " vs oracle/yardstick")
        if name + "/" + k in g.files:
            assert_level_close_yardstick(a, g[name + "/" + k], q, what=name + " " + k + " vs reference")
    lvc = pb.compress_thermal(d["nwno"], lv[1], d["gweight"], d["tweight"])
    assert lvc.shape == (d["nlevel"], d["nwno"])
    assert_level_close(lvc, oracle.compress_thermal(d["nwno"], olv[1], d["gweight"], d["tweight"]),
                       what="compress_thermal 4-D")


@pytest.mark.parametrize("kernel", ["chain", "angle", "wave"])
@pytest.mark.parametrize("name", sorted(C.thermal_cases()))
def test_thermal_toa_kernels(name, kernel, monkeypatch):
    """every TOA kernel of get_thermal_1d (PB_THERM_KERNEL): the chain-warp kernel (elimination once per wavelength on an
    extra warp), the angle-parallel kernel and the one-thread-per-wavelength kernel against the reference"""
    g = golden("thermal")
    d = C.build_thermal(C.thermal_cases()[name])
    monkeypatch.setenv("PB_THERM_KERNEL", kernel)
    ftop, none, th = pb.get_thermal_1d(*C.thermal_args(d), level_fluxes=False, gweight=d["gweight"], tweight=d["tweight"],
                                       return_thermal=True)
    assert none is None
    assert_close(ftop, g[name + "/ftop"], RTOL, "%s ftop (%s kernel) vs reference" % (name, kernel))
    assert_close(th, g[name + "/thermal"], RTOL, "%s fused thermal (%s kernel) vs reference" % (name, kernel))
    monkeypatch.setenv("PB_THERM_WT", "17")    # narrow tiles: partial warps, several chunks per warp
    f2, _ = pb.get_thermal_1d(*C.thermal_args(d), level_fluxes=False)
    assert_close(f2, g[name + "/ftop"], RTOL, "%s ftop (%s kernel, 17-wide tiles) vs reference" % (name, kernel))


@pytest.mark.parametrize("name", sorted(C.sh_cases()))
def test_reflected_sh_vs_golden_and_oracle(name):
    g = golden("sh")
    case = C.sh_cases()[name]
    d = C.build_sh(case)
    a = C.sh_args(d, case)
    xint, flux, alb = pb.get_reflected_SH(*a, gweight=d["gweight"], tweight=d["tweight"], return_albedo=True)
    assert_close(xint, g[name + "/xint"], RTOL, name + " xint vs reference")
    assert_close(alb, g[name + "/albedo"], RTOL, name + " fused albedo vs reference")
    assert_close(a[10], g[name + "/f_deltaM_after"], 1e-13, name + " f_deltaM side effect")
    assert flux.shape == (d["numg"], d["numt"], case["stream"] * d["nlevel"], d["nwno"]) and not flux.any()
    ox, _ = oracle.get_reflected_SH(*C.sh_args(d, case))
    assert_close(xint, ox, RTOL, name + " xint vs oracle")


@pytest.mark.parametrize("name", sorted(C.sh_flux_cases()))
def test_reflected_sh_layer_fluxes(name):
    """get_reflected_SH(flx=1) (calculate_fluxes='on', fluxes.py:2889-2890): flux = F.X + G at every level against
    the unmodified reference (tests/golden/sh_flux.npz) and the oracle.  The kernel substitutes X from the pivot
    rows of its windowed elimination where the reference calls LAPACK; level fluxes are judged by the level-flux
    criterion (rtol 1e-6 + 1e-9 of the column maximum; yardstick slack only where the fp64 reference itself is
    off its binary128 evaluation).  xint_at_top of the same call must not change."""
    g = golden("sh_flux")
    case = C.sh_flux_cases()[name]
    d = C.build_sh(case)
    xint, flux, alb = pb.get_reflected_SH(*C.sh_args(d, case, flx=1), gweight=d["gweight"], tweight=d["tweight"],
                                          return_albedo=True)
    assert flux.shape == g[name + "/flux"].shape
    assert_close(xint, g[name + "/xint"], RTOL, name + " xint (flx=1) vs reference")
    x0, _, alb0 = pb.get_reflected_SH(*C.sh_args(d, case), gweight=d["gweight"], tweight=d["tweight"], return_albedo=True)
    # same algorithm, separately compiled (FMA contraction differs): near-resonant columns move by ~1e-12
    assert_close(xint, x0, 1e-9, name + " xint flx=1 vs flx=0")
    assert_close(alb, alb0, 1e-9, name + " albedo flx=1 vs flx=0")
    _, oflux = oracle.get_reflected_SH(*C.sh_args(d, case, flx=1))
    _, exact = oracle.get_reflected_SH(*C.sh_args(d, case, flx=1), quad=True)
    assert_level_close_yardstick(flux, g[name + "/flux"], exact, what=name + " flux vs reference")
    assert_level_close_yardstick(flux, oflux, exact, what=name + " flux vs oracle")


def test_reflected_sh_layer_fluxes_chunked(monkeypatch):
    """the pivot-row scratch is sized per wavelength chunk: several chunks give the same bits as one"""
    case = C.sh_flux_cases()["sh4_cfg3_tthg"]
    d = C.build_sh(case)
    x1, f1 = pb.get_reflected_SH(*C.sh_args(d, case, flx=1))
    monkeypatch.setenv("PB_SH_FLX_WCAP", "32")
    x2, f2 = pb.get_reflected_SH(*C.sh_args(d, case, flx=1))
    assert np.array_equal(x1, x2) and np.array_equal(f1, f2)


@pytest.mark.parametrize("name", sorted(C.transit_cases()))
def test_transit_vs_golden_and_oracle(name):
    g = golden("transit")
    d = synth.transit_inputs(**C.transit_cases()[name])
    F = pb.get_transit_1d(*C.transit_args(d))
    assert_close(F, g[name + "/F"], RTOL, name + " vs reference")
    assert_close(F, oracle.get_transit_1d(*C.transit_args(d)), RTOL, name + " vs oracle")


def test_upload_mirror_tracks_changing_geometry():
    """The small host vectors of a call (geometry, weights) are not copied again when the pinned slot's device block
    already holds the same bytes (pb_upload_flush).  Alternating geometries through more calls than the ring has slots
    must give, call for call, the results of a context that always copies."""
    d = synth.reflected_inputs(L=7, W=45, seed=77)
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
    geos = [synth.geometry_1d(5, 0.0), synth.geometry_1d(6, 0.7), synth.geometry_1d(5, 0.3)]
    ctx = pb.Context(0)
    want = []
    for gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta in geos:
        dd = dict(d, ubar0=ubar0, ubar1=ubar1, cos_theta=cos_theta, numg=ubar0.shape[0], numt=ubar0.shape[1])
        want.append(pb.get_reflected_1d(*C.reflected_args(dd, kw), gweight=gweight, tweight=tweight, return_albedo=True))
    for i in range(40):     # ring of 8 slots: every slot sees every geometry, in changing order
        k = (i * 7 + i // 3) % 3
        gangle, gweight, tangle, tweight, ubar0, ubar1, cos_theta = geos[k]
        dd = dict(d, ubar0=ubar0, ubar1=ubar1, cos_theta=cos_theta, numg=ubar0.shape[0], numt=ubar0.shape[1])
        x, _, alb = pb.get_reflected_1d(*C.reflected_args(dd, kw), gweight=gweight, tweight=tweight, return_albedo=True,
                                        ctx=ctx)
        assert np.array_equal(x, want[k][0]) and np.array_equal(alb, want[k][2]), (i, k)
    ctx.close()


def test_empty_wave_axis():
    d = synth.reflected_inputs(L=5, W=0, seed=1)
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
    xint, lv = pb.get_reflected_1d(*C.reflected_args(d, kw))
    assert xint.shape == (5, 1, 0)


def test_strided_ck_slice():
    """X[:, :, ig] views of [L, W, K] arrays (ngauss>1 path of picaso(), justdoit.py:256-283)."""
    d = synth.reflected_inputs(L=9, W=41, seed=21)
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
    ref, _ = oracle.get_reflected_1d(*C.reflected_args(d, kw))
    d2 = dict(d)
    for k in ("dtau", "tau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og", "tau_og",
              "w0_og", "cosb_og"):
        big = np.random.default_rng(0).random(d[k].shape + (3,))
        big[:, :, 1] = d[k]
        d2[k] = big[:, :, 1]
    got, _ = pb.get_reflected_1d(*C.reflected_args(d2, kw))
    assert_close(got, ref, RTOL, "strided")


def test_row_padded_leading_dimension():
    d = synth.thermal_inputs(L=11, W=50, seed=4)
    ref, _ = oracle.get_thermal_1d(*C.thermal_args(dict(d, calc_type=0)))
    d2 = dict(d, calc_type=0)
    for k in ("dtau", "w0", "cosb"):
        big = np.zeros((11, 64))
        big[:, :50] = d[k]
        d2[k] = big[:, :50]
    got, _ = pb.get_thermal_1d(*C.thermal_args(d2), level_fluxes=False)
    assert_close(got, ref, RTOL, "padded ld")


def test_many_angles_3d_geometry():
    """ng x nt = 6 x 4 > 8 angles: un-fused disk integration path."""
    gangle, gweight, tangle, tweight = pb.get_angles_3d(6, 4)
    ubar0, ubar1, cos_theta, lat, lon = pb.compute_disco(6, 4, gangle, tangle, 0.4)
    d = synth.reflected_inputs(L=8, W=70, seed=31)
    d.update(numg=6, numt=4, ubar0=ubar0, ubar1=ubar1, cos_theta=float(cos_theta), gweight=gweight,
             tweight=tweight)
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
    args = C.reflected_args(d, kw)
    xint, _, alb = pb.get_reflected_1d(*args, gweight=gweight, tweight=tweight, return_albedo=True)
    ox, _ = oracle.get_reflected_1d(*args)
    assert_close(xint, ox, RTOL, "3d xint")
    assert_close(alb, oracle.compress_disco(70, float(cos_theta), ox, gweight, tweight, d["F0PI"]),
                 RTOL, "3d albedo")


def test_full_size_properties():
    """BASELINE headline size (60 x 10000 x 5): oracle comparison on a wavelength sample plus
    size-independent properties - linearity in F0PI and invariance to wavelength permutation."""
    d = synth.reflected_inputs(L=60, W=10000, seed=1000)
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
    xint, _, alb = pb.get_reflected_1d(*C.reflected_args(d, kw), gweight=d["gweight"],
                                       tweight=d["tweight"], return_albedo=True)
    assert np.isfinite(xint).all()
    ox, _ = oracle.get_reflected_1d(*C.reflected_args(d, kw), nthreads=8)
    assert_close(xint, ox, RTOL, "headline xint")
    # scaling the stellar flux by a power of two scales every intermediate exactly
    d2 = dict(d, F0PI=d["F0PI"] * 2.0)
    x2, _, alb2 = pb.get_reflected_1d(*C.reflected_args(d2, kw), gweight=d["gweight"],
                                      tweight=d["tweight"], return_albedo=True)
    assert np.array_equal(x2, 2.0 * xint), "linearity in F0PI"
    assert np.array_equal(alb2, alb), "albedo independent of F0PI"
    perm = np.random.default_rng(0).permutation(d["nwno"])
    dp = dict(d)
    for k, v in d.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[-1] == d["nwno"]:
            dp[k] = np.ascontiguousarray(v[..., perm])
    xp, _ = pb.get_reflected_1d(*C.reflected_args(dp, kw))
    assert np.array_equal(xp, xint[..., perm]), "wavelengths are not independent"


def test_tau_not_cumsum_of_dtau():
    """tau / tau_og are independent arguments of the reference signature.  The top-down kernel
    (toon_reflected_toa4.cuh) replaces exp(-tau/u0) by running products only where
    tau[l+1] == tau[l] + dtau[l]; everywhere else it must follow the caller's arrays exactly."""
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
    rng = np.random.default_rng(5)
    for phase in (0.0, 0.7):
        d = synth.reflected_inputs(L=23, W=97, seed=61, phase=phase)
        d = dict(d)
        d["tau"] = d["tau"] * 1.07 + 0.01                      # inconsistent everywhere, tau[0] != 0
        tau_og = d["tau_og"].copy()
        rows = rng.integers(1, 24, size=40)
        cols = rng.integers(0, 97, size=40)
        tau_og[rows, cols] *= 1.0 + 1e-9                        # inconsistent at scattered levels only
        d["tau_og"] = tau_og
        args = C.reflected_args(d, kw)
        xint, _ = pb.get_reflected_1d(*args)
        ox, _ = oracle.get_reflected_1d(*args)
        assert_close(xint, ox, RTOL, "inconsistent tau, phase %.1f" % phase)


def test_chunked_host_pipeline_ragged_padded(monkeypatch):
    """Opt-in PB_REFL_CHUNKS: PB_HOST calls split into wavelength chunks (copy stream / compute
    stream, toon_reflected.cu); ragged last chunk, padded leading dimension, xint-only and albedo."""
    monkeypatch.setenv("PB_REFL_CHUNKS", "4")
    W = 4099
    d = synth.reflected_inputs(L=12, W=W, seed=77)
    kw = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)
    ox, _ = oracle.get_reflected_1d(*C.reflected_args(d, kw), nthreads=8)
    oalb = oracle.compress_disco(W, d["cos_theta"], ox, d["gweight"], d["tweight"], d["F0PI"])
    d2 = dict(d)
    for k in ("dtau", "tau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og", "tau_og",
              "w0_og", "cosb_og"):
        big = np.full((d[k].shape[0], W + 61), np.nan)
        big[:, :W] = d[k]
        d2[k] = big[:, :W]
    for dd in (d, d2):
        xint, _, alb = pb.get_reflected_1d(*C.reflected_args(dd, kw), gweight=d["gweight"],
                                           tweight=d["tweight"], return_albedo=True)
        assert_close(xint, ox, RTOL, "chunked xint")
        assert_close(alb, oalb, RTOL, "chunked albedo")
    x_only, _ = pb.get_reflected_1d(*C.reflected_args(d2, kw))
    assert np.array_equal(x_only, xint)


def test_table_exp():
    """shared-memory table exp of the v5 reflected kernel (pb_math.cuh: exp_tab) vs libm on its domain x <= 709:
    relative error <= 1.2e-16 |x| + 1 ulp; x <= -708 comes out <= 2.3e-308 (never inf / NaN); NaN propagates."""
    from picaso_b200 import _lib
    ctx = _lib.default_context()
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-707, 700, 300000), rng.uniform(-40, 40, 300000), rng.normal(0, 1e-3, 1000),
                        -10.0 ** rng.uniform(-300, 300, 50000), 10.0 ** rng.uniform(-300, 2, 20000),
                        [0.0, -0.0, -np.inf, np.nan, 35.0, -35.0, 709.0, -745.1, -1023.9, -1024.0, -1e6, -2.4e7, -1e15,
                         -1e100, -1e300, -1e-310, 1e-310]])
    x = x[np.isnan(x) | (x <= 709.0)]
    e = np.empty_like(x)
    ctx.check(ctx.lib.pb_selftest_exp_tab(ctx.h, x.ctypes.data, x.size, e.ctypes.data))
    with np.errstate(all="ignore"):
        we = np.exp(x)
    assert np.array_equal(np.isnan(e), np.isnan(x))
    reg = (x > -707.0) & ~np.isnan(x)
    err = np.abs(e[reg] - we[reg]) / we[reg]
    assert np.all(err <= 1.2e-16 * np.abs(x[reg]) + 2.3e-16), err.max()
    assert np.max(err[np.abs(x[reg]) <= 40]) < 4e-15
    low = (x <= -707.0)
    assert np.all((e[low] >= 0.0) & (e[low] < 1e-306))
    assert e[np.flatnonzero(x == 0.0)[0]] == 1.0


def test_kernel_math_primitives():
    """branch-free exp / reciprocal used inside the kernels vs numpy (libm)."""
    from picaso_b200 import _lib
    ctx = _lib.default_context()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-745, 709, 200000), rng.uniform(-40, 40, 200000),
                        rng.normal(0, 1e-3, 1000), 10.0 ** rng.uniform(-300, 300, 50000),
                        -10.0 ** rng.uniform(-300, 300, 50000),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 35.0, -35.0, 709.7, -745.1, -800.0, 800.0,
                         1e-310, -1e-310]])
    e = np.empty_like(x)
    r = np.empty_like(x)
    ctx.check(ctx.lib.pb_selftest_math(ctx.h, x.ctypes.data, x.size, e.ctypes.data, r.ctypes.data))
    with np.errstate(all="ignore"):
        we, wr = np.exp(x), 1.0 / x
    # documented deviations of pbm::exp from libm: results < 2^-1021 flush to 0 (x <= -708),
    # saturation to +inf starts at x = 709.4 instead of 709.78
    window = (x >= 709.4) & (x < 709.79)
    fin = np.isfinite(we) & (x > -708.0) & ~window
    assert np.max(np.abs(e[fin] - we[fin]) / we[fin]) < 5e-16
    low = (x <= -708.0)
    assert np.all((e[low] == 0.0) | (np.abs(e[low] - we[low]) <= 5e-16 * we[low])) and np.all(we[low] < 3.4e-308)
    assert np.all(np.isinf(e[window]) | (np.abs(e[window] - we[window]) <= 5e-16 * we[window]))
    assert np.array_equal(np.isnan(e), np.isnan(we))
    assert np.array_equal(np.isinf(e[~window]), np.isinf(we[~window]))
    reg = np.isfinite(x) & (np.abs(x) >= 2.3e-308) & (np.abs(x) < 1e307)
    assert np.max(np.abs(r[reg] - wr[reg]) / np.abs(wr[reg])) < 4e-16
    assert np.array_equal(np.isnan(r), np.isnan(wr))
    for v in (0.0, -0.0, np.inf, -np.inf):
        i = np.flatnonzero((x == v) & (np.signbit(x) == np.signbit(v)))[0]
        assert r[i] == wr[i] and np.signbit(r[i]) == np.signbit(wr[i])
