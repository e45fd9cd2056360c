"""climate.get_fluxes: oracle pinned to the reference's own get_fluxes
(tests/golden/climate.npz), CUDA path (pb_climate_get_fluxes) against both.

Tolerances.  [nlevel] net fluxes are differences of wavelength sums: rtol 1e-6 of the entry plus 1e-9 of
the largest |net flux| of the column (deep levels cancel to ~0).  [nlevel, nwno] level arrays: the mixed
level-flux criterion of tests/util.py (rtol 1e-6 + 1e-9 of the column maximum).  The reference's level-flux
formulas are ill-conditioned in optically thick layers: where the fp64 reference
itself is further than that from the binary128 evaluation of its own formulas, an independent fp64
implementation is held to 32 x (level arrays) / 4 x (net fluxes) the reference's own error instead (the
yardstick criterion; the factors follow from the measured need, profiles/r2_yardstick.jsonl)."""
import numpy as np
import pytest

import cases as C
import oracle
from oracle import climate as oclim
from util import assert_level_close, assert_level_close_yardstick, golden, log_yardstick

NET_SLACK = 4.0    # measured need <= 0.95 (profiles/r2_yardstick.jsonl); 64 in round 1


def check_net(a, b, x, what):
    """[nlevel] net-flux vector `a` against the fp64 reference `b` and its binary128 evaluation `x`.
    (i) yardstick: 1e-6 |x| + 1e-9 max|x| + NET_SLACK x the fp64 reference's own largest error;
    (ii) plain rtol 1e-6 against the fp64 reference at every level where the reference itself is well conditioned
    (within 1e-8 of its binary128 evaluation): there an independent fp64 implementation has no excuse."""
    scale = np.max(np.abs(x)) if x.size else 0.0
    ref_err = np.max(np.abs(b - x)) if x.size else 0.0
    err = np.abs(a - x)
    plain = 1e-6 * np.abs(x) + 1e-9 * scale
    bad = err > plain + NET_SLACK * ref_err
    excess = err - plain
    need = float(np.max(excess) / ref_err) if (x.size and ref_err > 0 and np.max(excess) > 0) else 0.0
    log_yardstick(what, need, int((excess > 0).sum()), int(a.size), NET_SLACK)
    assert not bad.any(), "%s: %d entries off, max abs err %.3e (scale %.3e, fp64 reference err %.3e)" % (
        what, int(bad.sum()), float(np.max(err)), scale, ref_err)
    well = np.abs(b - x) <= 1e-8 * np.abs(x)
    off = well & (np.abs(a - b) > 1e-6 * np.abs(b))
    assert not off.any(), "%s: %d well-conditioned levels differ from the fp64 reference by more than rtol 1e-6" % (
        what, int(off.sum()))


def check(got, ref, what, exact=None):
    """exact: binary128-based evaluation; then `ref` is the fp64 evaluation of the same formulas and the
    comparison is the yardstick one."""
    for i, (k, a, b) in enumerate(zip(C.CLIMATE_OUT, got, ref)):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        if exact is not None:
            x = np.asarray(exact[i])
            if k.startswith("flux_net"):
                check_net(a, b, x, what + " " + k)
            else:
                assert_level_close_yardstick(a, b, x, what=what + " " + k)
            continue
        if k.startswith("flux_net"):
            scale = np.max(np.abs(b)) if b.size else 0.0
            bad = np.abs(a - b) > 1e-6 * np.abs(b) + 1e-9 * scale
            assert not bad.any(), "%s %s: %d entries off, max abs err %.3e (scale %.3e)" % (
                what, k, int(bad.sum()), float(np.max(np.abs(a - b))), scale)
        else:
            assert_level_close(a, b, what=what + " " + k)


@pytest.mark.parametrize("name", sorted(C.climate_cases()))
def test_oracle_vs_reference(name):
    case = C.climate_cases()[name]
    d = C.build_climate(case)
    g = golden("climate")
    got = oclim.get_fluxes(*C.climate_args(d, case), nthreads=4)
    check(got, [g[name + "/" + k] for k in C.CLIMATE_OUT], name + " oracle vs reference")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(C.climate_cases()))
def test_gpu_vs_reference_and_oracle(name):
    import picaso_b200 as pb
    case = C.climate_cases()[name]
    d = C.build_climate(case)
    g = golden("climate")
    got = pb.get_fluxes(*C.climate_args(d, case))
    exact = oclim.get_fluxes(*C.climate_args(d, case), nthreads=8, quad=True)
    check(got, [g[name + "/" + k] for k in C.CLIMATE_OUT], name + " gpu vs reference", exact=exact)
    check(got, oclim.get_fluxes(*C.climate_args(d, case), nthreads=4), name + " gpu vs oracle", exact=exact)


@pytest.mark.gpu
def test_gpu_device_resident_opacities():
    """DeviceArray inputs (what compute_opacity(device_outputs=True) returns) == numpy inputs, bit for bit"""
    import picaso_b200 as pb
    from picaso_b200.optics import DeviceArray
    case = C.climate_cases()["clim_k4"]
    d = C.build_climate(case)
    ref = pb.get_fluxes(*C.climate_args(d, case))
    ctx = pb.default_context()
    dd = dict(d)
    dd["OpacityWEd"] = type(d["OpacityWEd"])(*[DeviceArray.from_numpy(ctx, a) for a in d["OpacityWEd"]])
    dd["OpacityNoEd"] = type(d["OpacityNoEd"])(*[DeviceArray.from_numpy(ctx, a) for a in d["OpacityNoEd"]])
    got = pb.get_fluxes(*C.climate_args(dd, case))
    for k, a, b in zip(C.CLIMATE_OUT, got, ref):
        assert np.array_equal(a, b), k


@pytest.mark.gpu
def test_gpu_net_fluxes_only():
    """full_arrays=False: the [nlevel, nwno] arrays are not copied back; the net-flux vectors are the same bits"""
    import picaso_b200 as pb
    case = C.climate_cases()["clim_k8_disk5"]
    d = C.build_climate(case)
    full = pb.get_fluxes(*C.climate_args(d, case))
    net = pb.get_fluxes(*C.climate_args(d, case), full_arrays=False)
    for i, k in enumerate(C.CLIMATE_OUT):
        if k.startswith("flux_net"):
            assert np.array_equal(net[i], full[i]), k
        else:
            assert net[i] is None


def test_bound_call_is_callable_from_numba_nopython():
    """CPU: pb_climate_run_bound takes two integers, so numba's ctypes support can call it from jitted code (the
    reference's solver loop is numba-nopython, climate.py:804).  Without a context it answers PB_ERR_ARG."""
    import numba
    from picaso_b200 import _lib
    run = _lib.load_library().pb_climate_run_bound

    @numba.njit
    def call(addr, handle):
        return run(addr, handle)

    assert call(0, 0) == 2      # PB_ERR_ARG: no such binding - but the call itself went through nopython code


def _perturbed_profiles(t_level):
    """the Jacobian loop of t_start (climate.py:1108-1180): one level temperature perturbed per profile"""
    V = t_level.size
    tl = np.tile(t_level, (V + 1, 1))
    for j in range(V):
        tl[j + 1, j] += 0.01 * t_level[j]
    return tl


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["clim_k4", "clim_k8_disk5"])
def test_gpu_jacobian_batch(name):
    """get_fluxes_jacobian: all perturbed temperature profiles in one call == one thermal get_fluxes per profile"""
    import picaso_b200 as pb
    case = C.climate_cases()[name]
    d = C.build_climate(case)
    args = list(C.climate_args(d, case))
    atm = args[0]
    tls = _perturbed_profiles(np.asarray(atm.t_level, dtype=np.float64))
    lay, net = pb.get_fluxes_jacobian(*args[:6], tls)
    assert lay.shape == net.shape == tls.shape
    for p in (0, 1, tls.shape[0] // 2, tls.shape[0] - 1):
        a2 = list(args)
        a2[0] = atm._replace(t_level=tls[p])
        a2[7], a2[8] = False, True
        one = pb.get_fluxes(*a2, full_arrays=False)
        # same kernels, same per-column arithmetic (the batch size may select the fused or the record-based level kernel:
        # identical expressions, so allow rounding-level slack only)
        assert np.allclose(lay[p], one[4], rtol=1e-12, atol=1e-12 * np.max(np.abs(one[4]))), (name, p)
        assert np.allclose(net[p], one[5], rtol=1e-12, atol=1e-12 * np.max(np.abs(one[5]))), (name, p)
    # and against the oracle with the net-flux criterion
    a2 = list(args)
    a2[0] = atm._replace(t_level=tls[3])
    a2[7], a2[8] = False, True
    ref = oclim.get_fluxes(*a2, nthreads=4)
    exact = oclim.get_fluxes(*a2, nthreads=8, quad=True)
    check_net(net[3], np.asarray(ref[5]), np.asarray(exact[5]), name + " jacobian profile 3 flux_net_ir")
    check_net(lay[3], np.asarray(ref[4]), np.asarray(exact[4]), name + " jacobian profile 3 flux_net_ir_layer")


@pytest.mark.gpu
def test_gpu_bound_call_from_numba():
    """a bound thermal get_fluxes driven from a numba-jitted loop that rewrites the temperatures in place"""
    import numba
    import picaso_b200 as pb
    case = C.climate_cases()["clim_k4"]
    d = C.build_climate(case)
    args = list(C.climate_args(d, case))
    atm = args[0]
    bf = pb.BoundFluxes(*args[:7], reflected=False, thermal=True)
    run, addr, h = bf.run, bf.ctx_address, bf.handle

    @numba.njit
    def sweep(t_level, net_ir, out):
        for j in range(out.shape[0]):
            keep = t_level[j]
            t_level[j] = keep * 1.01
            rc = run(addr, h)
            if rc != 0:
                return rc
            out[j, :] = net_ir
            t_level[j] = keep
        return 0

    V = bf.t_level.size
    out = np.zeros((3, V))
    assert sweep(bf.t_level, bf.flux_net_ir, out) == 0
    tls = _perturbed_profiles(np.asarray(atm.t_level, dtype=np.float64))
    _, net = pb.get_fluxes_jacobian(*args[:6], tls)
    assert np.allclose(out, net[1:4], rtol=1e-12, atol=1e-12 * np.max(np.abs(net)))
    bf.close()
