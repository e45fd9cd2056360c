"""climate.get_fluxes: oracle pinned to the reference's own get_fluxes
(tests/golden/climate.npz), CUDA path (pb_climate_get_fluxes) against both.

Tolerances.  [nlevel] net fluxes are differences of wavelength sums: rtol 1e-6 of the entry plus 1e-9 of
the largest |net flux| of the column (deep levels cancel to ~0).  [nlevel, nwno] level arrays: the mixed
level-flux criterion of tests/util.py (rtol 1e-6 + 1e-9 of the column maximum).  The reference's level-flux
formulas are ill-conditioned in optically thick layers: where the fp64 reference
itself is further than that from the binary128 evaluation of its own formulas, an independent fp64
implementation is held to 64 x the reference's own error instead (the yardstick criterion)."""
import numpy as np
import pytest

import cases as C
import oracle
from oracle import climate as oclim
from util import assert_level_close, assert_level_close_yardstick, golden


def check(got, ref, what, exact=None):
    """exact: binary128-based evaluation; then `ref` is the fp64 evaluation of the same formulas and the
    comparison is the yardstick one."""
    for i, (k, a, b) in enumerate(zip(C.CLIMATE_OUT, got, ref)):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        if exact is not None:
            x = np.asarray(exact[i])
            if k.startswith("flux_net"):
                scale = np.max(np.abs(x)) if x.size else 0.0
                ref_err = np.max(np.abs(b - x)) if x.size else 0.0
                bad = np.abs(a - x) > 1e-6 * np.abs(x) + 1e-9 * scale + 64.0 * ref_err
                assert not bad.any(), "%s %s: %d entries off, max abs err %.3e (scale %.3e, fp64 reference err %.3e)" % (
                    what, k, int(bad.sum()), float(np.max(np.abs(a - x))), scale, ref_err)
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
