"""3-D (per-facet) flux functions: CPU oracle vs reference golden vectors; GPU vs golden + oracle."""
import numpy as np
import pytest

import cases as C
from oracle import facets as of
from util import assert_close, golden


@pytest.mark.parametrize("name", sorted(C.facets_cases()))
def test_oracle_facets(name):
    case = C.facets_cases()[name]
    d = C.build_facets(case)
    fn = of.get_reflected_3d if case["kind"] == "refl" else of.get_thermal_3d
    assert_close(fn(*C.facets_args(d, case)), golden("facets")[name], 1e-10, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(C.facets_cases()))
def test_gpu_facets(name):
    import picaso_b200 as pb
    case = C.facets_cases()[name]
    d = C.build_facets(case)
    args = C.facets_args(d, case)
    fn = pb.get_reflected_3d if case["kind"] == "refl" else pb.get_thermal_3d
    got = fn(*args)
    assert got.shape == (d["ng"], d["nt"], d["nwno"])
    assert_close(got, golden("facets")[name], 1e-6, name + " vs reference")
    if case["kind"] == "refl":
        alb = pb.compress_disco(d["nwno"], d["cos_theta"], got, d["gweight"], d["tweight"], d["F0PI"])
        import oracle
        want = oracle.compress_disco(d["nwno"], d["cos_theta"], golden("facets")[name], d["gweight"], d["tweight"],
                                     d["F0PI"])
        assert_close(alb, want, 1e-6, name + " disk-integrated albedo")
