"""get_thermal_SH: CPU oracle vs reference golden vectors; GPU vs golden + oracle."""
import numpy as np
import pytest

import cases as C
import oracle
from util import assert_close, golden


@pytest.mark.parametrize("name", sorted(C.thermal_sh_cases()))
def test_oracle_thermal_sh(name):
    g = golden("thermal_sh")
    case = C.thermal_sh_cases()[name]
    d = C.build_thermal_sh(case)
    x, _ = oracle.get_thermal_SH(*C.thermal_sh_args(d, case))
    assert_close(x, g[name + "/xint"], 1e-9, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(C.thermal_sh_cases()))
def test_gpu_thermal_sh(name):
    import picaso_b200 as pb
    g = golden("thermal_sh")
    case = C.thermal_sh_cases()[name]
    d = C.build_thermal_sh(case)
    args = C.thermal_sh_args(d, case)
    x, flux, th = pb.get_thermal_SH(*args, gweight=d["gweight"], tweight=d["tweight"], return_thermal=True)
    assert_close(x, g[name + "/xint"], 1e-6, name + " xint vs reference")
    assert_close(th, g[name + "/thermal"], 1e-6, name + " fused thermal vs reference")
    ox, _ = oracle.get_thermal_SH(*args)
    assert_close(x, ox, 1e-6, name + " vs oracle")
    assert flux.shape == (d["numg"], d["numt"], case["stream"] * d["nlevel"], d["nwno"]) and not flux.any()
    if case["same"]:
        # passing the very same object for cosb and cosb_og takes the aliasing shortcut
        a2 = list(args)
        a2[9] = a2[14]
        x2, _ = pb.get_thermal_SH(*a2)
        assert np.array_equal(x, x2)
