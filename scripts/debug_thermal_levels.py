import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np
import cases as C, oracle, picaso_b200 as pb
d = C.build_thermal(C.thermal_cases()["therm_cfg2_small"])
args = C.thermal_args(d)
ftop, lv = pb.get_thermal_1d(*args)
oftop, olv = oracle.get_thermal_1d(*args)
for k, a, o in zip(("fm", "fp", "fmm", "fpm"), lv, olv):
    colmax = np.max(np.abs(o), axis=-2, keepdims=True)
    err = np.abs(a - o)
    bad = err > 1e-6 * np.abs(o) + 1e-9 * colmax
    print(k, "bad", bad.sum(), "max rel", np.max(err / np.maximum(np.abs(o), 1e-300)))
    idx = np.argwhere(bad)
    for ii in idx[:12]:
        g, t, l, w = ii
        print("   ", ii, "gpu", a[g, t, l, w], "orc", o[g, t, l, w], "colmax", colmax[g, t, 0, w],
              "dtau", d["dtau"][min(l, 89), w], "w0", d["w0"][min(l, 89), w], "cosb", d["cosb"][min(l,89), w], "u", d["ubar1"][g, t])
