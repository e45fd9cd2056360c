"""Golden vectors for get_thermal_SH from the unmodified reference (build container only)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref_loader  # noqa: E402
import cases as C  # noqa: E402


def main():
    F = ref_loader.load("fluxes")
    D = ref_loader.load("disco")
    out = {}
    for name, case in C.thermal_sh_cases().items():
        d = C.build_thermal_sh(case)
        x, _ = F.get_thermal_SH(*C.thermal_sh_args(d, case))
        out[name + "/xint"] = x
        out[name + "/thermal"] = D.compress_thermal(d["nwno"], x, d["gweight"], d["tweight"])
    np.savez_compressed(os.path.join(HERE, "thermal_sh.npz"), ref_commit="0369089", **out)
    print("thermal_sh:", len(out), "arrays")


if __name__ == "__main__":
    main()
