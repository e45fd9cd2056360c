"""Golden vectors for get_reflected_3d / get_thermal_3d from the unmodified reference."""
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
    out = {}
    for name, case in C.facets_cases().items():
        d = C.build_facets(case)
        fn = F.get_reflected_3d if case["kind"] == "refl" else F.get_thermal_3d
        out[name] = fn(*C.facets_args(d, case))
    np.savez_compressed(os.path.join(HERE, "facets.npz"), ref_commit="0369089", **out)
    print("facets:", len(out), "arrays")


if __name__ == "__main__":
    main()
