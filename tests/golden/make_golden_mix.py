"""Golden vectors for resort-rebin k-mixing from the unmodified reference (deq_chem.py +
RetrieveCKs.mix_my_opacities_gasesfly / get_mixing_indices bound to a bare instance)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref_loader  # noqa: E402
import cases as C  # noqa: E402
import make_golden_optics as MO  # noqa: E402


def main():
    O = ref_loader.load_optics()
    out = {}
    for name, case in C.mix_cases().items():
        db, gases, kappas, atm, gauss_pts, gauss_wts = C.build_mix(case)
        opa = object.__new__(O.RetrieveCKs)
        opa.pressures, opa.temps, opa.nc_p = db["pressures"], db["temps"], db["nc_p"]
        opa.kappas, opa.nwno, opa.ngauss = kappas, db["nwno"], case["K"]
        opa.gauss_pts, opa.gauss_wts = gauss_pts, gauss_wts
        a = MO.duck_atmosphere(dict(db, molecules=gases), atm)
        opa.mix_my_opacities_gasesfly(a)
        out[name + "/molecular_opa"] = opa.molecular_opa
        ind, ti, pi = opa.get_mixing_indices(a)
        out[name + "/indices"] = ind
    np.savez_compressed(os.path.join(HERE, "mix.npz"), ref_commit="0369089", **out)
    print("mix:", len(out), "arrays")


if __name__ == "__main__":
    main()
