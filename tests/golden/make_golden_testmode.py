"""Golden vectors for compute_opacity(test_mode='rayleigh' | 'constant_tau') from the UNMODIFIED reference
(optics.py:372-399).  Build container only:

    python tests/golden/make_golden_testmode.py   ->  tests/golden/testmode.npz
"""
import os
import sqlite3
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import cases as C  # noqa: E402
import make_golden_optics as MG  # noqa: E402
from oracle import ref_loader  # noqa: E402
from picaso_b200 import synth  # noqa: E402

CASES = {"opt_linear_raman": 2, "opt_nearest_noraman": 2}   # optics case -> raman option used here
NAMES = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "DTAU_OG", "TAU_OG", "W0_OG", "COSB_OG",
         "W0_no_raman", "f_deltaM")


def main():
    O = ref_loader.load_optics()
    sqlite3.register_adapter(np.int64, int)
    sqlite3.register_adapter(np.int32, int)
    raman_txt = os.path.join(ref_loader.REF_ROOT, "reference", "opacities", "raman.txt")
    out = {}
    for cname, raman in CASES.items():
        case = C.optics_cases()[cname]
        db = synth.opacity_database(**case["db"])
        for mode in ("rayleigh", "constant_tau"):
            atm = synth.atmosphere_profile(db, **case["atm"])
            atm["cloud_w0"][::3, ::5] = 0.0        # exercise the w0 <= 0 -> 1e-10 replacement (optics.py:393)
            with tempfile.TemporaryDirectory() as tmp:
                path = os.path.join(tmp, "opa.db")
                MG.write_db(path, db)
                opa = O.RetrieveOpacities(path, raman_txt, query_method=case["query"])
                a = MG.duck_atmosphere(db, atm)
                opa.get_opacities(a)
                res = O.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"],
                                        test_mode=mode, raman=raman)
            for n, arr in zip(NAMES, res):
                arr = np.asarray(arr)
                out[f"{cname}/{mode}/{n}"] = arr[:, :, 0] if arr.ndim == 3 else arr
            out[f"{cname}/{mode}/cloud_w0_after"] = np.asarray(a.layer["cloud"]["w0"])
    np.savez_compressed(os.path.join(HERE, "testmode.npz"), ref_commit=MG.REF_COMMIT, **out)
    print("testmode:", len(out), "arrays")


if __name__ == "__main__":
    main()
