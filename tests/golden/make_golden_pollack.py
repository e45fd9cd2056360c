"""Golden vectors for compute_opacity(raman=1 'pollack', full_output=True) from the UNMODIFIED reference
(optics.py:298-300, :322-325, :584-660).  raman_fortran.txt is not part of the reference checkout (it belongs to
the downloadable reference data), so a synthetic two-column table (wavelength [um], factor) is written into a
temporary $picaso_refdata/opacities/ and stored beside the outputs.  Build container only:

    python tests/golden/make_golden_pollack.py   ->  tests/golden/pollack.npz
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

CASE = "opt_linear_raman"   # db / atmosphere of this optics case, run with raman = 1 instead of 0


def main():
    case = C.optics_cases()[CASE]
    db = synth.opacity_database(**case["db"])
    atm = synth.atmosphere_profile(db, **case["atm"])
    rng = np.random.default_rng(77)
    wave = 1e4 / db["wno"]
    tw = np.linspace(wave.min() * 0.9, wave.max() * 1.1, 41)
    tf = np.clip(0.9 + 0.12 * rng.random(41), 0.0, 1.02)     # some entries above the 0.99999 cap
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "opacities"))
        np.savetxt(os.path.join(tmp, "opacities", "raman_fortran.txt"), np.column_stack([tw, tf]))
        os.environ["picaso_refdata"] = tmp
        O = ref_loader.load_optics()
        sqlite3.register_adapter(np.int64, int)
        sqlite3.register_adapter(np.int32, int)
        path = os.path.join(tmp, "opa.db")
        MG.write_db(path, db)
        raman_txt = os.path.join(ref_loader.REF_ROOT, "reference", "opacities", "raman.txt")
        opa = O.RetrieveOpacities(path, raman_txt, query_method=case["query"])
        a = MG.duck_atmosphere(db, atm)
        opa.get_opacities(a)
        res = O.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"],
                                test_mode=None, raman=1, full_output=True)
    out = {"table_w": tw, "table_f": tf}
    names = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "DTAU_OG", "TAU_OG", "W0_OG",
             "COSB_OG", "W0_no_raman", "f_deltaM")
    for n, arr in zip(names, res):
        arr = np.asarray(arr)
        out["out/" + n] = arr[:, :, 0] if arr.ndim == 3 else arr
    for n in ("taugas", "tauray", "taucld"):
        out["full/" + n] = np.asarray(getattr(a, n))[:, :, 0]
    np.savez_compressed(os.path.join(HERE, "pollack.npz"), ref_commit=MG.REF_COMMIT, case=CASE, **out)
    print("pollack:", len(out), "arrays")


if __name__ == "__main__":
    main()
