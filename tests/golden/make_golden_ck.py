"""Golden vectors for the pre-mixed correlated-k opacity path: the UNMODIFIED reference methods
RetrieveCKs.get_pre_mix_ck / get_continuum (bound to a bare instance whose attributes are filled from
a synthetic table + sqlite continuum DB) and compute_opacity(ngauss=K).  Build container only."""
import os
import sqlite3
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref_loader  # noqa: E402
from picaso_b200 import synth  # noqa: E402
import cases as C  # noqa: E402
import make_golden_optics as MO  # noqa: E402

NAMES = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "DTAU_OG", "TAU_OG", "W0_OG", "COSB_OG",
         "W0_no_raman", "f_deltaM")


def main():
    O = ref_loader.load_optics()
    R = ref_loader.load_optics().Rayleigh
    out = {}
    for name, case in C.ck_cases().items():
        db = synth.ck_database(**case["db"])
        atm = synth.atmosphere_profile(dict(db, molecules=[]), **case["atm"])
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "cont.db")
            MO.write_db(path, dict(db, molecules=[], tables={}, pt_pairs=[]))
            opa = object.__new__(O.RetrieveCKs)
            opa.db_filename = path
            opa.continuum_db = path
            opa.pressures, opa.temps, opa.nc_p = db["pressures"], db["temps"], db["nc_p"]
            opa.kappa, opa.nwno, opa.wno, opa.ngauss = db["kappa"], db["nwno"], db["wno"], db["ngauss"]
            opa.cia_temps = db["cia_temps"]
            opa.gauss_wts = db["gauss_wts"]
            ray = R(db["wno"])
            opa.rayleigh_opa = {m: ray.compute_sigma(m) for m in db["rayleigh_molecules"]}
            a = MO.duck_atmosphere(dict(db, molecules=[]), atm)
            opa.get_continuum(a)
            opa.get_pre_mix_ck(a)
        for m in db["rayleigh_molecules"]:
            out[f"{name}/in/rayleigh/{m}"] = opa.rayleigh_opa[m]
        out[f"{name}/molecular_opa"] = opa.molecular_opa
        for k, v in opa.continuum_opa.items():
            out[f"{name}/continuum_opa/{k}"] = v
        res = O.compute_opacity(a, opa, ngauss=db["ngauss"], stream=case["stream"], delta_eddington=case["dedd"],
                                test_mode=None, raman=2)
        for n, arr in zip(NAMES, res):
            out[f"{name}/out/{n}"] = np.asarray(arr)
    np.savez_compressed(os.path.join(HERE, "ck.npz"), ref_commit="0369089", **out)
    print("ck:", len(out), "arrays")


if __name__ == "__main__":
    main()
