"""Golden vectors for the opacity path (a9-a11), produced by the UNMODIFIED reference classes:
RetrieveOpacities (on a synthetic sqlite DB with the reference schema), compute_opacity and
compute_raman of /root/reference/picaso/optics.py.  Build container only.

    python tests/golden/make_golden_optics.py
"""
import io
import os
import sqlite3
import sys
import tempfile
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_loader  # noqa: E402
from picaso_b200 import synth  # noqa: E402
import cases as C  # noqa: E402

REF_COMMIT = "0369089"


def _adapt(arr):
    out = io.BytesIO()
    np.save(out, arr)
    out.seek(0)
    return sqlite3.Binary(out.read())


def write_db(path, db):
    """schema of opacity_factory.build_skeleton (opacity_factory.py:622-668)."""
    sqlite3.register_adapter(np.ndarray, _adapt)
    conn = sqlite3.connect(path, detect_types=sqlite3.PARSE_DECLTYPES)
    cur = conn.cursor()
    cur.execute("CREATE TABLE header (id INTEGER PRIMARY KEY, pressure_unit VARCHAR, temperature_unit VARCHAR, "
                "wavenumber_grid array, continuum_unit VARCHAR, molecular_unit VARCHAR)")
    cur.execute("CREATE TABLE molecular (id INTEGER PRIMARY KEY, ptid INTEGER, molecule VARCHAR, "
                "pressure FLOAT, temperature FLOAT, opacity array)")
    cur.execute("CREATE TABLE continuum (id INTEGER PRIMARY KEY, molecule VARCHAR, temperature FLOAT, "
                "opacity array)")
    cur.execute("INSERT INTO header (pressure_unit, temperature_unit, wavenumber_grid, continuum_unit, "
                "molecular_unit) VALUES (?,?,?,?,?)", ("bar", "kelvin", db["wno"], "cm-1 amagat-2", "cm2/molecule"))
    for m in db["molecules"]:
        for (ptid, p, t), row in zip(db["pt_pairs"], db["tables"][m]):
            cur.execute("INSERT INTO molecular (ptid, molecule, temperature, pressure, opacity) VALUES (?,?,?,?,?)",
                        (ptid, m, float(t), float(p), row))
    for key, tab in db["continuum"].items():
        for t, row in zip(db["cia_temps"], tab):
            cur.execute("INSERT INTO continuum (molecule, temperature, opacity) VALUES (?,?,?)",
                        (key, float(t), row))
    conn.commit()
    conn.close()


def duck_atmosphere(db, atm):
    """the attributes compute_opacity / get_opacities read from ATMSETUP."""
    a = types.SimpleNamespace()
    a.c = types.SimpleNamespace(nlayer=atm["nlayer"], pconv=atm["pconv"], rgas=atm["rgas"], amu=atm["amu"],
                                k_b=atm["k_b"])
    a.level = {"temperature": atm["tlevel"], "pressure": atm["plevel"]}
    a.layer = {"temperature": atm["tlayer"], "pressure": atm["player"], "colden": atm["colden"],
               "mmw": atm["mmw"], "mixingratios": pd.DataFrame(atm["mixingratios"]),
               "electrons": atm["electrons"],
               "cloud": {"opd": atm["cloud_opd"].copy(), "w0": atm["cloud_w0"].copy(), "g0": atm["cloud_g0"].copy()}}
    a.planet = types.SimpleNamespace(gravity=atm["gravity"])
    a.molecules = list(db["molecules"])
    a.continuum_molecules = [list(x) for x in db["continuum_molecules"]]
    a.rayleigh_molecules = list(db["rayleigh_molecules"])
    return a


def main():
    O = ref_loader.load_optics()
    # the reference binds numpy integers in its SQL query (optics.py:2222-2229); python's sqlite3
    # only matches them when an adapter is registered (environment set-up, not a reference change)
    sqlite3.register_adapter(np.int64, int)
    sqlite3.register_adapter(np.int32, int)
    raman_txt = os.path.join(ref_loader.REF_ROOT, "reference", "opacities", "raman.txt")
    out = {}
    for name, case in C.optics_cases().items():
        db = synth.opacity_database(**case["db"])
        atm = synth.atmosphere_profile(db, **case["atm"])
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "opa.db")
            write_db(path, db)
            opa = O.RetrieveOpacities(path, raman_txt, query_method=case["query"])
            assert np.array_equal(opa.wno, db["wno"])
            a = duck_atmosphere(db, atm)
            opa.get_opacities(a)
        # reference-derived INPUTS that cannot be regenerated from a seed elsewhere
        for m in db["rayleigh_molecules"]:
            out[f"{name}/in/rayleigh/{m}"] = opa.rayleigh_opa[m]
        rdb = opa.raman_db
        out[f"{name}/in/raman_c"] = rdb["c"].values
        out[f"{name}/in/raman_ji"] = rdb["ji"].values
        out[f"{name}/in/raman_deltanu"] = rdb["deltanu"].values
        shifts = 0.6 + 0.8 * np.random.default_rng(case["db"]["seed"] + 5).random((db["nwno"], len(rdb)))
        opa.raman_stellar_shifts = shifts
        out[f"{name}/in/stellar_shifts"] = shifts
        # reference OUTPUTS
        for m in db["molecules"]:
            out[f"{name}/molecular_opa/{m}"] = opa.molecular_opa[m]
        for k, v in opa.continuum_opa.items():
            out[f"{name}/continuum_opa/{k}"] = v
        out[f"{name}/pt_opa_index"] = np.asarray(a.layer["pt_opa_index"])
        if case["raman"] == 0:
            out[f"{name}/raman_factor"] = O.compute_raman(db["nwno"], atm["nlayer"], opa.wno, shifts,
                                                          atm["tlayer"], rdb["c"].values, rdb["ji"].values,
                                                          rdb["deltanu"].values)
        res = O.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"],
                                test_mode=None, raman=case["raman"])
        names = ("DTAU", "TAU", "W0", "COSB", "ftau_cld", "ftau_ray", "GCOS2", "DTAU_OG", "TAU_OG", "W0_OG",
                 "COSB_OG", "W0_no_raman", "f_deltaM")
        for n, arr in zip(names, res):
            arr = np.asarray(arr)
            out[f"{name}/out/{n}"] = arr[:, :, 0] if arr.ndim == 3 else arr
    np.savez_compressed(os.path.join(HERE, "optics.npz"), ref_commit=REF_COMMIT, **out)
    print("optics:", len(out), "arrays")


if __name__ == "__main__":
    main()
