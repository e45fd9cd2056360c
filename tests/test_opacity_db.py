"""Host I/O: the sqlite opacity-database reader (picaso_b200/opacity_db.py) on a synthetic file written with the
reference schema (opacity_factory.py:622-668, same writer as tests/golden/make_golden_optics.py), and - on a GPU -
the `opannection` factory feeding compute_opacity."""
import io
import os
import sqlite3

import numpy as np
import pytest

import cases as C
from picaso_b200 import opacity_db, synth


def _adapt(arr):
    out = io.BytesIO()
    np.save(out, arr)
    out.seek(0)
    return sqlite3.Binary(out.read())


def write_db(path, db, shuffle_seed=None):
    conn = sqlite3.connect(path)
    cur = conn.cursor()
    cur.execute("CREATE TABLE header (id INTEGER PRIMARY KEY, pressure_unit VARCHAR, temperature_unit VARCHAR, "
                "wavenumber_grid array, continuum_unit VARCHAR, molecular_unit VARCHAR)")
    cur.execute("CREATE TABLE molecular (id INTEGER PRIMARY KEY, ptid INTEGER, molecule VARCHAR, "
                "pressure FLOAT, temperature FLOAT, opacity array)")
    cur.execute("CREATE TABLE continuum (id INTEGER PRIMARY KEY, molecule VARCHAR, temperature FLOAT, opacity array)")
    cur.execute("INSERT INTO header (pressure_unit, temperature_unit, wavenumber_grid, continuum_unit, molecular_unit) "
                "VALUES (?,?,?,?,?)", ("bar", "kelvin", _adapt(db["wno"]), "cm-1 amagat-2", "cm2/molecule"))
    rows = [(ptid, m, float(t), float(p), _adapt(row)) for m in db["molecules"]
            for (ptid, p, t), row in zip(db["pt_pairs"], db["tables"][m])]
    crow = [(k, float(t), _adapt(row)) for k, tab in db["continuum"].items() for t, row in zip(db["cia_temps"], tab)]
    if shuffle_seed is not None:      # row order in the file must not matter
        rng = np.random.default_rng(shuffle_seed)
        rows = [rows[i] for i in rng.permutation(len(rows))]
        crow = [crow[i] for i in rng.permutation(len(crow))]
    cur.executemany("INSERT INTO molecular (ptid, molecule, temperature, pressure, opacity) VALUES (?,?,?,?,?)", rows)
    cur.executemany("INSERT INTO continuum (molecule, temperature, opacity) VALUES (?,?,?)", crow)
    conn.commit()
    conn.close()


def test_reader_round_trip(tmp_path):
    db = synth.opacity_database(W=120, nmol=3, seed=5)
    path = os.path.join(tmp_path, "opa.db")
    write_db(path, db, shuffle_seed=1)
    got = opacity_db.read_opacity_db(path)
    assert np.array_equal(got["wno"], db["wno"])
    assert got["pt_pairs"] == [(int(a), float(b), float(c)) for a, b, c in db["pt_pairs"]]
    assert sorted(got["tables"]) == sorted(db["molecules"])
    for m in db["molecules"]:
        assert np.array_equal(got["tables"][m], db["tables"][m])
    assert np.array_equal(got["cia_temps"], np.unique(db["cia_temps"]))
    for k, tab in db["continuum"].items():
        assert np.array_equal(got["continuum"][k], tab[np.argsort(db["cia_temps"])])


def test_reader_wave_range_and_resample(tmp_path):
    """opacity[::resample][loc], loc from the resampled grid - optics.py:2027-2036, :2238"""
    db = synth.opacity_database(W=200, nmol=2, seed=6)
    path = os.path.join(tmp_path, "opa.db")
    write_db(path, db)
    got = opacity_db.read_opacity_db(path, wave_range=[0.4, 0.8], resample=3, molecules=["CH4"])
    wno = db["wno"][::3]
    wave = 1e4 / wno
    loc = np.where((wave > 0.4) & (wave < 0.8))
    assert np.array_equal(got["wno"], wno[loc]) and got["wno"].size > 5
    assert list(got["tables"]) == ["CH4"]
    assert np.array_equal(got["tables"]["CH4"], db["tables"]["CH4"][:, ::3][:, loc[0]])
    for k, tab in db["continuum"].items():
        assert np.array_equal(got["continuum"][k], tab[np.argsort(db["cia_temps"])][:, ::3][:, loc[0]])


def test_reader_rejects_incomplete_molecule(tmp_path):
    db = synth.opacity_database(W=30, nmol=2, seed=7)
    db["tables"][db["molecules"][1]] = db["tables"][db["molecules"][1]][:-2]   # two (P, T) rows missing
    path = os.path.join(tmp_path, "opa.db")
    write_db(path, db)
    with pytest.raises(ValueError):
        opacity_db.read_opacity_db(path)


def test_raman_table(tmp_path):
    p = os.path.join(tmp_path, "raman.txt")
    ji, c, dnu = synth.raman_table(seed=3, n=12)
    with open(p, "w") as f:
        f.write("# header\n" * 16)
        for a, b, d in zip(ji, c, dnu):
            f.write("%d %d %d %.17e %.17e\n" % (a, a + 2, 0, b, d))
    c2, ji2, dnu2 = opacity_db.read_raman_table(p)
    assert np.array_equal(ji2, ji) and np.array_equal(c2, c) and np.array_equal(dnu2, dnu)


@pytest.mark.gpu
def test_opannection_feeds_compute_opacity(tmp_path):
    import picaso_b200 as pb
    from optics_util import OUT_NAMES, duck_atmosphere, load_case
    from util import assert_close
    name = "opt_linear_raman" if "opt_linear_raman" in C.optics_cases() else sorted(C.optics_cases())[0]
    case, g, db, atm, ins = load_case(name)
    path = os.path.join(tmp_path, "opa.db")
    write_db(path, db, shuffle_seed=2)
    opa = opacity_db.opannection(path, ins["rayleigh"], query_method=case["query"])
    opa.raman_db = (ins["raman_c"], ins["raman_ji"], ins["raman_deltanu"])
    if case["raman"] == 0:
        opa.raman_stellar_shifts = ins["stellar_shifts"]
    a = duck_atmosphere(db, atm)
    opa.get_opacities(a)
    res = pb.compute_opacity(a, opa, ngauss=1, stream=case["stream"], delta_eddington=case["dedd"], test_mode=None,
                             raman=case["raman"])
    for n, arr in zip(OUT_NAMES, res):
        assert_close(arr[:, :, 0], g[f"{name}/out/{n}"], 1e-10, name + " " + n)
    opa.close()
