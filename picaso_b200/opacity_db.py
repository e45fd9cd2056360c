"""Reader for the reference's sqlite opacity databases -> a device-resident `DeviceOpacities`.

Host I/O only.  Follows `RetrieveOpacities.get_available_data` (picaso/optics.py:1999-2046), `open_local` /
`convert_array` (:1960-1985: blobs are `np.save` bytes) and the row queries of `_get_query_molecular` /
`_get_query_continuum` (:2160-2239) against the schema written by `opacity_factory.build_skeleton`
(picaso/opacity_factory.py:622-668):

    header(wavenumber_grid array, ...), molecular(ptid, molecule, pressure, temperature, opacity array),
    continuum(molecule, temperature, opacity array)

The reference keeps the file open and fetches the rows an atmosphere needs on every spectrum; here every
row is read once and handed to `DeviceOpacities`, which uploads them to HBM (`pb_optab_*`).  Wavelength
selection is the reference's: `opacity[::resample][loc]` with `loc = (wave > min(wave_range)) & (wave <
max(wave_range))`.
"""
import io
import sqlite3

import numpy as np

__all__ = ["read_opacity_db", "read_raman_table", "opannection"]


def _blob(b):
    if isinstance(b, np.ndarray):
        return b
    return np.load(io.BytesIO(b))


def read_opacity_db(db_filename, wave_range=None, resample=1, molecules=None):
    """-> dict(wno, pt_pairs, tables {molecule: [npt, nwno]}, cia_temps, continuum {pair: [ntemp, nwno]}).

    pt_pairs are (ptid, pressure[bar], temperature[K]) sorted by ptid, tables follow that order; continuum rows
    are in ascending temperature order.  `molecules` restricts the molecular tables that are read."""
    conn = sqlite3.connect(db_filename)
    try:
        cur = conn.cursor()
        cur.execute("SELECT wavenumber_grid FROM header")
        wno_all = np.asarray(_blob(cur.fetchone()[0]), dtype=np.float64)[::resample]
        wave = 1e4 / wno_all
        if wave_range is None:
            loc = np.arange(wno_all.size)
        else:
            loc = np.where((wave > min(wave_range)) & (wave < max(wave_range)))[0]
        wno = np.ascontiguousarray(wno_all[loc])
        cur.execute("SELECT ptid, pressure, temperature FROM molecular")
        pt_pairs = sorted(set(cur.fetchall()), key=lambda x: x[0])
        row_of = {int(p[0]): i for i, p in enumerate(pt_pairs)}
        cur.execute("SELECT DISTINCT molecule FROM molecular")
        avail = sorted(r[0] for r in cur.fetchall())
        want = avail if molecules is None else [m for m in molecules if m in avail]
        tables = {}
        for m in want:
            tab = np.zeros((len(pt_pairs), wno.size))
            seen = np.zeros(len(pt_pairs), dtype=bool)
            cur.execute("SELECT ptid, opacity FROM molecular WHERE molecule = ?", (m,))
            for ptid, blob in cur.fetchall():
                i = row_of[int(ptid)]
                tab[i] = np.asarray(_blob(blob), dtype=np.float64)[::resample][loc]
                seen[i] = True
            if not seen.all():
                raise ValueError("opacity DB %s: molecule %s lacks %d of the %d (P, T) points"
                                 % (db_filename, m, int((~seen).sum()), len(pt_pairs)))
            tables[m] = tab
        cur.execute("SELECT temperature FROM continuum")
        cia_temps = np.unique(np.array([r[0] for r in cur.fetchall()], dtype=np.float64))
        t_row = {float(t): i for i, t in enumerate(cia_temps)}
        cur.execute("SELECT DISTINCT molecule FROM continuum")
        continuum = {}
        for (pair,) in cur.fetchall():
            tab = np.zeros((cia_temps.size, wno.size))
            c2 = conn.cursor()
            c2.execute("SELECT temperature, opacity FROM continuum WHERE molecule = ?", (pair,))
            for t, blob in c2.fetchall():
                tab[t_row[float(t)]] = np.asarray(_blob(blob), dtype=np.float64)[::resample][loc]
            continuum[pair] = tab
    finally:
        conn.close()
    return dict(wno=wno, pt_pairs=[(int(a), float(b), float(c)) for a, b, c in pt_pairs], tables=tables,
                cia_temps=cia_temps, continuum=continuum)


def read_raman_table(raman_data):
    """reference/opacities/raman.txt as RetrieveOpacities reads it (optics.py:1957-1961: whitespace separated,
    16 header lines, columns ji jf vf c deltanu) -> (c, ji, deltanu)"""
    rows = np.loadtxt(raman_data, skiprows=16, ndmin=2)
    return rows[:, 3].copy(), rows[:, 0].astype(np.int64), rows[:, 4].copy()


def opannection(db_filename, rayleigh_opa, raman_data=None, wave_range=None, resample=1, query_method="nearest",
                molecules=None, ctx=None):
    """The reference's `justdoit.opannection(...)` for a resampled-opacity sqlite file, returning a
    `DeviceOpacities` with every table resident in HBM.  `rayleigh_opa` is {molecule: sigma[nwno]} on the
    selected grid, or a callable wno -> that dict (the reference computes it with `rayleigh.Rayleigh(wno)`,
    optics.py:2040-2046)."""
    from .optics import DeviceOpacities
    d = read_opacity_db(db_filename, wave_range=wave_range, resample=resample, molecules=molecules)
    ray = rayleigh_opa(d["wno"]) if callable(rayleigh_opa) else rayleigh_opa
    raman_db = read_raman_table(raman_data) if raman_data is not None else None
    return DeviceOpacities(d["wno"], d["pt_pairs"], d["tables"], d["cia_temps"], d["continuum"], ray,
                           raman_db=raman_db, query_method=query_method, ctx=ctx)
