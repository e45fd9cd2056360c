"""CPU: the vectorised host-side planning code of picaso_b200/optics.py (what a spectrum costs on the host between two
kernel launches: bin search, interpolation weights, per-layer multipliers) against its statement-per-species versions
(tests/support/host_plan_reference.py), bit for bit.  No GPU: a DeviceOpacities object is assembled without its
device tables."""
import os
import sys
import types

import numpy as np
import pandas as pd
import pytest

from picaso_b200 import optics, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "support"))
import host_plan_reference as ref  # noqa: E402


def host_only_opacities(db, ray_names, query="linear"):
    """DeviceOpacities minus everything that needs a device (the attributes get_opacities / _layer_scalars read)"""
    opa = optics.DeviceOpacities.__new__(optics.DeviceOpacities)
    pp = [(int(p[0]), float(p[1]), float(p[2])) for p in db["pt_pairs"]]
    P = np.array([p[1] for p in pp])
    T = np.array([p[2] for p in pp])
    opa.query_method = query
    opa.wno, opa.nwno = db["wno"], db["wno"].size
    opa._ptid, opa._lnP, opa._T = np.array([p[0] for p in pp]), np.log(P), T
    opa.temps = T[np.sort(np.unique(T, return_index=True)[1])]
    opa.pressures = P[np.sort(np.unique(P, return_index=True)[1])]
    opa.nc_p = np.array([np.sum(T == t) for t in np.unique(T)])
    opa.t_inv_grid, opa.p_log_grid = 1 / opa.temps, np.log10(opa.pressures)
    opa.cia_temps = np.asarray(db["cia_temps"], dtype=np.float64)
    opa._cia_unique = np.unique(opa.cia_temps)
    opa._mol_index = {m: i for i, m in enumerate(db["tables"])}
    opa._cont_index = {k: i for i, k in enumerate(db["continuum"])}
    opa._ray_index = {k: i for i, k in enumerate(ray_names)}
    return opa


def duck(db, atm, frame):
    a = types.SimpleNamespace()
    a.c = types.SimpleNamespace(nlayer=atm["nlayer"], pconv=atm["pconv"], rgas=atm["rgas"], amu=atm["amu"], k_b=atm["k_b"])
    a.level = {"temperature": atm["tlevel"], "pressure": atm["plevel"]}
    a.layer = {"temperature": atm["tlayer"], "pressure": atm["player"], "colden": atm["colden"], "mmw": atm["mmw"],
               "mixingratios": pd.DataFrame(atm["mixingratios"]) if frame else atm["mixingratios"],
               "electrons": atm["electrons"], "cloud": None}
    a.planet = types.SimpleNamespace(gravity=atm["gravity"])
    a.molecules = list(db["molecules"])
    a.continuum_molecules = [list(x) for x in db["continuum_molecules"]]
    a.rayleigh_molecules = list(db["rayleigh_molecules"])
    return a


@pytest.mark.parametrize("seed,ragged", [(1, True), (2, False), (3, True)])
def test_bin_search_fast_path_matches_masks(seed, ragged):
    """searchsorted on monotonic grids == the comparison masks: random profiles, profiles outside the grid, layers
    exactly on grid points; and a shuffled (non-monotonic) temperature grid takes the general path"""
    rng = np.random.default_rng(seed)
    db = synth.opacity_database(W=8, nmol=1, seed=seed, nT=14, nP=11, ragged=ragged)
    tlayer = np.concatenate([rng.uniform(40.0, 5000.0, size=80), db["temps"], [db["temps"][-1]] * 3])
    pbar = np.concatenate([10.0 ** rng.uniform(-7.5, 4.0, size=80 + db["temps"].size - db["pressures"].size),
                           db["pressures"], [db["pressures"][0]] * 3])
    args = (1 / db["temps"], np.log10(db["pressures"]), db["nc_p"], tlayer, pbar)
    for a, b in zip(optics.find_needed_pts_grid(*args), ref.find_needed_pts_grid(*args)):
        assert np.array_equal(a, b)
    # equal-length pressure columns so that a permuted temperature axis is still a valid grid description
    db2 = synth.opacity_database(W=8, nmol=1, seed=seed, nT=9, nP=7, ragged=False)
    perm = rng.permutation(db2["temps"].size)
    args = (1 / db2["temps"][perm], np.log10(db2["pressures"]), db2["nc_p"][perm], tlayer, pbar)
    for a, b in zip(optics.find_needed_pts_grid(*args), ref.find_needed_pts_grid(*args)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("frame", [False, True])
@pytest.mark.parametrize("seed", [11, 12])
def test_plan_and_layer_scalars_bit_for_bit(seed, frame):
    db = synth.opacity_database(W=16, nmol=5, seed=seed, nT=12, nP=10, ragged=True)
    atm = synth.atmosphere_profile(db, L=23, seed=seed + 100)
    ray_names = list(db["rayleigh_molecules"])
    opa = host_only_opacities(db, ray_names)
    a = duck(db, atm, frame)
    opa.get_opacities(a)
    plan = opa._plan
    # the plan, assembled the statement-per-column way
    pbar = atm["player"] / atm["pconv"]
    t, p, ill, ihl, ilh, ihh = ref.find_needed_pts_grid(opa.t_inv_grid, opa.p_log_grid, opa.nc_p, atm["tlayer"], pbar)
    t, p = t[:, 0], p[:, 0]
    assert plan["idx"].dtype == np.int32 and np.array_equal(plan["idx"], np.stack([ill, ihl, ihh, ilh], axis=1))
    want_w = np.stack([(1 - t) * (1 - p), t * (1 - p), t * p, (1 - t) * p], axis=1)
    assert np.array_equal(plan["wts"], want_w)
    assert np.array_equal(np.asarray(a.layer["pt_opa_index"]), 1 + np.unique(np.concatenate([ill, ihl, ilh, ihh])))
    assert np.array_equal(plan["cia"], np.abs(opa._cia_unique[None, :] - atm["tlayer"][:, None]).argmin(axis=1))
    assert all(plan["fac"][m] == 1 for m in a.molecules)
    # the multipliers
    got = optics._layer_scalars(a, opa)
    want = ref._layer_scalars(a, opa)
    for g_, w_, n in zip(got, want, ("mol", "cont", "ray")):
        assert g_.shape == w_.shape and np.array_equal(g_, w_), n
    # excluded molecules (exclude_mol dict, justdoit.py:223) scale their rows
    ex = {m: (0 if i % 2 else 1) for i, m in enumerate(a.molecules)}
    opa.get_opacities(a, exclude_mol=ex)
    got = optics._layer_scalars(a, opa)
    want = ref._layer_scalars(a, opa)
    for g_, w_ in zip(got, want):
        assert np.array_equal(g_, w_)


def c_plan(t_inv_grid, p_log_grid, nc_p, tlayer, pbar, cia_unique):
    """pb_host_plan_bilinear called directly (what DeviceOpacities.get_opacities does for query_method='linear')"""
    import ctypes
    from picaso_b200 import _lib
    fn = _lib.load_library().pb_host_plan_bilinear
    L = tlayer.size
    tg, pg = np.ascontiguousarray(t_inv_grid), np.ascontiguousarray(p_log_grid)
    ncp = np.ascontiguousarray(nc_p, dtype=np.int64)
    off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(ncp)]), dtype=np.int64)
    t_mono, p_mono = optics.grid_is_monotonic(tg, pg)
    tl = np.ascontiguousarray(tlayer)
    t_inv, p_log = 1 / tl, np.log10(pbar)
    idx, wts = np.empty((L, 4), dtype=np.int32), np.empty((L, 4))
    cia, rows, n = np.empty(L, dtype=np.int32), np.empty(4 * L, dtype=np.int64), ctypes.c_int(0)
    cu = np.ascontiguousarray(cia_unique)
    rc = fn(L, t_inv.ctypes.data, p_log.ctypes.data, tl.ctypes.data, tg.size, tg.ctypes.data, pg.size, pg.ctypes.data,
            ncp.ctypes.data, off.ctypes.data, int(t_mono), int(p_mono), cu.size, cu.ctypes.data, idx.ctypes.data,
            wts.ctypes.data, cia.ctypes.data, rows.ctypes.data, ctypes.addressof(n))
    assert rc == 0
    return idx, wts, cia, rows[:n.value], (t_mono, p_mono)


@pytest.mark.parametrize("seed,ragged,shuffle", [(1, True, False), (2, False, False), (3, True, False), (4, False, True)])
def test_c_planner_bit_for_bit(seed, ragged, shuffle):
    """the C planner (csrc/host_plan.cu) against the mask-based numpy version: random profiles, profiles far outside the
    grid, layers exactly on grid points, ragged pressure columns, and a non-monotonic temperature axis (scan path)"""
    rng = np.random.default_rng(seed)
    db = synth.opacity_database(W=8, nmol=1, seed=seed, nT=13, nP=9, ragged=ragged)
    temps, press, nc_p = db["temps"], db["pressures"], db["nc_p"]
    if shuffle:
        perm = rng.permutation(temps.size)
        temps, nc_p = temps[perm], nc_p[perm]
    n_extra = temps.size + press.size
    tlayer = np.concatenate([rng.uniform(30.0, 6000.0, size=120), temps, rng.uniform(200, 2000, size=press.size)])
    pbar = np.concatenate([10.0 ** rng.uniform(-8.0, 4.5, size=120), 10.0 ** rng.uniform(-6, 2, size=temps.size), press])
    cia_unique = np.unique(rng.uniform(50, 3000, size=17))
    idx, wts, cia, rows, mono = c_plan(1 / temps, np.log10(press), nc_p, tlayer, pbar, cia_unique)
    assert mono == (not shuffle, True)
    t, p, ill, ihl, ilh, ihh = ref.find_needed_pts_grid(1 / temps, np.log10(press), nc_p, tlayer, pbar)
    t, p = t[:, 0], p[:, 0]
    assert np.array_equal(idx, np.stack([ill, ihl, ihh, ilh], axis=1))
    assert np.array_equal(wts, np.stack([(1 - t) * (1 - p), t * (1 - p), t * p, (1 - t) * p], axis=1), equal_nan=True)
    assert np.array_equal(cia, np.abs(cia_unique[None, :] - tlayer[:, None]).argmin(axis=1))
    assert np.array_equal(rows, 1 + np.unique(np.concatenate([ill, ihl, ilh, ihh])))
    assert n_extra > 0
