"""Straightforward (one numpy statement per reference line) versions of the host-side planning code of
picaso_b200/optics.py, kept as the bit-for-bit reference of its vectorised production versions
(tests/test_host_plan_cpu.py).  They are the round-1 implementations, validated on the GPU against the reference
golden vectors through compute_opacity (tests/test_gpu_optics.py)."""
import numpy as np


def find_needed_pts_grid(t_inv_grid, p_log_grid, nc_p, tlayer, player):
    """RetrieveOpacities.find_needed_pts (picaso/optics.py:2048-2123) for all layers at once: the bilinear
    neighbours in (1/T, log10 P) of every layer on a (T-major, P-minor, possibly ragged) table grid.
    Returns t_interp[:, None], p_interp[:, None] and the four 0-based row indices (ll, hl, lh, hh)."""
    t_inv = 1 / np.asarray(tlayer, dtype=np.float64)
    p_log = np.log10(np.asarray(player, dtype=np.float64))
    nT = t_inv_grid.size

    def last_true(mask):
        """per row: index of the last True (np.where(row)[0][-1]), 0 if none - any grid ordering"""
        n = mask.shape[1]
        return np.where(mask.any(axis=1), n - 1 - np.argmax(mask[:, ::-1], axis=1), 0)

    # last grid temperature strictly below T, last grid pressure <= P
    t_low = last_true(t_inv_grid[None, :] > t_inv[:, None])
    t_low = np.where(t_low == nT - 1, nT - 2, t_low)
    t_hi = t_low + 1
    p_low = last_true(p_log_grid[None, :] <= p_log[:, None])
    p_low = np.minimum(p_low, nc_p[t_hi] - 3)
    p_hi = p_low + 1
    off = np.concatenate([[0], np.cumsum(nc_p)])
    t_interp = ((t_inv - t_inv_grid[t_low]) / (t_inv_grid[t_hi] - t_inv_grid[t_low]))[:, np.newaxis]
    p_interp = ((p_log - p_log_grid[p_low]) / (p_log_grid[p_hi] - p_log_grid[p_low]))[:, np.newaxis]
    return (t_interp, p_interp, off[t_low] + p_low, off[t_hi] + p_low, off[t_low] + p_hi, off[t_hi] + p_hi)


def _layer_scalars(atm, opa):
    """per-layer multipliers exactly as compute_opacity parenthesises them (optics.py:147-271)."""
    L = atm.c.nlayer
    tlevel = np.asarray(atm.level["temperature"], dtype=np.float64)
    plevel = np.asarray(atm.level["pressure"], dtype=np.float64) / atm.c.pconv
    tlayer = np.asarray(atm.layer["temperature"], dtype=np.float64)
    gravity = atm.planet.gravity / 100.0
    mmw = np.asarray(atm.layer["mmw"], dtype=np.float64)
    colden = np.asarray(atm.layer["colden"], dtype=np.float64)
    player = np.asarray(atm.layer["pressure"], dtype=np.float64)
    mix = atm.layer["mixingratios"]
    x = lambda s: np.asarray(mix[s].values if hasattr(mix[s], "values") else mix[s], dtype=np.float64)
    ACOEF = (tlayer / (tlevel[:-1] * tlevel[1:])) * (
        tlevel[1:] * plevel[1:] - tlevel[:-1] * plevel[:-1]) / (plevel[1:] - plevel[:-1])
    BCOEF = (tlayer / (tlevel[:-1] * tlevel[1:])) * (tlevel[:-1] - tlevel[1:]) / (plevel[1:] - plevel[:-1])
    COEF1 = atm.c.rgas * 273.15 ** 2 * .5E5 * (
        ACOEF * (plevel[1:] ** 2 - plevel[:-1] ** 2) + BCOEF * (2. / 3.) * (plevel[1:] ** 3 - plevel[:-1] ** 3)) / (
        1.01325 ** 2 * gravity * tlayer * mmw)
    cont = np.zeros((len(opa._cont_index), L))
    used = set()
    for m in atm.continuum_molecules:
        key = m[0] + m[1]
        if key not in opa._cont_index:
            raise KeyError(f"continuum pair {key} is not in the uploaded tables")
        used.add(key)
        if m[0] == "H-" and m[1] == "bf":
            s = (x("H-") * colden / (mmw * atm.c.amu))
        elif m[0] == "H-" and m[1] == "ff":
            s = (player * x("H") * np.asarray(atm.layer["electrons"]) * colden / (tlayer * mmw * atm.c.amu * atm.c.k_b))
        elif m[0] == "H2-" and m[1] == "":
            s = (player * x("H2") * np.asarray(atm.layer["electrons"]) * colden / (mmw * atm.c.amu))
        else:
            s = (COEF1 * x(m[0]) * x(m[1]))
        cont[opa._cont_index[key]] = s
    mol = np.zeros((len(opa._mol_index), L))
    for m in atm.molecules:
        if m not in opa._mol_index:
            raise KeyError(f"molecule {m} is not in the uploaded tables")
        mol[opa._mol_index[m]] = opa._plan["fac"][m] * (colden * x(m) / mmw)
    ray = np.zeros((len(opa._ray_index), L))
    for m in atm.rayleigh_molecules:
        ray[opa._ray_index[m]] = (colden * x(m) / mmw)
    return mol, cont, ray
