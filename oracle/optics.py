"""CPU (numpy) restatement of the reference's per-layer opacity path - TEST INFRASTRUCTURE ONLY.

Follows /root/reference/picaso/optics.py (commit 0369089); each function cites its lines.
Pinned against the unmodified reference classes (RetrieveOpacities on a synthetic sqlite DB,
compute_opacity, compute_raman) by tests/golden/make_golden.py -> tests/golden/optics.npz.
Only tests/, smoke() and bench.py's cpu_baseline may import this.
"""
import numpy as np

N_A = 6.02214086e+23  # the constant the reference multiplies cross-sections by (optics.py:2294)


# ---- a10: where to look in the (T, P) grid ---------------------------------------------------
def find_needed_pts(temps, pressures, nc_p, tlayer, player_bar):
    """optics.py:2048-2123.  temps ascending [nT], pressures [nP] (bar, ascending), nc_p[nT]
    pressures available per temperature.  Returns t_interp[L], p_interp[L] and the four 0-based
    row indices (t_low,p_low), (t_hi,p_low), (t_low,p_hi), (t_hi,p_hi) into the T-major table."""
    t_inv = 1 / np.asarray(tlayer, dtype=np.float64)
    p_log = np.log10(np.asarray(player_bar, dtype=np.float64))
    t_inv_grid = 1 / np.asarray(temps, dtype=np.float64)
    p_log_grid = np.log10(np.asarray(pressures, dtype=np.float64))
    nc_p = np.asarray(nc_p)
    L = t_inv.size
    t_low = np.zeros(L, dtype=np.int64)
    for i in range(L):
        find = np.where(t_inv_grid > t_inv[i])[0]
        t_low[i] = 0 if len(find) == 0 else find[-1]
    t_low[t_low == (len(t_inv_grid) - 1)] = len(t_inv_grid) - 2
    t_hi = t_low + 1
    p_low = np.zeros(L, dtype=np.int64)
    for i in range(L):
        find = np.where(p_log_grid <= p_log[i])[0]
        p_low[i] = 0 if len(find) == 0 else find[-1]
        p_low[i] = min(p_low[i], nc_p[t_hi[i]] - 3)
    p_hi = p_low + 1
    off = np.concatenate([[0], np.cumsum(nc_p)])
    t_interp = (t_inv - t_inv_grid[t_low]) / (t_inv_grid[t_hi] - t_inv_grid[t_low])
    p_interp = (p_log - p_log_grid[p_low]) / (p_log_grid[p_hi] - p_log_grid[p_low])
    return (t_interp, p_interp, off[t_low] + p_low, off[t_hi] + p_low, off[t_low] + p_hi,
            off[t_hi] + p_hi)


def nearest_pt(pt_pairs, tlayer, player_bar):
    """optics.py:2330-2332: first (ptid-ordered) minimiser of hypot(ln P diff, T diff); returns
    0-based table rows."""
    P = np.array([p[1] for p in pt_pairs])
    T = np.array([p[2] for p in pt_pairs])
    out = []
    for p, t in zip(player_bar, tlayer):
        out.append(int(np.argmin(np.hypot(np.log(P) - np.log(p), T - t))))
    return np.array(out, dtype=np.int64)


def nearest_cia_temp(cia_temps, tlayer):
    """optics.py:2298 / :2355 with find_nearest :2418-2421: index into unique(cia_temps)."""
    u = np.unique(cia_temps)
    return np.array([int(np.abs(u - t).argmin()) for t in tlayer], dtype=np.int64)


def interp_molecular(table, t_interp, p_interp, i_ll, i_hl, i_lh, i_hh):
    """optics.py:2277-2294 for one molecule: table [nPT, W] raw cross-sections -> [L, W]."""
    lg = lambda a: np.log10(np.where(a != 0, a, 1e-50))
    t = np.asarray(t_interp)[:, None]
    p = np.asarray(p_interp)[:, None]
    cx = 10 ** (((1 - t) * (1 - p) * lg(table[i_ll])) + ((t) * (1 - p) * lg(table[i_hl])) +
                ((t) * (p) * lg(table[i_hh])) + ((1 - t) * (p) * lg(table[i_lh])))
    return cx * N_A


def nearest_molecular(table, ind_pt):
    """optics.py:2350-2351."""
    return table[ind_pt] * N_A


# ---- a11: Raman ---------------------------------------------------------------------------------
def _partition_function(j, T):
    """optics.py:541-549 (as coded: b_energy already carries j(j+1))."""
    k = 1.38064852e-16
    b = 60.853
    c = 29979245800
    h = 6.62607004e-27
    b_energy = (b * (h) * (c) * j * (j + 1) / k)
    g = (2.0 * j + 1.0) if j % 2 == 0 else 3.0 * (2.0 * j + 1.0)
    return g * np.exp(-0.5 * b_energy * j * (j + 1) / T)


def j_fraction(j, T):
    """optics.py:552-581."""
    Z = np.zeros(np.size(T))
    for jj in range(0, 20):
        Z += _partition_function(jj, T)
    return _partition_function(j, T) / Z


def compute_raman(nwno, nlayer, wno, stellar_shifts, tlayer, cross_sections, j_initial, deltanu):
    """optics.py:467-494."""
    w_shift = np.zeros((nlayer, nwno))
    wo_shift = np.zeros((nlayer, nwno))
    ray = np.zeros((nlayer, nwno))
    jt = np.zeros((10, nlayer))
    for i in range(10):
        jt[i, :] = j_fraction(i, tlayer)
    for i in range(len(cross_sections)):
        ji = int(j_initial[i])
        Q = cross_sections[i] / wno ** 3.0 / (wno + deltanu[i])
        if deltanu[i] == 0:
            ray += np.outer(jt[ji, :], Q)
        else:
            w_shift += np.outer(jt[ji, :], Q * stellar_shifts[:, i])
            wo_shift += np.outer(jt[ji, :], Q)
    return (ray + w_shift) / (ray + wo_shift)


# ---- a9: assembling the layer optical properties --------------------------------------------------
def raman_pollack(wno, table_w, table_f, nlayer):
    """optics.py:652-660: np.interp of the raman_fortran.txt table onto 1e4 / wno, one identical row per layer"""
    row = np.interp(1e4 / np.asarray(wno, dtype=np.float64), table_w, table_f)
    return np.array([row] * nlayer)


def compute_opacity(atm, molecular_opa, continuum_opa, rayleigh_opa, raman_factor, stream=2,
                    delta_eddington=True, fthin_cld=None, do_holes=False, full=None, test_mode=None):
    """optics.py:147-431 for ngauss = 1; test_mode None, 'rayleigh' or any other string (optics.py:372-399; like the
    reference, a test mode replaces non-positive cloud single-scattering albedos by 1e-10 IN atm["cloud_w0"]).

    atm: dict from picaso_b200.synth.atmosphere_profile; molecular_opa {mol: [L, W]} (already x N_A),
    continuum_opa {pair: [L, W]}, rayleigh_opa {mol: [W]}, raman_factor [L, W] BEFORE the 0.99999
    cap, or None for raman = 2 ("none").  Returns the reference's 13-tuple of [L|V, W] arrays; a dict passed as
    `full` receives taugas / tauray / taucld (full_output)."""
    L = atm["nlayer"]
    W = next(iter(rayleigh_opa.values())).shape[0]
    mix = atm["mixingratios"]
    tlevel = atm["tlevel"]
    plevel = atm["plevel"] / atm["pconv"]
    tlayer = atm["tlayer"]
    gravity = atm["gravity"] / 100.0
    ACOEF = (tlayer / (tlevel[:-1] * tlevel[1:])) * (
        tlevel[1:] * plevel[1:] - tlevel[:-1] * plevel[:-1]) / (plevel[1:] - plevel[:-1])
    BCOEF = (tlayer / (tlevel[:-1] * tlevel[1:])) * (tlevel[:-1] - tlevel[1:]) / (plevel[1:] - plevel[:-1])
    COEF1 = atm["rgas"] * 273.15 ** 2 * .5E5 * (
        ACOEF * (plevel[1:] ** 2 - plevel[:-1] ** 2) + BCOEF * (2. / 3.) * (plevel[1:] ** 3 - plevel[:-1] ** 3)) / (
        1.01325 ** 2 * gravity * tlayer * atm["mmw"])
    colden = atm["colden"][:, None]
    mmw = atm["mmw"][:, None]
    player = atm["player"][:, None]
    tl = atm["tlayer"][:, None]
    TAUGAS = np.zeros((L, W))
    for key, kap in continuum_opa.items():
        if key == "H-bf":
            TAUGAS += kap * (mix["H-"][:, None] * colden / (mmw * atm["amu"]))
        elif key == "H-ff":
            TAUGAS += kap * (player * mix["H"][:, None] * atm["electrons"][:, None] * colden /
                             (tl * mmw * atm["amu"] * atm["k_b"]))
        elif key == "H2-":
            TAUGAS += kap * (player * mix["H2"][:, None] * atm["electrons"][:, None] * colden /
                             (mmw * atm["amu"]))
        else:
            a, b = atm["cia_pairs"][key]
            TAUGAS += kap * (COEF1[:, None] * mix[a][:, None] * mix[b][:, None])
    for m, kap in molecular_opa.items():
        TAUGAS += kap * (colden * mix[m][:, None] / mmw)
    TAURAY = np.zeros((L, W))
    for m, sig in rayleigh_opa.items():
        TAURAY += np.array([sig] * L) * (colden * mix[m][:, None] / mmw)
    if raman_factor is None:
        rf = 0.99999 + np.zeros((L, W))
    else:
        rf = np.minimum(raman_factor, raman_factor * 0 + 0.99999)
    TAUCLD = atm["cloud_opd"].copy()
    g0 = atm["cloud_g0"]
    w0c = atm["cloud_w0"]
    if do_holes:
        TAUCLD = fthin_cld * TAUCLD
    if full is not None:   # full_output (optics.py:322-325): atmosphere.taugas / tauray / taucld
        full.update(taugas=TAUGAS.copy(), tauray=TAURAY.copy(), taucld=TAUCLD.copy())
    with np.errstate(all="ignore"):
        DTAU = TAUGAS + TAURAY + TAUCLD
        ftau_cld = (w0c * TAUCLD) / (w0c * TAUCLD + TAURAY)
        COSB = g0
        ftau_ray = TAURAY / (TAURAY + w0c * TAUCLD)
        GCOS2 = 0.5 * ftau_ray
        W0 = (TAURAY * rf + TAUCLD * w0c) / (TAUGAS + TAURAY + TAUCLD)
        W0_no_raman = (TAURAY * 0.99999 + TAUCLD * w0c) / (TAUGAS + TAURAY + TAUCLD)
        TAU = np.zeros((L + 1, W))
        TAU[1:] = np.cumsum(DTAU, axis=0)
        if test_mode is not None:
            # optics.py:372-399 (check against Dlugach & Yanovitskij): Rayleigh-only or cloud-only optical depths with
            # the cloud's single-scattering albedo and asymmetry everywhere
            if test_mode == 'rayleigh':
                DTAU = TAURAY
                GCOS2 = np.zeros(DTAU.shape) + 0.5
                ftau_ray = np.zeros(DTAU.shape) + 1.0
                ftau_cld = np.zeros(DTAU.shape)
            else:
                DTAU = np.zeros(DTAU.shape)
                DTAU[:, :] = atm["cloud_opd"]
                GCOS2 = np.zeros(DTAU.shape)
                ftau_ray = np.zeros(DTAU.shape)
                ftau_cld = np.zeros(DTAU.shape) + 1.
            atm["cloud_w0"][atm["cloud_w0"] <= 0] = 1e-10
            DTAU[DTAU <= 0] = 1e-10
            COSB = np.zeros(DTAU.shape) + atm["cloud_g0"]
            W0 = np.zeros(DTAU.shape) + atm["cloud_w0"]
            W0_no_raman = W0
            TAU = np.zeros((L + 1, W))
            TAU[1:] = np.cumsum(DTAU, axis=0)
        if delta_eddington:
            f = COSB ** stream
            w0_d = W0 * (1. - f) / (1.0 - W0 * f)
            cosb_d = (COSB - f) / (1. - f)
            dtau_d = DTAU * (1. - W0 * f)
            tau_d = np.zeros((L + 1, W))
            tau_d[1:] = np.cumsum(dtau_d, axis=0)
            return (dtau_d, tau_d, w0_d, cosb_d, ftau_cld, ftau_ray, GCOS2, DTAU, TAU, W0, COSB,
                    W0_no_raman, f)
    return (DTAU, TAU, W0, COSB, ftau_cld, ftau_ray, GCOS2, DTAU, TAU, W0, COSB, W0_no_raman, 0 * COSB)


# ---- correlated-k (RetrieveCKs, pre-mixed tables) ---------------------------------------------------
def ck_find_pts(pressures, temps, nc_p, tlayer, player_bar):
    """index / weight logic of RetrieveCKs.get_pre_mix_ck (optics.py:1086-1149)."""
    t_inv = 1 / np.asarray(tlayer, dtype=np.float64)
    p_log = np.log10(np.asarray(player_bar, dtype=np.float64))
    p_grid = np.unique(pressures)
    p_log_grid = np.log10(p_grid[p_grid > 0])
    t_inv_grid = 1 / np.array(np.unique(temps))
    L = t_inv.size
    t_low = np.zeros(L, dtype=np.int64)
    for i in range(L):
        find = np.where(t_inv_grid > t_inv[i])[0]
        t_low[i] = 0 if len(find) == 0 else find[-1]
    t_low[t_low == (len(t_inv_grid) - 1)] = len(t_inv_grid) - 2
    t_hi = t_low + 1
    p_low = np.zeros(L, dtype=np.int64)
    for i in range(L):
        find = np.where(p_log_grid <= p_log[i])[0]
        p_low[i] = 0 if len(find) == 0 else find[-1]
        p_low[i] = min(p_low[i], nc_p[t_hi[i]] - 3)
    p_hi = p_low + 1
    t_interp = (t_inv - t_inv_grid[t_low]) / (t_inv_grid[t_hi] - t_inv_grid[t_low])
    p_interp = (p_log - p_log_grid[p_low]) / (p_log_grid[p_hi] - p_log_grid[p_low])
    return t_interp, p_interp, p_low, t_low, p_hi, t_hi


def premix_ck(ln_kappa, t_interp, p_interp, p_low, t_low, p_hi, t_hi):
    """optics.py:1151-1161: ln_kappa [nP, nT, W, K] -> molecular_opa [L, W, K]."""
    t = np.asarray(t_interp)[:, None, None]
    p = np.asarray(p_interp)[:, None, None]
    out = np.exp(((1 - t) * (1 - p) * ln_kappa[p_low, t_low, :, :]) + ((t) * (1 - p) * ln_kappa[p_low, t_hi, :, :]) +
                 ((t) * (p) * ln_kappa[p_hi, t_hi, :, :]) + ((1 - t) * (p) * ln_kappa[p_hi, t_low, :, :]))
    return out * N_A


def continuum_loglinear(cia_temps, table, tlayer):
    """RetrieveCKs.get_continuum (optics.py:1410-1497) for one CIA pair: table [nTc, W] in the order of
    cia_temps -> [L, W]; returns also the (low, high) row indices into the ascending-sorted table and t."""
    order = np.argsort(cia_temps)
    st = np.asarray(cia_temps, dtype=np.float64)[order]
    tab = np.asarray(table)[order]
    L = len(tlayer)
    lo = np.zeros(L, dtype=np.int64)
    hi = np.zeros(L, dtype=np.int64)
    for i, t in enumerate(tlayer):
        if t <= st[0]:
            lo[i], hi[i] = 0, 1
        elif t >= st[-1]:
            lo[i], hi[i] = len(st) - 2, len(st) - 1
        else:
            lo[i] = np.where(st - t <= 0)[0][-1]
            hi[i] = np.where(st - t > 0)[0][0]
    t_inv = 1 / np.asarray(tlayer, dtype=np.float64)
    ti = (t_inv - 1 / st[lo]) / (1 / st[hi] - 1 / st[lo])
    out = np.exp(((1 - ti)[:, None] * np.log(tab[lo])) + ((ti)[:, None] * np.log(tab[hi])))
    return out, lo, hi, ti


def compute_opacity_ck(atm, molecular_opa, continuum_opa, rayleigh_opa, stream=2, delta_eddington=True):
    """optics.py:147-431 for ngauss > 1 (pre-mixed CK), raman = 2, test_mode = None: molecular_opa
    [L, W, K] (already x N_A); returns the 13-tuple of [L|V, W, K] arrays."""
    L, W, K = molecular_opa.shape
    # continuum, Rayleigh and cloud terms are identical for every gauss point (optics.py:238, :277, :309-312):
    # run the monochromatic assembly once without molecules / Rayleigh / cloud to get the continuum sum
    bare = dict(atm, cloud_opd=np.zeros((L, W)), cloud_w0=np.zeros((L, W)), cloud_g0=np.zeros((L, W)))
    zero_ray = {m: np.zeros(W) for m in rayleigh_opa}
    with np.errstate(all="ignore"):
        TAUGAS0 = compute_opacity(bare, {}, continuum_opa, zero_ray, None, stream=stream, delta_eddington=False)[7]
    colden = atm["colden"][:, None, None]
    mmw = atm["mmw"][:, None, None]
    TAURAY = np.zeros((L, W))
    for m, sig in rayleigh_opa.items():
        TAURAY += np.array([sig] * L) * (atm["colden"][:, None] * atm["mixingratios"][m][:, None] / atm["mmw"][:, None])
    TAUCLD = atm["cloud_opd"]
    TAUGAS = TAUGAS0[:, :, None] + molecular_opa * (colden / mmw)
    TAURAY = np.repeat(TAURAY[:, :, None], K, axis=2)
    TAUCLD = np.repeat(TAUCLD[:, :, None], K, axis=2)
    w0c = np.repeat(atm["cloud_w0"][:, :, None], K, axis=2)
    g0 = np.repeat(atm["cloud_g0"][:, :, None], K, axis=2)
    rf = 0.99999
    with np.errstate(all="ignore"):
        DTAU = TAUGAS + TAURAY + TAUCLD
        ftau_cld = (w0c * TAUCLD) / (w0c * TAUCLD + TAURAY)
        ftau_ray = TAURAY / (TAURAY + w0c * TAUCLD)
        GCOS2 = 0.5 * ftau_ray
        W0 = (TAURAY * rf + TAUCLD * w0c) / (TAUGAS + TAURAY + TAUCLD)
        W0nr = (TAURAY * 0.99999 + TAUCLD * w0c) / (TAUGAS + TAURAY + TAUCLD)
        TAU = np.zeros((L + 1, W, K))
        TAU[1:] = np.cumsum(DTAU, axis=0)
        COSB = g0
        if delta_eddington:
            f = COSB ** stream
            w0_d = W0 * (1. - f) / (1.0 - W0 * f)
            cosb_d = (COSB - f) / (1. - f)
            dtau_d = DTAU * (1. - W0 * f)
            tau_d = np.zeros((L + 1, W, K))
            tau_d[1:] = np.cumsum(dtau_d, axis=0)
            return (dtau_d, tau_d, w0_d, cosb_d, ftau_cld, ftau_ray, GCOS2, DTAU, TAU, W0, COSB, W0nr, f)
    return (DTAU, TAU, W0, COSB, ftau_cld, ftau_ray, GCOS2, DTAU, TAU, W0, COSB, W0nr, 0 * COSB)
