"""Numerical study for DESIGN.md section 7 item 5 (no GPU): can the Toon tridiagonal be solved layer-parallel?

Builds the reference's 2L x 2L system (setup_tri_diag, fluxes.py:139-183) for the reflected golden cases in
numpy, then compares, against an mpmath (50 digits) Thomas solve:
  thomas   - the reference's bottom-up Thomas sweep in fp64 (tri_diag_solve, fluxes.py:289-323)
  scan     - the same recurrences evaluated as a TREE-ordered (Blelloch-style) suffix product of the 3 x 3
             homogeneous matrices  v_i = M_i v_{i+1},  (AS_i, DS_i) = (v_i[0], v_i[1]) / v_i[2],  with
             renormalisation of every partial product, followed by the affine forward substitution as a scan
  pcr      - plain parallel cyclic reduction on (A, B, C, D)
Reported: max over wavelengths of  max_i |X_i - X_exact| / max_i |X_exact|.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases as C  # noqa: E402


def system(d, kw, ig=0):
    """A, B, C, D [2L, W] of angle ig (numpy restatement of fluxes.py:1132-1205 + setup_tri_diag)"""
    L = d["nlevel"] - 1
    W = d["nwno"]
    dtau, tau, w0, cosb, fc = d["dtau"], d["tau"], d["w0"], d["cosb"], d["ftau_cld"]
    surf = np.broadcast_to(np.asarray(d["surf_reflect"], float), (W,))
    F0 = d["F0PI"]
    u0 = d["ubar0"][ig, 0]
    sq3 = np.sqrt(3.0)
    g = fc * cosb
    if kw["toon_coefficients"] == 1:
        g1 = (7 - w0 * (4 + 3 * g)) / 4
        g2 = -(1 - w0 * (4 - 3 * g)) / 4
        g3 = (2 - 3 * g * u0) / 4
    else:
        g1 = (sq3 * 0.5) * (2.0 - w0 * (1.0 + g))
        g2 = (sq3 * w0 * 0.5) * (1.0 - g)
        g3 = 0.5 * (1.0 - sq3 * g * u0)
    lam = np.sqrt(g1 ** 2 - g2 ** 2)
    gam = (g1 - lam) / g2
    g4 = 1 - g3
    den = lam ** 2 - 1 / u0 ** 2
    am = F0 * w0 * (g4 * (g1 + 1 / u0) + g2 * g3) / den
    ap = F0 * w0 * (g3 * (g1 - 1 / u0) + g2 * g4) / den
    xu, xd = np.exp(-tau[:-1] / u0), np.exp(-tau[1:] / u0)
    cmu, cpu, cmd, cpd = am * xu, ap * xu, am * xd, ap * xd
    EP = np.exp(np.minimum(lam * dtau, 35.0))
    EM = 1 / EP
    e1, e2, e3, e4 = EP + gam * EM, EP - gam * EM, gam * EP + EM, gam * EP - EM
    bs = surf * u0 * F0 * np.exp(-tau[-1] / u0)
    A = np.zeros((2 * L, W)); B = np.zeros((2 * L, W)); Cc = np.zeros((2 * L, W)); D = np.zeros((2 * L, W))
    B[0] = gam[0] + 1; Cc[0] = gam[0] - 1; D[0] = 0.0 - cmu[0]
    A[1::2][:-1] = (e1[:-1] + e3[:-1]) * (gam[1:] - 1); B[1::2][:-1] = (e2[:-1] + e4[:-1]) * (gam[1:] - 1)
    Cc[1::2][:-1] = 2 * (1 - gam[1:] ** 2)
    D[1::2][:-1] = (gam[1:] - 1) * (cpu[1:] - cpd[:-1]) + (1 - gam[1:]) * (cmd[:-1] - cmu[1:])
    A[::2][1:] = 2 * (1 - gam[:-1] ** 2); B[::2][1:] = (e1[:-1] - e3[:-1]) * (gam[1:] + 1)
    Cc[::2][1:] = (e1[:-1] + e3[:-1]) * (gam[1:] - 1)
    D[::2][1:] = e3[:-1] * (cpu[1:] - cpd[:-1]) + e1[:-1] * (cmd[:-1] - cmu[1:])
    A[-1] = e1[-1] - surf * e3[-1]; B[-1] = e2[-1] - surf * e4[-1]; D[-1] = bs - cpd[-1] + surf * cmd[-1]
    return A, B, Cc, D


def thomas(A, B, Cc, D, mp=None):
    n = len(A)
    one = (mp.mpf(1) if mp else 1.0)
    AS = [None] * n; DS = [None] * n
    AS[-1] = A[-1] / B[-1]; DS[-1] = D[-1] / B[-1]
    for i in range(n - 2, -1, -1):
        x = one / (B[i] - Cc[i] * AS[i + 1])
        AS[i] = A[i] * x
        DS[i] = (D[i] - Cc[i] * DS[i + 1]) * x
    X = [None] * n
    X[0] = DS[0]
    for i in range(1, n):
        X[i] = DS[i] - AS[i] * X[i - 1]
    return X


def _affine_scan(m, t, reverse=False):
    """x_i = m_i x_{i-1} + t_i (x_{-1} = 0) for all i by recursive doubling; reverse: x_i = m_i x_{i+1} + t_i"""
    if reverse:
        return _affine_scan(m[::-1], t[::-1])[::-1]
    m, t = m.copy(), t.copy()
    n = m.shape[0]
    m[0] = 0.0
    step = 1
    while step < n:
        m2, t2 = m.copy(), t.copy()
        t2[step:] = m[step:] * t[:-step] + t[step:]
        m2[step:] = m[step:] * m[:-step]
        m, t = m2, t2
        step *= 2
    return t


def scan_solve(A, B, Cc, D, split=True, surface_divide=False):
    """The reference's bottom-up Thomas recurrences evaluated as tree-ordered scans.
    split=False: one 3 x 3 homogeneous (Moebius + affine) suffix product for (AS, DS) - DS loses RELATIVE
    accuracy where it is tiny against AS (deep layers), which the e^{35}-scaled level fluxes cannot tolerate.
    split=True : 2 x 2 Moebius suffix product for AS only (both components O(1)), then DS and X as affine scans
    whose partial compositions carry their own magnitude."""
    n, W = A.shape

    def norm(P):
        return P / np.max(np.abs(P), axis=(-1, -2), keepdims=True)
    if not split:
        M = np.zeros((n, W, 3, 3))
        M[:, :, 0, 2] = A; M[:, :, 1, 1] = -Cc; M[:, :, 1, 2] = D; M[:, :, 2, 0] = -Cc; M[:, :, 2, 2] = B
    else:
        # AS_i = a_i / (b_i - c_i AS_{i+1}):  (p, r)_i = [[0, a_i], [-c_i, b_i]] (p, r)_{i+1}
        M = np.zeros((n, W, 2, 2))
        M[:, :, 0, 1] = A; M[:, :, 1, 0] = -Cc; M[:, :, 1, 1] = B
    S = norm(M.copy())
    step = 1
    while step < n:
        S2 = S.copy()
        S2[:n - step] = norm(np.einsum("nwij,nwjk->nwik", S[:n - step], S[step:]))
        S = S2
        step *= 2
    v = S[:, :, :, -1]                    # S_i e_last
    if not split:
        AS, DS = v[:, :, 0] / v[:, :, 2], v[:, :, 1] / v[:, :, 2]
    else:
        AS = v[:, :, 0] / v[:, :, 1]
        ASn = np.vstack([AS[1:], np.zeros((1, W))])
        x = 1.0 / (B - Cc * ASn)          # the Thomas pivots, all rows at once
        DS = _affine_scan(-Cc * x, D * x, reverse=True)   # DS_i = (d_i - c_i DS_{i+1}) x_i
        if surface_divide:                # the reference DIVIDES in the surface row (fluxes.py:305): AS = a / b
            AS[-1] = A[-1] / B[-1]
            DS[-1] = D[-1] / B[-1]
    X = _affine_scan(-AS, DS)             # X_i = DS_i - AS_i X_{i-1}
    if split:
        # Y+ = X[2l] + X[2l+1] cancels to ~1e-30 of its terms in optically thick layers and is multiplied by
        # e^{35} in the level fluxes: the pair must carry the reference's own rounding relation, so the odd
        # entries are re-derived from their even neighbours with the reference's (local) substitution step
        X[1::2] = DS[1::2] - AS[1::2] * X[0::2]
    return X


def pcr(A, B, Cc, D):
    a, b, c, d = A.copy(), B.copy(), Cc.copy(), D.copy()
    n = a.shape[0]
    step = 1
    with np.errstate(all="ignore"):
        while step < n:
            an, bn, cn, dn = a.copy(), b.copy(), c.copy(), d.copy()
            for i in range(n):
                lo, hi = i - step, i + step
                al = a[i] / b[lo] if lo >= 0 else 0 * a[i]
                ga = c[i] / b[hi] if hi < n else 0 * c[i]
                bn[i] = b[i] - (al * c[lo] if lo >= 0 else 0) - (ga * a[hi] if hi < n else 0)
                dn[i] = d[i] - (al * d[lo] if lo >= 0 else 0) - (ga * d[hi] if hi < n else 0)
                an[i] = -al * a[lo] if lo >= 0 else 0 * a[i]
                cn[i] = -ga * c[hi] if hi < n else 0 * c[i]
            a, b, c, d = an, bn, cn, dn
            step *= 2
        return d / b


def reflected_levels_from_scan(d, kw, ig=0):
    """all four level arrays of get_reflected_1d(get_lvl_flux=1) (fluxes.py:1219-1257) with X from the scan:
    once X is known every level is independent (no further recurrence)"""
    A, B, Cc, D = system(d, kw, ig)
    X = scan_solve(A, B, Cc, D)
    L = d["nlevel"] - 1
    pos, neg = X[::2] + X[1::2], X[::2] - X[1::2]
    # recompute the per-layer quantities (as system())
    W = d["nwno"]
    dtau, tau, w0, cosb, fc = d["dtau"], d["tau"], d["w0"], d["cosb"], d["ftau_cld"]
    F0 = d["F0PI"]; u0 = d["ubar0"][ig, 0]
    sq3 = np.sqrt(3.0); g = fc * cosb
    if kw["toon_coefficients"] == 1:
        g1 = (7 - w0 * (4 + 3 * g)) / 4; g2 = -(1 - w0 * (4 - 3 * g)) / 4; g3 = (2 - 3 * g * u0) / 4
    else:
        g1 = (sq3 * 0.5) * (2.0 - w0 * (1.0 + g)); g2 = (sq3 * w0 * 0.5) * (1.0 - g); g3 = 0.5 * (1.0 - sq3 * g * u0)
    lam = np.sqrt(g1 ** 2 - g2 ** 2); gam = (g1 - lam) / g2; g4 = 1 - g3
    den = lam ** 2 - 1 / u0 ** 2
    am = F0 * w0 * (g4 * (g1 + 1 / u0) + g2 * g3) / den
    ap = F0 * w0 * (g3 * (g1 - 1 / u0) + g2 * g4) / den
    E = np.minimum(lam * dtau, 35.0)
    fm = np.zeros((L + 1, W)); fp = np.zeros((L + 1, W)); fmm = np.zeros((L + 1, W)); fpm = np.zeros((L + 1, W))
    xu = np.exp(-tau[:-1] / u0)
    fm[:-1] = pos * gam + neg + am * xu + u0 * F0 * xu
    fp[:-1] = pos + gam * neg + ap * xu
    xdn = np.exp(-tau[-1] / u0)
    fm[-1] = gam[-1] * pos[-1] * np.exp(E[-1]) + neg[-1] / np.exp(E[-1]) + am[-1] * xdn + u0 * F0 * xdn
    fp[-1] = pos[-1] * np.exp(E[-1]) + gam[-1] * neg[-1] / np.exp(E[-1]) + ap[-1] * xdn
    xm = np.exp(-(tau[:-1] + 0.5 * dtau) / u0)
    EPm = np.exp(0.5 * E)
    fmm[:-1] = gam * pos * EPm + neg / EPm + am * xm + u0 * F0 * xm
    fpm[:-1] = pos * EPm + gam * neg / EPm + ap * xm
    return fm, fp, fmm, fpm


def main():
    import mpmath as mp
    mp.mp.dps = 50
    names = ["refl_cfg1_tthg_ray", "refl_adversarial", "refl_lvl", "refl_combo_sp1_mp0_tc1"]
    print("%-28s %12s %12s %12s %12s" % ("case", "thomas", "scan 3x3", "scan split", "pcr"))
    for name in names:
        case = C.reflected_cases()[name]
        d = C.build_reflected(case)
        A, B, Cc, D = system(d, case["kw"])
        W = min(A.shape[1], 24)
        A, B, Cc, D = A[:, :W], B[:, :W], Cc[:, :W], D[:, :W]
        n = A.shape[0]
        exact = np.zeros((n, W))
        for w in range(W):
            X = thomas([mp.mpf(x) for x in A[:, w]], [mp.mpf(x) for x in B[:, w]], [mp.mpf(x) for x in Cc[:, w]],
                       [mp.mpf(x) for x in D[:, w]], mp=mp)
            exact[:, w] = [float(x) for x in X]
        scale = np.max(np.abs(exact), axis=0)
        err = lambda X: float(np.nanmax(np.max(np.abs(np.asarray(X, dtype=float) - exact), axis=0) / scale)) \
            if np.all(np.isfinite(np.asarray(X, dtype=float))) else float("nan")
        Xt = np.array(thomas(A, B, Cc, D))
        print("%-28s %12.2e %12.2e %12.2e %12.2e" % (name, err(Xt), err(scan_solve(A, B, Cc, D, split=False)),
                                                     err(scan_solve(A, B, Cc, D)), err(pcr(A, B, Cc, D))))
    # level arrays through the scan against the reference golden vectors, with the tests' own criterion
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from util import assert_level_close_yardstick, golden
    g = golden("reflected")
    for name in ("refl_lvl", "refl_lvl_edd"):
        case = C.reflected_cases()[name]
        d = C.build_reflected(case)
        args = C.reflected_args(d, case["kw"])
        _, o64 = oracle.get_reflected_1d(*args)
        _, q = oracle.get_reflected_1d(*args, quad=True, nthreads=8)
        for ig in range(d["numg"]):
            lv = reflected_levels_from_scan(d, case["kw"], ig)
            for k, a, o, x in zip(("fm", "fp", "fmm", "fpm"), lv, o64, q):
                assert_level_close_yardstick(a, g[name + "/" + k][ig, 0], x[ig, 0], what=name + " " + k)
        print("%-28s level arrays from the scan pass the level-flux (yardstick) criterion of tests/util.py" % name)


if __name__ == "__main__" and "--thermal" not in sys.argv:
    main()


# ------------------------------------------------------------------------------------------------------
# Thermal (get_thermal_1d, fluxes.py:1746-1912): the same split scan for X, then the downward / upward
# source-function recurrences (:1875-1907) as affine scans over layers.
# ------------------------------------------------------------------------------------------------------
def thermal_levels_from_scan(d, ia=0):
    h, c, k = 6.62607004e-27, 2.99792458e+10, 1.38064852e-16
    L = d["nlevel"] - 1
    W = d["nwno"]
    dt, om, g = d["dtau"], d["w0"], d["cosb"]
    wl = 1.0 / d["wno"]
    B = ((2.0 * h * c * c) / wl ** 5.0)[None, :] * (1.0 / (np.exp((h * c) / (wl[None, :] * k) / d["tlevel"][:, None]) - 1.0))
    mu1 = 0.5
    u = d["ubar1"].reshape(-1)[ia]
    surf = d["surf_reflect"]
    b0 = B[:-1]
    b1 = (B[1:] - B[:-1]) / dt
    g1 = 2.0 - om * (1 + g)
    g2 = om * (1 - g)
    lam = np.sqrt(g1 * g1 - g2 * g2)
    gam = (g1 - lam) / g2
    q = 1.0 / (g1 + g2)
    tp = 2 * np.pi * mu1
    cpu, cmu = tp * (b0 + b1 * q), tp * (b0 - b1 * q)
    cpd, cmd = tp * (b0 + b1 * dt + b1 * q), tp * (b0 + b1 * dt - b1 * q)
    E = np.minimum(lam * dt, 35.0)
    EP = np.exp(E); EM = 1 / EP
    e1, e2, e3, e4 = EP + gam * EM, EP - gam * EM, gam * EP + EM, gam * EP - EM
    pl = d["plevel"]
    tau_top = dt[0] * pl[0] / (pl[1] - pl[0])
    b_top = (1.0 - np.exp(-tau_top / mu1)) * B[0] * np.pi
    if d["hard_surface"]:
        b_surface = (1.0 - surf) * B[-1] * np.pi
    else:
        b_surface = (B[-1] + b1[-1] * mu1) * np.pi
    A = np.zeros((2 * L, W)); Bm = np.zeros((2 * L, W)); Cc = np.zeros((2 * L, W)); D = np.zeros((2 * L, W))
    Bm[0] = gam[0] + 1; Cc[0] = gam[0] - 1; D[0] = b_top - cmu[0]
    A[1::2][:-1] = (e1[:-1] + e3[:-1]) * (gam[1:] - 1); Bm[1::2][:-1] = (e2[:-1] + e4[:-1]) * (gam[1:] - 1)
    Cc[1::2][:-1] = 2 * (1 - gam[1:] ** 2)
    D[1::2][:-1] = (gam[1:] - 1) * (cpu[1:] - cpd[:-1]) + (1 - gam[1:]) * (cmd[:-1] - cmu[1:])
    A[::2][1:] = 2 * (1 - gam[:-1] ** 2); Bm[::2][1:] = (e1[:-1] - e3[:-1]) * (gam[1:] + 1)
    Cc[::2][1:] = (e1[:-1] + e3[:-1]) * (gam[1:] - 1)
    D[::2][1:] = e3[:-1] * (cpu[1:] - cpd[:-1]) + e1[:-1] * (cmd[:-1] - cmu[1:])
    A[-1] = e1[-1] - surf * e3[-1]; Bm[-1] = e2[-1] - surf * e4[-1]; D[-1] = b_surface - cpd[-1] + surf * cmd[-1]
    X = scan_solve(A, Bm, Cc, D, surface_divide=True)
    pos, neg = X[::2] + X[1::2], X[::2] - X[1::2]
    xa, xh = np.exp(-dt / u), np.exp(-0.5 * dt / u)
    EPh = np.exp(0.5 * E); EMh = 1 / EPh
    lu = lam * u
    # downward: f_{l+1} = xa f_l + src_l  (fluxes.py:1883-1888), prefix scan
    J = gam * (lam + 1 / mu1) * pos
    K = (1 / mu1 - lam) * neg
    si1 = 2 * np.pi * (b0 - b1 * (q - mu1)); si2 = 2 * np.pi * b1
    src_dn = (J / (lu + 1.0)) * (EP - xa) + (K / (lu - 1.0)) * (xa - EM) + si1 * (1. - xa) + si2 * (u * xa + dt - u)
    f0 = (1 - np.exp(-tau_top / u)) * B[0] * 2 * np.pi
    t = src_dn.copy(); t[0] = xa[0] * f0 + src_dn[0]
    fnext = _affine_scan(xa, t)                              # f at levels 1..L
    fm = np.vstack([f0[None, :], fnext])
    fmm = np.zeros_like(fm)
    fmm[:-1] = fm[:-1] * xh + (J / (lu + 1.0)) * (EPh - xh) + (K / (-lu + 1.0)) * (EMh - xh) + si1 * (1. - xh) + \
        si2 * (u * xh + 0.5 * dt - u)
    # upward: f_l = xa f_{l+1} + src_l  (fluxes.py:1897-1901), suffix scan
    Gt = (1 / mu1 - lam) * pos
    Ht = gam * (lam + 1 / mu1) * neg
    al1 = 2 * np.pi * (b0 + b1 * (q - mu1)); al2 = 2 * np.pi * b1
    src_up = (Gt / (lu - 1.0)) * (EP * xa - 1.0) + (Ht / (lu + 1.0)) * (1.0 - EM * xa) + al1 * (1. - xa) + al2 * (u - (dt + u) * xa)
    fL = (1.0 - surf) * B[-1] * 2 * np.pi if d["hard_surface"] else (B[-1] + b1[-1] * u) * 2 * np.pi
    t = src_up.copy(); t[-1] = xa[-1] * fL + src_up[-1]
    fup = _affine_scan(xa, t, reverse=True)                  # f at levels 0..L-1
    fp = np.vstack([fup, fL[None, :]])
    fpm = np.zeros_like(fp)
    fpm[:-1] = fp[1:] * xh + (Gt / (lu - 1.0)) * (EP * xh - EPh) - (Ht / (lu + 1.0)) * (EM * xh - EMh) + al1 * (1. - xh) + \
        al2 * (u + 0.5 * dt - (dt + u) * xh)
    return fm, fp, fmm, fpm


def thermal_main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from util import assert_level_close_yardstick, golden
    g = golden("thermal")
    for name in ("therm_ct0_hs0", "therm_ct0_hs1", "therm_cfg2_small", "therm_cold"):
        case = C.thermal_cases()[name]
        d = C.build_thermal(case)
        args = C.thermal_args(d)
        _, o64 = oracle.get_thermal_1d(*args, nthreads=8)
        _, q = oracle.get_thermal_1d(*args, quad=True, nthreads=8)
        for ia in range(d["numg"]):
            lv = thermal_levels_from_scan(d, ia)
            for k, a, o, x in zip(("fm", "fp", "fmm", "fpm"), lv, o64, q):
                assert_level_close_yardstick(a, o[ia, 0], x[ia, 0], what=name + " " + k + " angle %d" % ia)
        print("%-28s thermal level arrays from the scans pass the level-flux (yardstick) criterion" % name)


if __name__ == "__main__" and "--thermal" in sys.argv:
    thermal_main()
