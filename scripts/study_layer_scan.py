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


def scan_solve(A, B, Cc, D):
    """tree-ordered suffix products of M_i = [[0,0,a],[0,-c,d],[-c,0,b]], renormalised; then affine scan"""
    n, W = A.shape
    M = np.zeros((n, W, 3, 3))
    M[:, :, 0, 2] = A; M[:, :, 1, 1] = -Cc; M[:, :, 1, 2] = D; M[:, :, 2, 0] = -Cc; M[:, :, 2, 2] = B

    def norm(P):
        return P / np.max(np.abs(P), axis=(-1, -2), keepdims=True)
    # inclusive suffix scan by recursive doubling (Hillis-Steele): S_i = M_i M_{i+1} ... M_{n-1}
    S = norm(M.copy())
    step = 1
    while step < n:
        S2 = S.copy()
        S2[:n - step] = norm(np.einsum("nwij,nwjk->nwik", S[:n - step], S[step:]))
        S = S2
        step *= 2
    v = S[:, :, :, 2]                     # S_i e3
    AS, DS = v[:, :, 0] / v[:, :, 2], v[:, :, 1] / v[:, :, 2]
    # X_i = DS_i - AS_i X_{i-1}: affine maps (m, t): x -> m x + t, prefix composition by recursive doubling
    m, t = -AS.copy(), DS.copy()
    m[0] = 0.0
    step = 1
    while step < n:
        m2, t2 = m.copy(), t.copy()
        t2[step:] = m[step:] * t[:-step] + t[step:]
        m2[step:] = m[step:] * m[:-step]
        m, t = m2, t2
        step *= 2
    return t


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


def main():
    import mpmath as mp
    mp.mp.dps = 50
    names = ["refl_cfg1_tthg_ray", "refl_adversarial", "refl_lvl", "refl_combo_sp1_mp0_tc1"]
    print("%-28s %12s %12s %12s" % ("case", "thomas", "scan", "pcr"))
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
        print("%-28s %12.2e %12.2e %12.2e" % (name, err(Xt), err(scan_solve(A, B, Cc, D)), err(pcr(A, B, Cc, D))))


if __name__ == "__main__":
    main()
