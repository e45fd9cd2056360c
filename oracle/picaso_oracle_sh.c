/*
 * picaso_oracle_sh.c - CPU restatement of the reference's spherical-harmonics reflected
 * solver (P1 / P3, "SH2" / "SH4").  TEST INFRASTRUCTURE ONLY - see picaso_oracle.h.
 *
 * Follows /root/reference/picaso/fluxes.py (commit 0369089):
 *   get_reflected_SH        :2796-2974
 *   setup_2_stream_fluxes   :3239-3309
 *   setup_4_stream_fluxes   :3387-3607
 *   solve_4_stream_banded   :3610-3628  -> scipy.linalg.solve_banded -> LAPACK dgbsv.
 * scipy/LAPACK are third-party and not vendored in the reference (pyproject.toml:26 pins
 * only "scipy"); the banded solve is restated here as the published dgbtf2/dgbtrs algorithm
 * (unblocked LU with partial pivoting in band storage), anchored on the reference call site.
 * The reference's in-place drift of f_deltaM across angles (fluxes.py:2823-2824, SURVEY.md
 * Appendix A1) is reproduced: `f_deltaM` is modified exactly as the reference modifies it.
 */
#include "picaso_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double f64;
#ifdef ORC_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_EXP expq
#define R_SQRT sqrtq
#define R_POW powq
#define R_FABS fabsq
#define ORC_NAME(x) x##_quad
#else
typedef double real;
#define R_EXP exp
#define R_SQRT sqrt
#define R_POW pow
#define R_FABS fabs
#define ORC_NAME(x) x
#endif

#define PI 3.14159265358979323846

/* slice_rav, fluxes.py:69-76 */
static inline real clip35(real x) { return x > 35.0 ? (real)35.0 : (x < -35.0 ? (real)-35.0 : x); }

/* legP, fluxes.py:3643-3646 (first four) */
static void legp(real mu, real *P)
{
    P[0] = 1;
    P[1] = mu;
    P[2] = (3 * mu * mu - 1) / 2;
    P[3] = (5 * mu * mu * mu - 3 * mu) / 2;
}

/* LAPACK dgbsv (dgbtf2 + dgbtrs, no transpose, one right-hand side) with kl = ku = k.
 * ab is (2k + k + 1) x n in LAPACK band storage, column major: ab[(kl+ku+i-j) + j*ldab]. */
static void band_solve(int n, int k, real *ab, int ldab, int *ipiv, real *b)
{
    const int kv = 2 * k; /* ku + kl: super-diagonals of U after pivoting */
    int ju = 0;
    for (int j = 0; j < n; ++j) {
        int km = (k < n - 1 - j) ? k : n - 1 - j;
        /* pivot: max |a(i,j)|, i = j..j+km */
        int jp = 0;
        real amax = R_FABS(ab[kv + j * ldab]);
        for (int i = 1; i <= km; ++i) {
            real v = R_FABS(ab[kv + i + j * ldab]);
            if (v > amax) { amax = v; jp = i; }
        }
        ipiv[j] = j + jp;
        int t = j + k + jp;
        if (t > n - 1) t = n - 1;
        if (t > ju) ju = t;
        if (jp != 0) /* swap rows j and j+jp over columns j..ju */
            for (int c = j; c <= ju; ++c) {
                real *p1 = &ab[kv + jp + j - c + c * ldab], *p2 = &ab[kv + j - c + c * ldab];
                real tmp = *p1; *p1 = *p2; *p2 = tmp;
            }
        real piv = ab[kv + j * ldab];
        if (km > 0) {
            real inv = 1 / piv;
            for (int i = 1; i <= km; ++i) ab[kv + i + j * ldab] *= inv;
            for (int c = j + 1; c <= ju; ++c) {
                real ujc = ab[kv + j - c + c * ldab];
                for (int i = 1; i <= km; ++i)
                    ab[kv + i + j - c + c * ldab] -= ab[kv + i + j * ldab] * ujc;
            }
        }
    }
    /* forward: apply P and L */
    for (int j = 0; j < n - 1; ++j) {
        int km = (k < n - 1 - j) ? k : n - 1 - j;
        int l = ipiv[j];
        if (l != j) { real tmp = b[l]; b[l] = b[j]; b[j] = tmp; }
        for (int i = 1; i <= km; ++i) b[j + i] -= ab[kv + i + j * ldab] * b[j];
    }
    /* back substitution with U (bandwidth kv) */
    for (int j = n - 1; j >= 0; --j) {
        b[j] /= ab[kv + j * ldab];
        int lo = j - kv < 0 ? 0 : j - kv;
        for (int i = lo; i < j; ++i) b[i] -= ab[kv + i - j + j * ldab] * b[j];
    }
}

#define MB(d, j) ab[(k + (d)) + (size_t)(j) * ldab] /* reference Mb[d, j] -> LAPACK row kl + d */

void ORC_NAME(orc_get_reflected_SH)(
    int nlevel, int nwno, int numg, int numt,
    const f64 *dtau, const f64 *tau, const f64 *w0, const f64 *cosb, const f64 *ftau_cld,
    const f64 *ftau_ray, f64 *f_deltaM /* modified like the reference does */,
    const f64 *dtau_og, const f64 *tau_og, const f64 *w0_og, const f64 *cosb_og,
    const f64 *surf_reflect, const f64 *ubar0, const f64 *ubar1, f64 cos_theta, const f64 *F0PI,
    int w_single_form, int w_multi_form, int psingle_form, int w_single_rayleigh,
    int w_multi_rayleigh, int psingle_rayleigh,
    f64 frac_a, f64 frac_b, f64 frac_c, f64 constant_back, f64 constant_forward,
    int stream, const f64 *b_top, int single_form, f64 *xint_at_top,
    f64 *flux /* flx = 1: [G][S * nlevel][W] layer fluxes F.X + G (fluxes.py:2889-2890, :3311-3331, :3551-3598); NULL: flx = 0 */,
    int nthreads)
{
    const int L = nlevel - 1, W = nwno, G = numg * numt, S = stream;
    const int n = S * L, k = 3 * S / 2 - 1, ldab = 3 * k + 1;
    (void)cosb;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        real *ab = (real *)malloc(sizeof(real) * (size_t)ldab * n);
        real *B = (real *)malloc(sizeof(real) * (size_t)n);
        int *ipiv = (int *)malloc(sizeof(int) * (size_t)n);
        real *lay = (real *)malloc(sizeof(real) * (size_t)L * 64);
        real *a_ = lay, *b_ = a_ + 4 * L, *ws = b_ + 4 * L, *wm = ws + 4 * L, *eta = wm + 4 * L,
             *lam1 = eta + 4 * L, *lam2 = lam1 + L, *qv = lam2 + L, *ps = qv + L, *fd = ps + L,
             *Am = fd + L /* 16 L */, *X = Am + 16 * L /* 4 L */, *zz = X + 4 * L /* 8 L */;
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
#define LW(arr, l) ((real)(arr)[(size_t)(l) * W + w])
            const real f0 = F0PI[w], r = surf_reflect[w], bt = b_top ? b_top[w] : 0.0;
            for (int l = 0; l < L; ++l) fd[l] = LW(f_deltaM, l);
            for (int ai = 0; ai < G; ++ai) {
                const real u0 = ubar0[ai], u1 = ubar1[ai];
                real Pu0[4], Pu1[4];
                legp(-u0, Pu0);
                legp(u1, Pu1);
                /* phase-function moments, fluxes.py:2805-2840 */
                for (int l = 0; l < L; ++l) {
                    real g = LW(cosb_og, l);
                    for (int m = 0; m < 4; ++m) ws[m * L + l] = wm[m * L + l] = 1;
                    if (w_single_form == 1 || w_multi_form == 1)
                        for (int m = 1; m < S; ++m) {
                            real wv = (2 * m + 1) * R_POW(g, (real)m);
                            real v = (wv - (2 * m + 1) * fd[l]) / (1 - fd[l]);
                            if (w_single_form == 1) ws[m * L + l] = v;
                            if (w_multi_form == 1) wm[m * L + l] = v;
                        }
                    if (w_single_form == 0 || w_multi_form == 0) {
                        real gf = constant_forward * g, gb = constant_back * g;
                        real f = frac_a + frac_b * R_POW(gb, (real)frac_c);
                        fd[l] *= (f * R_POW((real)constant_forward, (real)S) +
                                  (1 - f) * R_POW((real)constant_back, (real)S));
                        for (int m = 1; m < S; ++m) {
                            real wv = (2 * m + 1) * (f * R_POW(gf, (real)m) + (1 - f) * R_POW(gb, (real)m));
                            real v = (wv - (2 * m + 1) * fd[l]) / (1 - fd[l]);
                            if (w_single_form == 0) ws[m * L + l] = v;
                            if (w_multi_form == 0) wm[m * L + l] = v;
                        }
                    }
                    if (w_single_rayleigh == 1) {
                        for (int m = 1; m < S; ++m) ws[m * L + l] *= LW(ftau_cld, l);
                        if (S == 4) ws[2 * L + l] += 0.5 * LW(ftau_ray, l);
                    }
                    if (w_multi_rayleigh == 1) {
                        for (int m = 1; m < S; ++m) wm[m * L + l] *= LW(ftau_cld, l);
                        if (S == 4) wm[2 * L + l] += 0.5 * LW(ftau_ray, l);
                    }
                    /* single scattering, fluxes.py:2843-2855, :2954-2957 */
                    real p = 0;
                    if (single_form == 0) {
                        if (psingle_form == 1) {
                            real s = R_SQRT(1 + g * g + 2 * g * cos_theta);
                            p = (1 - g * g) / (s * s * s);
                        } else {
                            real gf = constant_forward * g, gb = constant_back * g;
                            real f = frac_a + frac_b * R_POW(gb, (real)frac_c);
                            real tf = 1 + gf * gf + 2 * gf * cos_theta, tb = 1 + gb * gb + 2 * gb * cos_theta;
                            p = f * (1 - gf * gf) / R_SQRT(tf * tf * tf) +
                                (1 - f) * (1 - gb * gb) / R_SQRT(tb * tb * tb);
                        }
                        if (psingle_rayleigh == 1)
                            p = LW(ftau_cld, l) * p + LW(ftau_ray, l) * (0.75 * (1 + cos_theta * cos_theta));
                    } else {
                        for (int m = 0; m < S; ++m) p = p + ws[m * L + l] * Pu0[m] * Pu1[m];
                    }
                    ps[l] = p;
                    for (int m = 0; m < S; ++m) {
                        a_[m * L + l] = (2 * m + 1) - LW(w0, l) * wm[m * L + l];
                        b_[m * L + l] = (f0 * (LW(w0, l) * ws[m * L + l])) * Pu0[m] / (4 * PI);
                    }
                }
                const real b_surface = (0. + r * u0 * f0 * R_EXP(-LW(tau, L) / u0));
                const real b_surface_SH4 = -(0. + r * u0 * f0 * R_EXP(-LW(tau, L) / u0)) / 4;
                memset(ab, 0, sizeof(real) * (size_t)ldab * n);
                memset(B, 0, sizeof(real) * (size_t)n);
                real flux_bot_row[4] = {0, 0, 0, 0}, G_bot = 0;
                if (S == 2) {
                    /* setup_2_stream_fluxes, fluxes.py:3239-3309 */
                    real *Q1 = X, *Q2 = X + L, *em = Am, *zmu = Am + L, *zpu = Am + 2 * L,
                         *zmd = Am + 3 * L, *zpd = Am + 4 * L;
                    for (int l = 0; l < L; ++l) {
                        real a0 = a_[l], a1 = a_[L + l], b0 = b_[l], b1 = b_[L + l];
                        real Del = ((1 / u0) * (1 / u0) - a0 * a1);
                        eta[l] = (b1 / u0 - a1 * b0) / Del;
                        eta[L + l] = (b0 / u0 - a0 * b1) / Del;
                        lam1[l] = R_SQRT(a0 * a1);
                        em[l] = R_EXP(-clip35(lam1[l] * LW(dtau, l)));
                        qv[l] = lam1[l] / a1;
                        Q1[l] = (0.5 + qv[l]) * 2 * PI;
                        Q2[l] = (0.5 - qv[l]) * 2 * PI;
                        real zmn = (0.5 * eta[l] - eta[L + l]) * 2 * PI;
                        real zpl = (0.5 * eta[l] + eta[L + l]) * 2 * PI;
                        real et = R_EXP(-LW(tau, l) / u0), eb = R_EXP(-LW(tau, l + 1) / u0);
                        zmu[l] = zmn * eb; zpu[l] = zpl * eb; zmd[l] = zmn * et; zpd[l] = zpl * et;
                    }
                    MB(2, 0) = Q1[0];
                    MB(1, 1) = Q2[0];
                    B[0] = bt - zmd[0];
                    int nn = L - 1;
                    MB(3, 2 * L - 2) = Q2[nn] * em[nn] - r * (Q1[nn] * em[nn]);
                    MB(2, 2 * L - 1) = Q1[nn] / em[nn] - r * (Q2[nn] / em[nn]);
                    B[2 * L - 1] = b_surface - zpu[nn] + r * zmu[nn];
                    for (int kk = 0; kk < L - 1; ++kk) {
                        MB(0, 2 * kk + 3) = -Q2[kk + 1];
                        MB(1, 2 * kk + 2) = -Q1[kk + 1];
                        MB(1, 2 * kk + 3) = -Q1[kk + 1];
                        MB(2, 2 * kk + 1) = Q2[kk] / em[kk];
                        MB(2, 2 * kk + 2) = -Q2[kk + 1];
                        MB(3, 2 * kk) = Q1[kk] * em[kk];
                        MB(3, 2 * kk + 1) = Q1[kk] / em[kk];
                        MB(4, 2 * kk) = Q2[kk] * em[kk];
                        B[2 * kk + 1] = zmd[kk + 1] - zmu[kk];
                        B[2 * kk + 2] = zpd[kk + 1] - zpu[kk];
                    }
                    flux_bot_row[0] = Q2[nn] * em[nn];
                    flux_bot_row[1] = Q1[nn] / em[nn];
                    G_bot = zpu[nn];
                    band_solve(n, k, ab, ldab, ipiv, B);
                    if (flux) {
                        /* calculate_flux(F, G, X), F and G of fluxes.py:3311-3331 (dot products in index order) */
                        f64 *fo = flux + (size_t)ai * 2 * nlevel * W + w;
                        fo[0] = (f64)((Q1[0] * B[0] + Q2[0] * B[1]) + zmd[0]);
                        fo[(size_t)W] = (f64)((Q2[0] * B[0] + Q1[0] * B[1]) + zpd[0]);
                        for (int kk = 0; kk < L; ++kk) {
                            fo[(size_t)(2 * kk + 2) * W] = (f64)(((Q1[kk] * em[kk]) * B[2 * kk] + (Q2[kk] / em[kk]) * B[2 * kk + 1]) + zmu[kk]);
                            fo[(size_t)(2 * kk + 3) * W] = (f64)(((Q2[kk] * em[kk]) * B[2 * kk] + (Q1[kk] / em[kk]) * B[2 * kk + 1]) + zpu[kk]);
                        }
                    }
                } else {
                    /* setup_4_stream_fluxes, fluxes.py:3387-3607.  per-layer scratch in Am:
                     * 8 p/q values, 2 exps, 8 z values */
                    real *pq = Am, *ex = Am + 8 * L;
                    for (int l = 0; l < L; ++l) {
                        real a0 = a_[l], a1 = a_[L + l], a2 = a_[2 * L + l], a3 = a_[3 * L + l];
                        real b0 = b_[l], b1 = b_[L + l], b2 = b_[2 * L + l], b3 = b_[3 * L + l];
                        real beta = a0 * a1 + 4 * a0 * a3 / 9 + a2 * a3 / 9;
                        real gama = a0 * a1 * a2 * a3 / 9;
                        real l1 = R_SQRT((beta + R_SQRT(beta * beta - 4 * gama)) / 2);
                        real l2 = R_SQRT((beta - R_SQRT(beta * beta - 4 * gama)) / 2);
                        lam1[l] = l1; lam2[l] = l2;
                        real iu = 1 / u0;
                        real x2 = iu * iu;
                        real Del = 9 * (x2 * x2 - beta * x2 + gama);
                        real D0 = ((a1 * b0 - b1 / u0) * (a2 * a3 - 9 / (u0 * u0)) +
                                   2 * (a3 * b2 - 2 * a3 * b0 - 3 * b3 / u0) / (u0 * u0));
                        real D1 = ((a0 * b1 - b0 / u0) * (a2 * a3 - 9 / (u0 * u0)) -
                                   2 * a0 * (a3 * b2 - 3 * b3 / u0) / u0);
                        real D2 = ((a3 * b2 - 3 * b3 / u0) * (a0 * a1 - 1 / (u0 * u0)) -
                                   2 * a3 * (a0 * b1 - b0 / u0) / u0);
                        real D3 = ((a2 * b3 - 3 * b2 / u0) * (a0 * a1 - 1 / (u0 * u0)) +
                                   2 * (3 * a0 * b1 - 2 * a0 * b3 - 3 * b0 / u0) / (u0 * u0));
                        real e0 = D0 / Del, e1 = D1 / Del, e2 = D2 / Del, e3 = D3 / Del;
                        eta[l] = e0; eta[L + l] = e1; eta[2 * L + l] = e2; eta[3 * L + l] = e3;
                        real z1pl = (e0 / 2 + e1 + 5 * e2 / 8) * 2 * PI;
                        real z1mn = (e0 / 2 - e1 + 5 * e2 / 8) * 2 * PI;
                        real z2pl = (-e0 / 8 + 5 * e2 / 8 + e3) * 2 * PI;
                        real z2mn = (-e0 / 8 + 5 * e2 / 8 - e3) * 2 * PI;
                        ex[l] = R_EXP(-clip35(l1 * LW(dtau, l)));
                        ex[L + l] = R_EXP(-clip35(l2 * LW(dtau, l)));
                        real R1 = -a0 / l1, R2 = -a0 / l2;
                        real Q1 = 0.5 * (a0 * a1 / (l1 * l1) - 1), Q2 = 0.5 * (a0 * a1 / (l2 * l2) - 1);
                        real S1 = -3 / (2 * a3) * (a0 * a1 / l1 - l1), S2 = -3 / (2 * a3) * (a0 * a1 / l2 - l2);
                        pq[0 * L + l] = (0.5 + R1 + 5 * Q1 / 8) * 2 * PI;   /* p1pl */
                        pq[1 * L + l] = (0.5 + R2 + 5 * Q2 / 8) * 2 * PI;   /* p2pl */
                        pq[2 * L + l] = (-0.125 + 5 * Q1 / 8 + S1) * 2 * PI; /* q1pl */
                        pq[3 * L + l] = (-0.125 + 5 * Q2 / 8 + S2) * 2 * PI; /* q2pl */
                        pq[4 * L + l] = (0.5 - R1 + 5 * Q1 / 8) * 2 * PI;   /* p1mn */
                        pq[5 * L + l] = (0.5 - R2 + 5 * Q2 / 8) * 2 * PI;   /* p2mn */
                        pq[6 * L + l] = (-0.125 + 5 * Q1 / 8 - S1) * 2 * PI; /* q1mn */
                        pq[7 * L + l] = (-0.125 + 5 * Q2 / 8 - S2) * 2 * PI; /* q2mn */
                        real et = R_EXP(-clip35(LW(tau, l) / u0)), eb = R_EXP(-clip35(LW(tau, l + 1) / u0));
                        zz[0 * L + l] = z1mn * eb; zz[1 * L + l] = z2mn * eb; /* up */
                        zz[2 * L + l] = z1pl * eb; zz[3 * L + l] = z2pl * eb;
                        zz[4 * L + l] = z1mn * et; zz[5 * L + l] = z2mn * et; /* down */
                        zz[6 * L + l] = z1pl * et; zz[7 * L + l] = z2pl * et;
                        /* keep R, Q, S (A tensor rows, fluxes.py:3601-3605) for the source function;
                         * b[] and w_single[] of this layer are not needed any more */
                        b_[l] = R1; b_[L + l] = R2; b_[2 * L + l] = Q1; b_[3 * L + l] = Q2;
                        ws[l] = S1; ws[L + l] = S2;
                    }
#define P1PL(l) pq[0 * L + (l)]
#define P2PL(l) pq[1 * L + (l)]
#define Q1PL(l) pq[2 * L + (l)]
#define Q2PL(l) pq[3 * L + (l)]
#define P1MN(l) pq[4 * L + (l)]
#define P2MN(l) pq[5 * L + (l)]
#define Q1MN(l) pq[6 * L + (l)]
#define Q2MN(l) pq[7 * L + (l)]
#define E1(l) ex[(l)]
#define E2(l) ex[L + (l)]
#define F00(l) (P1MN(l) * E1(l))
#define F01(l) (P1PL(l) / E1(l))
#define F02(l) (P2MN(l) * E2(l))
#define F03(l) (P2PL(l) / E2(l))
#define F10(l) (Q1MN(l) * E1(l))
#define F11(l) (Q1PL(l) / E1(l))
#define F12(l) (Q2MN(l) * E2(l))
#define F13(l) (Q2PL(l) / E2(l))
#define F20(l) (P1PL(l) * E1(l))
#define F21(l) (P1MN(l) / E1(l))
#define F22(l) (P2PL(l) * E2(l))
#define F23(l) (P2MN(l) / E2(l))
#define F30(l) (Q1PL(l) * E1(l))
#define F31(l) (Q1MN(l) / E1(l))
#define F32(l) (Q2PL(l) * E2(l))
#define F33(l) (Q2MN(l) / E2(l))
                    MB(5, 0) = P1MN(0); MB(5, 1) = Q1PL(0); MB(4, 1) = P1PL(0); MB(4, 2) = Q2MN(0);
                    MB(3, 2) = P2MN(0); MB(3, 3) = Q2PL(0); MB(2, 3) = P2PL(0); MB(6, 0) = Q1MN(0);
                    B[0] = bt - zz[4 * L + 0];
                    B[1] = -bt / 4 - zz[5 * L + 0];
                    int nn = L - 1;
                    MB(5, 4 * L - 2) = F22(nn) - r * F02(nn);
                    MB(5, 4 * L - 1) = F33(nn) - r * F13(nn);
                    MB(4, 4 * L - 1) = F23(nn) - r * F03(nn);
                    MB(6, 4 * L - 3) = F21(nn) - r * F01(nn);
                    MB(6, 4 * L - 2) = F32(nn) - r * F12(nn);
                    MB(7, 4 * L - 4) = F20(nn) - r * F00(nn);
                    MB(7, 4 * L - 3) = F31(nn) - r * F11(nn);
                    MB(8, 4 * L - 4) = F30(nn) - r * F10(nn);
                    B[4 * L - 2] = b_surface - zz[2 * L + nn] + r * zz[0 * L + nn];
                    B[4 * L - 1] = b_surface_SH4 - zz[3 * L + nn] + r * zz[1 * L + nn];
                    for (int kk = 0; kk < L - 1; ++kk) {
                        int c = 4 * kk;
                        MB(5, c + 2) = F02(kk); MB(5, c + 3) = F13(kk);
                        MB(5, c + 4) = -P1PL(kk + 1); MB(5, c + 5) = -Q1MN(kk + 1);
                        MB(4, c + 3) = F03(kk); MB(4, c + 4) = -Q1MN(kk + 1);
                        MB(4, c + 5) = -P1MN(kk + 1); MB(4, c + 6) = -Q2PL(kk + 1);
                        MB(3, c + 4) = -P1MN(kk + 1); MB(3, c + 5) = -Q1PL(kk + 1);
                        MB(3, c + 6) = -P2PL(kk + 1); MB(3, c + 7) = -Q2MN(kk + 1);
                        MB(2, c + 5) = -P1PL(kk + 1); MB(2, c + 6) = -Q2MN(kk + 1); MB(2, c + 7) = -P2MN(kk + 1);
                        MB(1, c + 6) = -P2MN(kk + 1); MB(1, c + 7) = -Q2PL(kk + 1);
                        MB(0, c + 7) = -P2PL(kk + 1);
                        MB(6, c + 1) = F01(kk); MB(6, c + 2) = F12(kk); MB(6, c + 3) = F23(kk);
                        MB(6, c + 4) = -Q1PL(kk + 1);
                        MB(7, c) = F00(kk); MB(7, c + 1) = F11(kk); MB(7, c + 2) = F22(kk); MB(7, c + 3) = F33(kk);
                        MB(8, c) = F10(kk); MB(8, c + 1) = F21(kk); MB(8, c + 2) = F32(kk);
                        MB(9, c) = F20(kk); MB(9, c + 1) = F31(kk);
                        MB(10, c) = F30(kk);
                        B[c + 2] = zz[4 * L + kk + 1] - zz[0 * L + kk];
                        B[c + 3] = zz[5 * L + kk + 1] - zz[1 * L + kk];
                        B[c + 4] = zz[6 * L + kk + 1] - zz[2 * L + kk];
                        B[c + 5] = zz[7 * L + kk + 1] - zz[3 * L + kk];
                    }
                    flux_bot_row[0] = F20(nn); flux_bot_row[1] = F21(nn);
                    flux_bot_row[2] = F22(nn); flux_bot_row[3] = F23(nn);
                    G_bot = zz[2 * L + nn];
                    band_solve(n, k, ab, ldab, ipiv, B);
                    if (flux) {
                        /* calculate_flux(F, G, X), F and G of fluxes.py:3551-3598 */
                        f64 *fo = flux + (size_t)ai * 4 * nlevel * W + w;
                        fo[0] = (f64)((((P1MN(0) * B[0] + P1PL(0) * B[1]) + P2MN(0) * B[2]) + P2PL(0) * B[3]) + zz[4 * L + 0]);
                        fo[(size_t)W] = (f64)((((Q1MN(0) * B[0] + Q1PL(0) * B[1]) + Q2MN(0) * B[2]) + Q2PL(0) * B[3]) + zz[5 * L + 0]);
                        fo[(size_t)2 * W] = (f64)((((P1PL(0) * B[0] + P1MN(0) * B[1]) + P2PL(0) * B[2]) + P2MN(0) * B[3]) + zz[6 * L + 0]);
                        fo[(size_t)3 * W] = (f64)((((Q1PL(0) * B[0] + Q1MN(0) * B[1]) + Q2PL(0) * B[2]) + Q2MN(0) * B[3]) + zz[7 * L + 0]);
                        for (int kk = 0; kk < L; ++kk) {
                            const real *x = B + 4 * kk;
                            fo[(size_t)(4 * kk + 4) * W] = (f64)((((F00(kk) * x[0] + F01(kk) * x[1]) + F02(kk) * x[2]) + F03(kk) * x[3]) + zz[0 * L + kk]);
                            fo[(size_t)(4 * kk + 5) * W] = (f64)((((F10(kk) * x[0] + F11(kk) * x[1]) + F12(kk) * x[2]) + F13(kk) * x[3]) + zz[1 * L + kk]);
                            fo[(size_t)(4 * kk + 6) * W] = (f64)((((F20(kk) * x[0] + F21(kk) * x[1]) + F22(kk) * x[2]) + F23(kk) * x[3]) + zz[2 * L + kk]);
                            fo[(size_t)(4 * kk + 7) * W] = (f64)((((F30(kk) * x[0] + F31(kk) * x[1]) + F32(kk) * x[2]) + F33(kk) * x[3]) + zz[3 * L + kk]);
                        }
                    }
                }
                /* B now holds X.  flux at the bottom, fluxes.py:2891 */
                real flux_bot = G_bot;
                for (int c = 0; c < S; ++c) flux_bot += flux_bot_row[c] * B[S * (L - 1) + c];
                /* source-function technique, fluxes.py:2898-2972 */
                const real mus = (u1 + u0) / (u1 * u0);
                real xi = flux_bot / PI;
                for (int l = L - 1; l >= 0; --l) {
                    real dt = LW(dtau, l);
                    real exptrm_mus = (1 - R_EXP(-clip35(mus * dt))) / mus;
                    real exptau_mu = R_EXP(-clip35(LW(tau, l) * 1 / u0));
                    real expon1 = exptrm_mus * exptau_mu;
                    real multi;
                    if (S == 2) {
                        real alpha = 1 / u1 + lam1[l], beta = 1 / u1 - lam1[l];
                        real ea = (1 - R_EXP(-clip35(alpha * dt))) / alpha;
                        real eb = (1 - R_EXP(-clip35(beta * dt))) / beta;
                        real wm0 = wm[l], wm1 = wm[L + l];
                        real A0 = B[2 * l] * (wm0 - wm1 * Pu1[1] * qv[l]) * ea;
                        real A1 = B[2 * l + 1] * (wm0 + wm1 * Pu1[1] * qv[l]) * eb;
                        real N0 = wm0 * (eta[l] * expon1);
                        real N1 = wm1 * Pu1[1] * (eta[L + l] * expon1);
                        multi = A0 + N0 + A1 + N1;
                    } else {
                        real R1 = b_[l], R2 = b_[L + l], Q1 = b_[2 * L + l], Q2 = b_[3 * L + l];
                        real S1 = ws[l], S2 = ws[L + l];
                        real al1 = 1 / u1 + lam1[l], al2 = 1 / u1 + lam2[l];
                        real be1 = 1 / u1 - lam1[l], be2 = 1 / u1 - lam2[l];
                        real et[4];
                        et[0] = (1 - R_EXP(-clip35(al1 * dt))) / al1 * B[4 * l];
                        et[1] = (1 - R_EXP(-clip35(be1 * dt))) / be1 * B[4 * l + 1];
                        et[2] = (1 - R_EXP(-clip35(al2 * dt))) / al2 * B[4 * l + 2];
                        et[3] = (1 - R_EXP(-clip35(be2 * dt))) / be2 * B[4 * l + 3];
                        real Arow[4][4] = {{1, 1, 1, 1}, {R1, -R1, R2, -R2}, {Q1, Q1, Q2, Q2}, {S1, -S1, S2, -S2}};
                        real Aint[4] = {0, 0, 0, 0};
                        for (int j = 0; j < 4; ++j)
                            for (int c = 0; c < 4; ++c)
                                Aint[c] = Aint[c] + wm[j * L + l] * Pu1[j] * Arow[j][c];
                        for (int c = 0; c < 4; ++c) Aint[c] = Aint[c] * et[c];
                        real N0 = wm[l] * Pu1[0] * eta[l] * expon1;
                        real N1 = wm[L + l] * Pu1[1] * eta[L + l] * expon1;
                        real N2 = wm[2 * L + l] * Pu1[2] * eta[2 * L + l] * expon1;
                        real N3 = wm[3 * L + l] * Pu1[3] * eta[3 * L + l] * expon1;
                        multi = (Aint[0] + N0 + Aint[1] + N1 + Aint[2] + N2 + Aint[3] + N3);
                    }
                    real e1m = R_EXP(-clip35(mus * LW(dtau_og, l)));
                    real integ = (LW(w0, l) * multi +
                                  LW(w0_og, l) * f0 / (4 * PI) * ps[l] * (1 - e1m) *
                                      R_EXP(-LW(tau_og, l) / u0) / mus);
                    xi = xi * R_EXP(-dt / u1) + integ / u1;
                }
                xint_at_top[(size_t)ai * W + w] = (f64)xi;
            }
            /* the reference leaves its caller's f_deltaM modified (Appendix A1) */
            for (int l = 0; l < L; ++l) f_deltaM[(size_t)l * W + w] = (f64)fd[l];
#undef LW
        }
        free(ab); free(B); free(ipiv); free(lay);
    }
}


/* blackbody(t, 1/wno), fluxes.py:1676-1680 */
static inline real sh_planck(real t, real wno)
{
    const real h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
    real w = 1.0 / wno;
    return ((2.0 * h * c * c) / R_POW(w, (real)5.0)) * (1.0 / (R_EXP((h * c) / (t * (w * k))) - 1.0));
}

/* get_thermal_SH, fluxes.py:2979-3186 with the calculation == 1 branches of
 * setup_2_stream_fluxes (:3266-3270) and setup_4_stream_fluxes (:3451-3459); flx = 0.
 * The banded system does not depend on the viewing angle and is solved once per wavelength. */
void ORC_NAME(orc_get_thermal_SH)(
    int nlevel, const f64 *wno, int nwno, int numg, int numt, const f64 *tlevel,
    const f64 *dtau, const f64 *w0, const f64 *cosb, const f64 *cosb_og, const f64 *plevel,
    const f64 *ubar1, const f64 *surf_reflect, int stream, int hard_surface,
    f64 *xint_at_top, int nthreads)
{
    const int L = nlevel - 1, W = nwno, G = numg * numt, S = stream;
    const int n = S * L, k = 3 * S / 2 - 1, ldab = 3 * k + 1;
    const real mu1 = 0.5;
    /* ff = 0 when cosb equals cosb_og everywhere (np.array_equal), else cosb_og**stream (:3044-3047) */
    int same = 1;
    for (size_t i = 0; i < (size_t)L * W && same; ++i) same = (cosb[i] == cosb_og[i]);
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        real *ab = (real *)malloc(sizeof(real) * (size_t)ldab * n);
        real *B = (real *)malloc(sizeof(real) * (size_t)n);
        int *ipiv = (int *)malloc(sizeof(int) * (size_t)n);
        real *lay = (real *)malloc(sizeof(real) * (size_t)(L + 1) * 40);
        real *bb = lay, *b0 = bb + (L + 1), *b1 = b0 + L, *a_ = b1 + L, *wm = a_ + 4 * L, *lam1 = wm + 4 * L,
             *lam2 = lam1 + L, *qv = lam2 + L, *pq = qv + L /* 8L */, *ex = pq + 8 * L /* 2L */,
             *zz = ex + 2 * L /* 8L */, *RQS = zz + 8 * L /* 6L */;
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
#define LW(arr, l) ((real)(arr)[(size_t)(l) * W + w])
            const real r = surf_reflect[w];
            for (int i = 0; i <= L; ++i) bb[i] = sh_planck(tlevel[i], wno[w]);
            for (int l = 0; l < L; ++l) {
                b0[l] = bb[l];
                b1[l] = (bb[l + 1] - b0[l]) / LW(dtau, l);
                real g = LW(cosb_og, l);
                real ff = same ? 0. * g : R_POW(g, (real)S);
                for (int m = 0; m < S; ++m) {
                    wm[m * L + l] = (2 * m + 1) * (R_POW(g, (real)m) - ff) / (1 - ff);
                    a_[m * L + l] = (2 * m + 1) - LW(w0, l) * wm[m * L + l];
                }
            }
            real tau_top = LW(dtau, 0) * plevel[0] / (plevel[1] - plevel[0]);
            real b_top = PI * (1.0 - R_EXP(-tau_top / mu1)) * bb[0];
            real b_surface = hard_surface ? PI * bb[L] : PI * (bb[L] + b1[L - 1] * mu1);
            real b_surface_SH4 = (-PI * bb[L] / 4);
            memset(ab, 0, sizeof(real) * (size_t)ldab * n);
            memset(B, 0, sizeof(real) * (size_t)n);
            if (S == 2) {
                real *Q1 = pq, *Q2 = pq + L, *em = ex, *zmu = zz, *zpu = zz + L, *zmd = zz + 2 * L, *zpd = zz + 3 * L;
                for (int l = 0; l < L; ++l) {
                    real a0 = a_[l], a1 = a_[L + l], om = LW(w0, l), dt = LW(dtau, l);
                    lam1[l] = R_SQRT(a0 * a1);
                    em[l] = R_EXP(-clip35(lam1[l] * dt));
                    qv[l] = lam1[l] / a1;
                    Q1[l] = (0.5 + qv[l]) * 2 * PI;
                    Q2[l] = (0.5 - qv[l]) * 2 * PI;
                    zmd[l] = ((1 - om) / a0 * (b0[l] / 2 - b1[l] / a1)) * 2 * PI;
                    zmu[l] = ((1 - om) / a0 * (b0[l] / 2 - b1[l] / a1 + b1[l] * dt / 2)) * 2 * PI;
                    zpd[l] = ((1 - om) / a0 * (b0[l] / 2 + b1[l] / a1)) * 2 * PI;
                    zpu[l] = ((1 - om) / a0 * (b0[l] / 2 + b1[l] / a1 + b1[l] * dt / 2)) * 2 * PI;
                }
                MB(2, 0) = Q1[0];
                MB(1, 1) = Q2[0];
                B[0] = b_top - zmd[0];
                int nn = L - 1;
                MB(3, 2 * L - 2) = Q2[nn] * em[nn] - r * (Q1[nn] * em[nn]);
                MB(2, 2 * L - 1) = Q1[nn] / em[nn] - r * (Q2[nn] / em[nn]);
                B[2 * L - 1] = b_surface - zpu[nn] + r * zmu[nn];
                for (int kk = 0; kk < L - 1; ++kk) {
                    MB(0, 2 * kk + 3) = -Q2[kk + 1];
                    MB(1, 2 * kk + 2) = -Q1[kk + 1];
                    MB(1, 2 * kk + 3) = -Q1[kk + 1];
                    MB(2, 2 * kk + 1) = Q2[kk] / em[kk];
                    MB(2, 2 * kk + 2) = -Q2[kk + 1];
                    MB(3, 2 * kk) = Q1[kk] * em[kk];
                    MB(3, 2 * kk + 1) = Q1[kk] / em[kk];
                    MB(4, 2 * kk) = Q2[kk] * em[kk];
                    B[2 * kk + 1] = zmd[kk + 1] - zmu[kk];
                    B[2 * kk + 2] = zpd[kk + 1] - zpu[kk];
                }
            } else {
                for (int l = 0; l < L; ++l) {
                    real a0 = a_[l], a1 = a_[L + l], a2 = a_[2 * L + l], a3 = a_[3 * L + l];
                    real om = LW(w0, l), dt = LW(dtau, l);
                    real beta = a0 * a1 + 4 * a0 * a3 / 9 + a2 * a3 / 9;
                    real gama = a0 * a1 * a2 * a3 / 9;
                    real l1 = R_SQRT((beta + R_SQRT(beta * beta - 4 * gama)) / 2);
                    real l2 = R_SQRT((beta - R_SQRT(beta * beta - 4 * gama)) / 2);
                    lam1[l] = l1; lam2[l] = l2;
                    ex[l] = R_EXP(-clip35(l1 * dt));
                    ex[L + l] = R_EXP(-clip35(l2 * dt));
                    real R1 = -a0 / l1, R2 = -a0 / l2;
                    real Q1 = 0.5 * (a0 * a1 / (l1 * l1) - 1), Q2 = 0.5 * (a0 * a1 / (l2 * l2) - 1);
                    real S1 = -3 / (2 * a3) * (a0 * a1 / l1 - l1), S2 = -3 / (2 * a3) * (a0 * a1 / l2 - l2);
                    RQS[l] = R1; RQS[L + l] = R2; RQS[2 * L + l] = Q1; RQS[3 * L + l] = Q2;
                    RQS[4 * L + l] = S1; RQS[5 * L + l] = S2;
                    pq[0 * L + l] = (0.5 + R1 + 5 * Q1 / 8) * 2 * PI;
                    pq[1 * L + l] = (0.5 + R2 + 5 * Q2 / 8) * 2 * PI;
                    pq[2 * L + l] = (-0.125 + 5 * Q1 / 8 + S1) * 2 * PI;
                    pq[3 * L + l] = (-0.125 + 5 * Q2 / 8 + S2) * 2 * PI;
                    pq[4 * L + l] = (0.5 - R1 + 5 * Q1 / 8) * 2 * PI;
                    pq[5 * L + l] = (0.5 - R2 + 5 * Q2 / 8) * 2 * PI;
                    pq[6 * L + l] = (-0.125 + 5 * Q1 / 8 - S1) * 2 * PI;
                    pq[7 * L + l] = (-0.125 + 5 * Q2 / 8 - S2) * 2 * PI;
                    /* fluxes.py:3452-3459; order (z1mn, z2mn, z1pl, z2pl), up then down */
                    zz[0 * L + l] = (1 - om) / a0 * (b0[l] / 2 - b1[l] / a1 + b1[l] * dt / 2) * 2 * PI;
                    zz[1 * L + l] = -0.5 * (1 - om) / (4 * a0) * (b0[l] + b1[l] * dt) * 2 * PI;
                    zz[2 * L + l] = (1 - om) / a0 * (b0[l] / 2 + b1[l] / a1 + b1[l] * dt / 2) * 2 * PI;
                    zz[3 * L + l] = -0.5 * (1 - om) / (4 * a0) * (b0[l] + b1[l] * dt) * 2 * PI;
                    zz[4 * L + l] = (1 - om) / a0 * (b0[l] / 2 - b1[l] / a1) * 2 * PI;
                    zz[5 * L + l] = -0.5 * (1 - om) / (4 * a0) * (b0[l]) * 2 * PI;
                    zz[6 * L + l] = (1 - om) / a0 * (b0[l] / 2 + b1[l] / a1) * 2 * PI;
                    zz[7 * L + l] = -0.5 * (1 - om) / (4 * a0) * (b0[l]) * 2 * PI;
                }
                MB(5, 0) = P1MN(0); MB(5, 1) = Q1PL(0); MB(4, 1) = P1PL(0); MB(4, 2) = Q2MN(0);
                MB(3, 2) = P2MN(0); MB(3, 3) = Q2PL(0); MB(2, 3) = P2PL(0); MB(6, 0) = Q1MN(0);
                B[0] = b_top - zz[4 * L + 0];
                B[1] = -b_top / 4 - zz[5 * L + 0];
                int nn = L - 1;
                MB(5, 4 * L - 2) = F22(nn) - r * F02(nn);
                MB(5, 4 * L - 1) = F33(nn) - r * F13(nn);
                MB(4, 4 * L - 1) = F23(nn) - r * F03(nn);
                MB(6, 4 * L - 3) = F21(nn) - r * F01(nn);
                MB(6, 4 * L - 2) = F32(nn) - r * F12(nn);
                MB(7, 4 * L - 4) = F20(nn) - r * F00(nn);
                MB(7, 4 * L - 3) = F31(nn) - r * F11(nn);
                MB(8, 4 * L - 4) = F30(nn) - r * F10(nn);
                B[4 * L - 2] = b_surface - zz[2 * L + nn] + r * zz[0 * L + nn];
                B[4 * L - 1] = b_surface_SH4 - zz[3 * L + nn] + r * zz[1 * L + nn];
                for (int kk = 0; kk < L - 1; ++kk) {
                    int c = 4 * kk;
                    MB(5, c + 2) = F02(kk); MB(5, c + 3) = F13(kk);
                    MB(5, c + 4) = -P1PL(kk + 1); MB(5, c + 5) = -Q1MN(kk + 1);
                    MB(4, c + 3) = F03(kk); MB(4, c + 4) = -Q1MN(kk + 1);
                    MB(4, c + 5) = -P1MN(kk + 1); MB(4, c + 6) = -Q2PL(kk + 1);
                    MB(3, c + 4) = -P1MN(kk + 1); MB(3, c + 5) = -Q1PL(kk + 1);
                    MB(3, c + 6) = -P2PL(kk + 1); MB(3, c + 7) = -Q2MN(kk + 1);
                    MB(2, c + 5) = -P1PL(kk + 1); MB(2, c + 6) = -Q2MN(kk + 1); MB(2, c + 7) = -P2MN(kk + 1);
                    MB(1, c + 6) = -P2MN(kk + 1); MB(1, c + 7) = -Q2PL(kk + 1);
                    MB(0, c + 7) = -P2PL(kk + 1);
                    MB(6, c + 1) = F01(kk); MB(6, c + 2) = F12(kk); MB(6, c + 3) = F23(kk);
                    MB(6, c + 4) = -Q1PL(kk + 1);
                    MB(7, c) = F00(kk); MB(7, c + 1) = F11(kk); MB(7, c + 2) = F22(kk); MB(7, c + 3) = F33(kk);
                    MB(8, c) = F10(kk); MB(8, c + 1) = F21(kk); MB(8, c + 2) = F32(kk);
                    MB(9, c) = F20(kk); MB(9, c + 1) = F31(kk);
                    MB(10, c) = F30(kk);
                    B[c + 2] = zz[4 * L + kk + 1] - zz[0 * L + kk];
                    B[c + 3] = zz[5 * L + kk + 1] - zz[1 * L + kk];
                    B[c + 4] = zz[6 * L + kk + 1] - zz[2 * L + kk];
                    B[c + 5] = zz[7 * L + kk + 1] - zz[3 * L + kk];
                }
            }
            band_solve(n, k, ab, ldab, ipiv, B);
            /* per-angle source-function integration, fluxes.py:3105-3182 */
            for (int ai = 0; ai < G; ++ai) {
                const real u1 = ubar1[ai];
                real Pu1[4];
                legp(u1, Pu1);
                real xi = hard_surface ? bb[L] * 2 * PI : (bb[L] + b1[L - 1] * u1) * 2 * PI;
                for (int l = L - 1; l >= 0; --l) {
                    real dt = LW(dtau, l), om = LW(w0, l);
                    real a0 = a_[l], a1 = a_[L + l];
                    real multi;
                    if (S == 2) {
                        real alpha = 1 / u1 + lam1[l], beta = 1 / u1 - lam1[l];
                        real ea = (1 - R_EXP(-clip35(alpha * dt))) / alpha;
                        real eb = (1 - R_EXP(-clip35(beta * dt))) / beta;
                        real wm0 = wm[l], wm1 = wm[L + l];
                        real A0 = B[2 * l] * (wm0 - wm1 * Pu1[1] * qv[l]) * ea;
                        real A1 = B[2 * l + 1] * (wm0 + wm1 * Pu1[1] * qv[l]) * eb;
                        real ed = R_EXP(-dt / u1);
                        real N0 = wm0 * ((1 - om) * u1 / a0 * (b0[l] * (1 - ed) + b1[l] * (u1 - (dt + u1) * ed)));
                        real N1 = wm1 * Pu1[1] * ((1 - om) * u1 / a0 * (b1[l] * (1 - ed) / a1));
                        multi = A0 + N0 + A1 + N1;
                    } else {
                        real R1 = RQS[l], R2 = RQS[L + l], Q1 = RQS[2 * L + l], Q2 = RQS[3 * L + l];
                        real S1 = RQS[4 * L + l], S2 = RQS[5 * L + l];
                        real al1 = 1 / u1 + lam1[l], al2 = 1 / u1 + lam2[l];
                        real be1 = 1 / u1 - lam1[l], be2 = 1 / u1 - lam2[l];
                        real et[4];
                        et[0] = (1 - R_EXP(-clip35(al1 * dt))) / al1 * B[4 * l];
                        et[1] = (1 - R_EXP(-clip35(be1 * dt))) / be1 * B[4 * l + 1];
                        et[2] = (1 - R_EXP(-clip35(al2 * dt))) / al2 * B[4 * l + 2];
                        et[3] = (1 - R_EXP(-clip35(be2 * dt))) / be2 * B[4 * l + 3];
                        real Arow[4][4] = {{1, 1, 1, 1}, {R1, -R1, R2, -R2}, {Q1, Q1, Q2, Q2}, {S1, -S1, S2, -S2}};
                        real Aint[4] = {0, 0, 0, 0};
                        for (int j = 0; j < 4; ++j)
                            for (int c = 0; c < 4; ++c) Aint[c] = Aint[c] + wm[j * L + l] * Pu1[j] * Arow[j][c];
                        for (int c = 0; c < 4; ++c) Aint[c] = Aint[c] * et[c];
                        real ed = R_EXP(-clip35(dt / u1));
                        real N0 = wm[l] * ((1 - om) * u1 / a0 * (b0[l] * (1 - ed) + b1[l] * (u1 - (dt + u1) * ed)));
                        real N1 = wm[L + l] * u1 * ((1 - om) * u1 / a0 * (b1[l] * (1 - ed) / a1));
                        multi = Aint[0] + Aint[1] + Aint[2] + Aint[3] + N0 + N1 + 0 + 0;
                    }
                    real ed = R_EXP(-(dt / u1));
                    real integ = (om * multi * 2 * PI +
                                  2 * PI * (1 - om) * u1 * (b0[l] * (1 - ed) + b1[l] * (u1 - (dt + u1) * ed)));
                    xi = xi * R_EXP(-dt / u1) + integ / u1;
                }
                xint_at_top[(size_t)ai * W + w] = (f64)xi;
            }
#undef LW
        }
        free(ab); free(B); free(ipiv); free(lay);
    }
}
