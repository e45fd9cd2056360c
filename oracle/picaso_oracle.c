/*
 * picaso_oracle.c - CPU restatement (plain C, fp64) of the reference's Toon89
 * reflected / thermal solvers, transit chord integration and disk integration.
 * TEST INFRASTRUCTURE ONLY - see picaso_oracle.h.  Each function cites the
 * reference lines it follows (/root/reference/picaso, commit 0369089).
 *
 * The restatement keeps the reference's algorithm (full 2L-row tridiagonal built
 * as in setup_tri_diag, bottom-up Thomas sweep as in tri_diag_solve, bottom-up
 * source-function recurrence) and works one wavelength column at a time.
 */
#include "picaso_oracle.h"

/* Precision-generic: compiled twice.  Default build = fp64 arithmetic (the oracle).
 * -DORC_QUAD = the same algorithm in IEEE binary128 (libquadmath) on the same fp64
 * inputs and constants, exported with a _quad suffix: the "exact arithmetic" yardstick
 * used to judge entries where the reference algorithm itself is ill-conditioned. */
typedef double f64;
#ifdef ORC_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_EXP expq
#define R_SQRT sqrtq
#define R_POW powq
#define ORC_NAME(x) x##_quad
#else
typedef double real;
#define R_EXP exp
#define R_SQRT sqrt
#define R_POW pow
#define ORC_NAME(x) x
#endif

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

/* 0: the 1-D reference functions; 1: the per-facet formulas of get_reflected_3d / get_thermal_3d
 * (fluxes.py:355-660, :2148-2352): |ubar|, exponent clip 40, the 3-D 'cahoy' phase function, and the
 * pi-based thermal boundary terms.  Set only by the *_3d entry points below. */
static int g_variant = 0;

/* tri_diag_solve, fluxes.py:311-323: eliminate from the last row upwards, then
 * substitute from the first row downwards. */
static void tri_solve(int n, const real *a, const real *b, const real *c, const real *d,
                      real *as, real *ds, real *x)
{
    as[n - 1] = a[n - 1] / b[n - 1];
    ds[n - 1] = d[n - 1] / b[n - 1];
    for (int i = n - 2; i >= 0; --i) {
        real inv = 1.0 / (b[i] - c[i] * as[i + 1]);
        as[i] = a[i] * inv;
        ds[i] = (d[i] - c[i] * ds[i + 1]) * inv;
    }
    x[0] = ds[0];
    for (int i = 1; i < n; ++i) x[i] = ds[i] - as[i] * x[i - 1];
}

/* setup_tri_diag, fluxes.py:139-183, for one wavelength column. */
static void build_tridiag(int L, const real *cpu, const real *cmu, const real *cpd,
                          const real *cmd, real b_top, real b_surface, real r,
                          const real *gam, const real *ep, const real *em,
                          real *A, real *B, real *C, real *D)
{
    int n = 2 * L;
    A[0] = 0.0;
    B[0] = gam[0] + 1.0;
    C[0] = gam[0] - 1.0;
    D[0] = b_top - cmu[0];
    for (int l = 0; l < L - 1; ++l) {
        real e1 = ep[l] + gam[l] * em[l];
        real e2 = ep[l] - gam[l] * em[l];
        real e3 = gam[l] * ep[l] + em[l];
        real e4 = gam[l] * ep[l] - em[l];
        real gn = gam[l + 1];
        int o = 2 * l + 1, e = 2 * l + 2;
        A[o] = (e1 + e3) * (gn - 1.0);
        B[o] = (e2 + e4) * (gn - 1.0);
        C[o] = 2.0 * (1.0 - gn * gn);
        D[o] = (gn - 1.0) * (cpu[l + 1] - cpd[l]) + (1.0 - gn) * (cmd[l] - cmu[l + 1]);
        A[e] = 2.0 * (1.0 - gam[l] * gam[l]);
        B[e] = (e1 - e3) * (gn + 1.0);
        C[e] = (e1 + e3) * (gn - 1.0);
        D[e] = e3 * (cpu[l + 1] - cpd[l]) + e1 * (cmd[l] - cmu[l + 1]);
    }
    {
        int l = L - 1;
        real e1 = ep[l] + gam[l] * em[l];
        real e2 = ep[l] - gam[l] * em[l];
        real e3 = gam[l] * ep[l] + em[l];
        real e4 = gam[l] * ep[l] - em[l];
        A[n - 1] = e1 - r * e3;
        B[n - 1] = e2 - r * e4;
        C[n - 1] = 0.0;
        D[n - 1] = b_surface - cpd[l] + r * cmd[l];
    }
}

static inline real hg_down(real g, real cos_theta)
{
    /* fluxes.py:1310: Henyey-Greenstein in the frame of the downward beam (+ sign) */
    real t = 1.0 + g * g + 2.0 * g * cos_theta;
    return (1.0 - g * g) / R_SQRT(t * t * t);
}

void ORC_NAME(orc_get_reflected_1d)(
    int nlevel, int nwno, int numg, int numt,
    const f64 *dtau, const f64 *tau, const f64 *w0, const f64 *cosb,
    const f64 *gcos2, const f64 *ftau_cld, const f64 *ftau_ray,
    const f64 *dtau_og, const f64 *tau_og, const f64 *w0_og, const f64 *cosb_og,
    const f64 *surf_reflect, const f64 *ubar0, const f64 *ubar1,
    f64 cos_theta, const f64 *F0PI,
    int single_phase, int multi_phase,
    f64 frac_a, f64 frac_b, f64 frac_c, f64 constant_back, f64 constant_forward,
    int get_toa_intensity, int get_lvl_flux, int toon_coefficients,
    const f64 *b_top,
    f64 *xint_at_top, f64 *flux_minus, f64 *flux_plus,
    f64 *flux_minus_mdpt, f64 *flux_plus_mdpt, int nthreads)
{
    const int L = nlevel - 1, W = nwno, G = numg * numt;
    const real sq3 = R_SQRT(3.0);
    (void)nthreads;
    memset(xint_at_top, 0, sizeof(f64) * (size_t)G * W);
    if (flux_minus) {
        size_t nb = sizeof(f64) * (size_t)G * nlevel * W;
        memset(flux_minus, 0, nb); memset(flux_plus, 0, nb);
        memset(flux_minus_mdpt, 0, nb); memset(flux_plus_mdpt, 0, nb);
    }
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        /* per-thread work arrays: 12 layer vectors + 7 row vectors of 2L */
        real *buf = (real *)malloc(sizeof(real) * (size_t)(14 * L + 14 * L + 2 * nlevel));
        real *g1 = buf, *g2 = g1 + L, *lam = g2 + L, *gam = lam + L, *cpu = gam + L,
               *cmu = cpu + L, *cpd = cmu + L, *cmd = cpd + L, *ep = cmd + L, *em = ep + L,
               *pos = em + L, *neg = pos + L, *ex = neg + L, *apl = ex + L;
        real *A = apl + L, *B = A + 2 * L, *C = B + 2 * L, *D = C + 2 * L, *AS = D + 2 * L,
               *DS = AS + 2 * L, *X = DS + 2 * L;
        real *xint = X + 2 * L; /* nlevel */
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
#define LW(a, l) ((a)[(size_t)(l) * W + w])
            /* angle independent, fluxes.py:1132-1141 */
            for (int l = 0; l < L; ++l) {
                real om = LW(w0, l), fc = LW(ftau_cld, l), g = LW(cosb, l);
                if (toon_coefficients == 1) {
                    g1[l] = (7.0 - om * (4.0 + 3.0 * fc * g)) / 4.0;
                    g2[l] = -(1.0 - om * (4.0 - 3.0 * fc * g)) / 4.0;
                } else {
                    g1[l] = (sq3 * 0.5) * (2.0 - om * (1.0 + fc * g));
                    g2[l] = (sq3 * om * 0.5) * (1.0 - fc * g);
                }
                lam[l] = R_SQRT(g1[l] * g1[l] - g2[l] * g2[l]);
                gam[l] = (g1[l] - lam[l]) / g2[l];
            }
            const real f0 = F0PI[w], r = surf_reflect[w], bt = b_top ? b_top[w] : 0.0;
            for (int a = 0; a < G; ++a) {
                const real u0 = g_variant ? (real)fabs(ubar0[a]) : (real)ubar0[a];
                const real u1 = g_variant ? (real)fabs(ubar1[a]) : (real)ubar1[a];
                /* fluxes.py:1146-1183 */
                for (int l = 0; l < L; ++l) {
                    real om = LW(w0, l), fc = LW(ftau_cld, l), g = LW(cosb, l);
                    real g3 = (toon_coefficients == 1) ? (2.0 - 3.0 * fc * g * u0) / 4.0
                                                         : 0.5 * (1.0 - sq3 * fc * g * u0);
                    real g4 = 1.0 - g3;
                    real den = lam[l] * lam[l] - 1.0 / (u0 * u0);
                    real a_minus = f0 * om * (g4 * (g1[l] + 1.0 / u0) + g2[l] * g3) / den;
                    real a_plus = f0 * om * (g3 * (g1[l] - 1.0 / u0) + g2[l] * g4) / den;
                    real xu = R_EXP(-LW(tau, l) / u0), xd = R_EXP(-LW(tau, l + 1) / u0);
                    cmu[l] = a_minus * xu; cpu[l] = a_plus * xu;
                    cmd[l] = a_minus * xd; cpd[l] = a_plus * xd;
                    apl[l] = a_plus; ex[l] = a_minus; /* keep a+- for the midpoint terms */
                    real e = lam[l] * LW(dtau, l);
                    if (e > (g_variant ? 40.0 : 35.0)) e = (g_variant ? 40.0 : 35.0);
                    ep[l] = R_EXP(e);
                    em[l] = 1.0 / ep[l];
                }
                real b_surface = 0.0 + r * u0 * f0 * R_EXP(-LW(tau, L) / u0);
                build_tridiag(L, cpu, cmu, cpd, cmd, bt, b_surface, r, gam, ep, em, A, B, C, D);
                tri_solve(2 * L, A, B, C, D, AS, DS, X);
                for (int l = 0; l < L; ++l) {
                    pos[l] = X[2 * l] + X[2 * l + 1];
                    neg[l] = X[2 * l] - X[2 * l + 1];
                }
                if (get_lvl_flux) {
                    /* fluxes.py:1219-1257 */
                    size_t base = (size_t)a * nlevel * W;
                    for (int l = 0; l < L; ++l) {
                        real fm = pos[l] * gam[l] + neg[l] + cmu[l];
                        real fp = pos[l] + gam[l] * neg[l] + cpu[l];
                        fm = fm + u0 * f0 * R_EXP(-LW(tau, l) / u0);
                        flux_minus[base + (size_t)l * W + w] = fm;
                        flux_plus[base + (size_t)l * W + w] = fp;
                        real e = lam[l] * LW(dtau, l);
                        if (e > (g_variant ? 40.0 : 35.0)) e = (g_variant ? 40.0 : 35.0);
                        real epm = R_EXP(0.5 * e), emm = 1.0 / epm;
                        real taumid = LW(tau, l) + 0.5 * LW(dtau, l);
                        real xm = R_EXP(-taumid / u0);
                        real cpm = apl[l] * xm, cmm = ex[l] * xm;
                        real fmm = gam[l] * pos[l] * epm + neg[l] * emm + cmm;
                        real fpm = pos[l] * epm + gam[l] * neg[l] * emm + cpm;
                        fmm = fmm + u0 * f0 * R_EXP(-taumid / u0);
                        flux_minus_mdpt[base + (size_t)l * W + w] = fmm;
                        flux_plus_mdpt[base + (size_t)l * W + w] = fpm;
                    }
                    int l = L - 1;
                    real fzm = gam[l] * pos[l] * ep[l] + neg[l] * em[l] + cmd[l];
                    real fzp = pos[l] * ep[l] + gam[l] * neg[l] * em[l] + cpd[l];
                    fzm = fzm + u0 * f0 * R_EXP(-LW(tau, L) / u0);
                    flux_minus[base + (size_t)L * W + w] = fzm;
                    flux_plus[base + (size_t)L * W + w] = fzp;
                }
                if (get_toa_intensity) {
                    /* fluxes.py:1262-1410 */
                    int lb = L - 1;
                    real flux_zero = pos[lb] * ep[lb] + gam[lb] * neg[lb] * em[lb] + cpd[lb];
                    xint[L] = flux_zero / PI;
                    for (int l = L - 1; l >= 0; --l) {
                        real om = LW(w0, l), fc = LW(ftau_cld, l), g = LW(cosb, l);
                        real mplus, mminus;
                        if (multi_phase == 0) {
                            const real ubar2 = 0.767;
                            real t2 = LW(gcos2, l) * (3.0 * ubar2 * ubar2 * u1 * u1 - 1.0) / 2.0;
                            mplus = 1.0 + 1.5 * fc * g * u1 + t2;
                            mminus = 1.0 - 1.5 * fc * g * u1 + t2;
                        } else {
                            mplus = 1.0 + 1.5 * fc * g * u1;
                            mminus = 1.0 - 1.5 * fc * g * u1;
                        }
                        real Gt = pos[l] * (mplus + gam[l] * mminus) * om * 0.5 / PI;
                        real Ht = neg[l] * (gam[l] * mplus + mminus) * om * 0.5 / PI;
                        real At = (mplus * cpu[l] + mminus * cmu[l]) * om * 0.5 / PI;
                        real go = LW(cosb_og, l), ps;
                        real gf = 0, gb = 0, f = 0;
                        if (single_phase != 1) {
                            gf = constant_forward * go;
                            gb = constant_back * go;
                            f = frac_a + frac_b * R_POW(gb, frac_c);
                        }
                        if (single_phase == 0 && g_variant) {
                            /* get_reflected_3d's 'cahoy' (fluxes.py:582-588): denominators use cosb_og
                             * and -cosb_og/2 instead of g_forward / g_back */
                            real tf = 1 + go * go + 2 * go * cos_theta;
                            real hb = -go / 2.;
                            real tb = 1 + hb * hb + 2 * hb * cos_theta;
                            ps = f * (1 - gf * gf) / R_SQRT(tf * tf * tf) +
                                 (1 - f) * (1 - gb * gb) / R_SQRT(tb * tb * tb) + (LW(gcos2, l));
                        } else if (single_phase == 0)
                            ps = f * hg_down(gf, cos_theta) + (1.0 - f) * hg_down(gb, cos_theta) +
                                 LW(gcos2, l);
                        else if (single_phase == 1)
                            ps = hg_down(go, cos_theta);
                        else if (single_phase == 2)
                            ps = f * hg_down(gf, cos_theta) + (1.0 - f) * hg_down(gb, cos_theta);
                        else
                            ps = fc * (f * hg_down(gf, cos_theta) +
                                       (1.0 - f) * hg_down(gb, cos_theta)) +
                                 LW(ftau_ray, l) * (0.75 * (1.0 + cos_theta * cos_theta));
                        real e = lam[l] * LW(dtau, l);
                        if (e > (g_variant ? 40.0 : 35.0)) e = (g_variant ? 40.0 : 35.0);
                        real dt = LW(dtau, l);
                        xint[l] = xint[l + 1] * R_EXP(-dt / u1) +
                                  (LW(w0_og, l) * f0 / (4.0 * PI)) * ps * R_EXP(-LW(tau_og, l) / u0) *
                                      (1.0 - R_EXP(-LW(dtau_og, l) * (u0 + u1) / (u0 * u1))) *
                                      (u0 / (u0 + u1)) +
                                  At * (1.0 - R_EXP(-dt * (u0 + u1) / (u0 * u1))) * (u0 / (u0 + u1)) +
                                  Gt * (R_EXP(e - dt / u1) - 1.0) / (lam[l] * u1 - 1.0) +
                                  Ht * (1.0 - R_EXP(-e - dt / u1)) / (lam[l] * u1 + 1.0);
                    }
                    xint_at_top[(size_t)a * W + w] = xint[0];
                }
            }
#undef LW
        }
        free(buf);
    }
}

/* blackbody, fluxes.py:1676-1680 with w = 1/wno (cm) */
static inline real planck_wavelength(real t, real wno)
{
    const real h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
    real w = 1.0 / wno;
    return ((2.0 * h * c * c) / R_POW(w, 5.0)) * (1.0 / (R_EXP((h * c) / (t * (w * k))) - 1.0));
}

/* blackbody_integrated, fluxes.py:1632-1656 (nbb = 1: three sub-bins) */
static inline real planck_binned(real t, real wave, real dwave)
{
    const real h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
    const real c1 = 2 * h * c * c, c2 = h * c / k;
    real s = 0.0;
    for (int kk = -1; kk <= 1; ++kk) {
        real wavenum = wave + kk * dwave / 2.0;
        s += c1 * (wavenum * wavenum * wavenum) / (R_EXP(c2 * wavenum / t) - 1.0);
    }
    return s / 3.0;
}

void ORC_NAME(orc_get_thermal_1d)(
    int nlevel, const f64 *wno, int nwno, int numg, int numt,
    const f64 *tlevel, const f64 *dtau, const f64 *w0, const f64 *cosb,
    const f64 *plevel, const f64 *ubar1, const f64 *surf_reflect,
    int hard_surface, const f64 *dwno, int calc_type,
    f64 *flux_at_top, f64 *flux_minus, f64 *flux_plus,
    f64 *flux_minus_mdpt, f64 *flux_plus_mdpt, int nthreads)
{
    const int L = nlevel - 1, W = nwno, G = numg * numt, V = nlevel;
    const real mu1 = 0.5;
    (void)nthreads;
    if (flux_minus) {
        size_t nb = sizeof(f64) * (size_t)G * V * W;
        memset(flux_minus, 0, nb); memset(flux_plus, 0, nb);
        memset(flux_minus_mdpt, 0, nb); memset(flux_plus_mdpt, 0, nb);
    }
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        real *buf = (real *)malloc(sizeof(real) * (size_t)(24 * L + 14 * L + 6 * V));
        real *bb = buf;            /* V */
        real *b0 = bb + V, *b1 = b0 + L, *lam = b1 + L, *gam = lam + L, *q = gam + L,
               *cpu = q + L, *cmu = cpu + L, *cpd = cmu + L, *cmd = cpd + L, *ep = cmd + L,
               *em = ep + L, *pos = em + L, *neg = pos + L, *epm = neg + L, *emm = epm + L,
               *Gt = emm + L, *Ht = Gt + L, *Jt = Ht + L, *Kt = Jt + L, *al1 = Kt + L,
               *al2 = al1 + L, *si1 = al2 + L, *si2 = si1 + L, *ext = si2 + L;
        real *A = ext + L, *B = A + 2 * L, *C = B + 2 * L, *D = C + 2 * L, *AS = D + 2 * L,
               *DS = AS + 2 * L, *X = DS + 2 * L;
        real *fm = X + 2 * L, *fp = fm + V, *fmm = fp + V, *fpm = fmm + V;
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
#define LW(a, l) ((a)[(size_t)(l) * W + w])
            for (int i = 0; i < V; ++i)
                bb[i] = (calc_type == 0) ? planck_wavelength(tlevel[i], wno[w])
                                         : planck_binned(tlevel[i], wno[w], dwno[w]);
            /* fluxes.py:1756-1789 */
            for (int l = 0; l < L; ++l) {
                real dt = LW(dtau, l), om = LW(w0, l), g = LW(cosb, l);
                b0[l] = bb[l];
                b1[l] = (bb[l + 1] - b0[l]) / dt;
                real g1 = 2.0 - om * (1 + g), g2 = om * (1 - g);
                lam[l] = R_SQRT(g1 * g1 - g2 * g2);
                gam[l] = (g1 - lam[l]) / g2;
                q[l] = 1.0 / (g1 + g2);
                cpu[l] = 2 * PI * mu1 * (b0[l] + b1[l] * q[l]);
                cmu[l] = 2 * PI * mu1 * (b0[l] - b1[l] * q[l]);
                cpd[l] = 2 * PI * mu1 * (b0[l] + b1[l] * dt + b1[l] * q[l]);
                cmd[l] = 2 * PI * mu1 * (b0[l] + b1[l] * dt - b1[l] * q[l]);
                real e = lam[l] * dt;
                if (e > 35.0) e = 35.0;
                ep[l] = R_EXP(e); em[l] = 1.0 / ep[l];
                epm[l] = R_EXP(0.5 * e); emm[l] = 1 / epm[l];
            }
            /* fluxes.py:1797-1806 */
            real tau_top = LW(dtau, 0) * plevel[0] / (plevel[1] - plevel[0]);
            real b_top = (1.0 - R_EXP(-tau_top / mu1)) * bb[0] * PI;
            real r = surf_reflect[w];
            real b_surface = hard_surface ? (g_variant ? PI * bb[L] : (1.0 - r) * bb[L] * PI)
                                            : (bb[L] + b1[L - 1] * mu1) * PI;
            build_tridiag(L, cpu, cmu, cpd, cmd, b_top, b_surface, r, gam, ep, em, A, B, C, D);
            tri_solve(2 * L, A, B, C, D, AS, DS, X);
            /* fluxes.py:1830-1849 */
            for (int l = 0; l < L; ++l) {
                pos[l] = X[2 * l] + X[2 * l + 1];
                neg[l] = X[2 * l] - X[2 * l + 1];
                Gt[l] = (1 / mu1 - lam[l]) * pos[l];
                Ht[l] = gam[l] * (lam[l] + 1 / mu1) * neg[l];
                Jt[l] = gam[l] * (lam[l] + 1 / mu1) * pos[l];
                Kt[l] = (1 / mu1 - lam[l]) * neg[l];
                al1[l] = 2 * PI * (b0[l] + b1[l] * (q[l] - mu1));
                al2[l] = 2 * PI * b1[l];
                si1[l] = 2 * PI * (b0[l] - b1[l] * (q[l] - mu1));
                si2[l] = 2 * PI * b1[l];
            }
            /* fluxes.py:1864-1910 */
            for (int a = 0; a < G; ++a) {
                real u = ubar1[a];
                for (int i = 0; i < V; ++i) fm[i] = fp[i] = fmm[i] = fpm[i] = 0.0;
                if (g_variant) { /* get_thermal_3d, fluxes.py:2302-2306 */
                    fp[L] = hard_surface ? PI * (b_surface) : PI * (bb[L] + b1[L - 1] * u);
                    fm[0] = PI * (1 - R_EXP(-tau_top / u)) * bb[0];
                } else {
                    fp[L] = hard_surface ? (1.0 - r) * bb[L] * 2 * PI
                                         : (bb[L] + b1[L - 1] * u) * 2 * PI;
                    fm[0] = (1 - R_EXP(-tau_top / u)) * bb[0] * 2 * PI;
                }
                for (int it = 0; it < L; ++it) {
                    real dt = LW(dtau, it);
                    real xa = R_EXP(-dt / u), xh = R_EXP(-0.5 * dt / u);
                    fm[it + 1] = fm[it] * xa + (Jt[it] / (lam[it] * u + 1.0)) * (ep[it] - xa) +
                                 (Kt[it] / (lam[it] * u - 1.0)) * (xa - em[it]) +
                                 si1[it] * (1. - xa) + si2[it] * (u * xa + dt - u);
                    fmm[it] = fm[it] * xh + (Jt[it] / (lam[it] * u + 1.0)) * (epm[it] - xh) +
                              (Kt[it] / (-lam[it] * u + 1.0)) * (emm[it] - xh) +
                              si1[it] * (1. - xh) + si2[it] * (u * xh + 0.5 * dt - u);
                    int ib = L - 1 - it;
                    dt = LW(dtau, ib);
                    xa = R_EXP(-dt / u); xh = R_EXP(-0.5 * dt / u);
                    fp[ib] = fp[ib + 1] * xa + (Gt[ib] / (lam[ib] * u - 1.0)) * (ep[ib] * xa - 1.0) +
                             (Ht[ib] / (lam[ib] * u + 1.0)) * (1.0 - em[ib] * xa) +
                             al1[ib] * (1. - xa) + al2[ib] * (u - (dt + u) * xa);
                    fpm[ib] = fp[ib + 1] * xh +
                              (Gt[ib] / (lam[ib] * u - 1.0)) * (ep[ib] * xh - epm[ib]) -
                              (Ht[ib] / (lam[ib] * u + 1.0)) * (em[ib] * xh - emm[ib]) +
                              al1[ib] * (1. - xh) + al2[ib] * (u + 0.5 * dt - (dt + u) * xh);
                }
                flux_at_top[(size_t)a * W + w] = fpm[0];
                if (flux_minus) {
                    size_t base = (size_t)a * V * W;
                    for (int i = 0; i < V; ++i) {
                        flux_minus[base + (size_t)i * W + w] = fm[i];
                        flux_plus[base + (size_t)i * W + w] = fp[i];
                        flux_minus_mdpt[base + (size_t)i * W + w] = fmm[i];
                        flux_plus_mdpt[base + (size_t)i * W + w] = fpm[i];
                    }
                }
            }
#undef LW
        }
        free(buf);
    }
}

void ORC_NAME(orc_get_transit_1d)(
    const f64 *z, const f64 *dz, int nlevel, int nwno, f64 rstar,
    const f64 *mmw, f64 k_b, f64 amu, const f64 *player, const f64 *tlayer,
    const f64 *colden, const f64 *DTAU, f64 *F, int nthreads)
{
    const int V = nlevel, L = nlevel - 1, W = nwno;
    (void)nthreads;
    /* path lengths, fluxes.py:2624-2644 */
    real *dl = (real *)calloc((size_t)V * V, sizeof(real));
    for (int i = 0; i < V; ++i)
        for (int j = 0; j < i; ++j) {
            real ref = z[i], inner = z[i - j], outer = z[i - j - 1], seg = 0.0;
            if (inner != ref && outer != ref)
                seg = R_SQRT(outer * outer - ref * ref) - R_SQRT(inner * inner - ref * ref);
            else if (inner == ref)
                seg = R_SQRT(outer * outer - ref * ref);
            dl[(size_t)i * V + j] = seg * player[i - j - 1] / tlayer[i - j - 1] / k_b;
        }
    real zmin = z[0];
    for (int i = 1; i < V; ++i) if (z[i] < zmin) zmin = z[i];
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        real *t = (real *)malloc(sizeof(real) * L);
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
            /* fluxes.py:2648-2661 */
            for (int k = 0; k < L; ++k) t[k] = DTAU[(size_t)k * W + w] / colden[k] * (mmw[k] * amu);
            real acc = 0.0;
            for (int i = 0; i < V; ++i) {
                real tauall = 0.0;
                for (int j = 0; j < i; ++j) tauall = tauall + 2 * t[i - j - 1] * dl[(size_t)i * V + j];
                acc += (1. - R_EXP(-tauall)) * (z[i] * dz[i]);
            }
            F[w] = (zmin / rstar) * (zmin / rstar) + 2. / (rstar * rstar) * acc;
        }
        free(t);
    }
    free(dl);
}

void ORC_NAME(orc_compress_disco)(int nwno, f64 cos_theta, const f64 *xint_at_top,
                        const f64 *gweight, int ng, const f64 *tweight, int nt,
                        const f64 *F0PI, f64 *albedo)
{
    /* disco.py:138-149 */
    real sym = (nt == 1) ? 2 * PI : 1.0;
    for (int w = 0; w < nwno; ++w) {
        real s = 0.0;
        for (int ig = 0; ig < ng; ++ig)
            for (int it = 0; it < nt; ++it)
                s = s + xint_at_top[((size_t)ig * nt + it) * nwno + w] * gweight[ig] * tweight[it];
        albedo[w] = sym * 0.5 * s / F0PI[w] * (cos_theta + 1.0);
    }
}

void ORC_NAME(orc_compress_thermal)(int n, const f64 *flux_at_top, const f64 *gweight, int ng,
                          const f64 *tweight, int nt, f64 *flux)
{
    /* disco.py:169-180 */
    real sym = (nt == 1) ? 1.0 : 1 / (2 * PI);
    for (int w = 0; w < n; ++w) {
        real s = 0.0;
        for (int ig = 0; ig < ng; ++ig)
            for (int it = 0; it < nt; ++it)
                s = s + flux_at_top[((size_t)ig * nt + it) * n + w] * gweight[ig] * tweight[it];
        flux[w] = s * sym;
    }
}


/* get_reflected_3d, fluxes.py:355-660.  Facet-major inputs: every layer/level array is
 * [numg*numt][nlayer|nlevel][nwno] (the Python wrapper transposes the reference's [.., nwno, ng, nt]). */
void ORC_NAME(orc_get_reflected_3d)(
    int nlevel, int nwno, int numg, int numt,
    const f64 *dtau, const f64 *tau, const f64 *w0, const f64 *cosb, const f64 *gcos2,
    const f64 *ftau_cld, const f64 *ftau_ray, const f64 *dtau_og, const f64 *tau_og,
    const f64 *w0_og, const f64 *cosb_og, const f64 *surf_reflect, const f64 *ubar0,
    const f64 *ubar1, f64 cos_theta, const f64 *F0PI, int single_phase, int multi_phase,
    f64 frac_a, f64 frac_b, f64 frac_c, f64 constant_back, f64 constant_forward,
    f64 *xint_at_top, int nthreads)
{
    const size_t nl = (size_t)(nlevel - 1) * nwno, nv = (size_t)nlevel * nwno;
    g_variant = 1;
    for (int a = 0; a < numg * numt; ++a)
        ORC_NAME(orc_get_reflected_1d)(nlevel, nwno, 1, 1, dtau + a * nl, tau + a * nv, w0 + a * nl, cosb + a * nl,
                             gcos2 + a * nl, ftau_cld + a * nl, ftau_ray + a * nl, dtau_og + a * nl,
                             tau_og + a * nv, w0_og + a * nl, cosb_og + a * nl, surf_reflect, ubar0 + a,
                             ubar1 + a, cos_theta, F0PI, single_phase, multi_phase, frac_a, frac_b, frac_c,
                             constant_back, constant_forward, 1, 0, 0, NULL, xint_at_top + (size_t)a * nwno,
                             NULL, NULL, NULL, NULL, nthreads);
    g_variant = 0;
}

/* get_thermal_3d, fluxes.py:2148-2352.  Facet-major inputs; tlevel/plevel are [numg*numt][nlevel]. */
void ORC_NAME(orc_get_thermal_3d)(
    int nlevel, const f64 *wno, int nwno, int numg, int numt, const f64 *tlevel, const f64 *dtau,
    const f64 *w0, const f64 *cosb, const f64 *plevel, const f64 *ubar1, const f64 *surf_reflect,
    int hard_surface, f64 *int_at_top, int nthreads)
{
    const size_t nl = (size_t)(nlevel - 1) * nwno;
    g_variant = 1;
    for (int a = 0; a < numg * numt; ++a)
        ORC_NAME(orc_get_thermal_1d)(nlevel, wno, nwno, 1, 1, tlevel + (size_t)a * nlevel, dtau + a * nl, w0 + a * nl,
                           cosb + a * nl, plevel + (size_t)a * nlevel, ubar1 + a, surf_reflect, hard_surface,
                           NULL, 0, int_at_top + (size_t)a * nwno, NULL, NULL, NULL, NULL, nthreads);
    g_variant = 0;
}
