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

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

/* tri_diag_solve, fluxes.py:311-323: eliminate from the last row upwards, then
 * substitute from the first row downwards. */
static void tri_solve(int n, const double *a, const double *b, const double *c, const double *d,
                      double *as, double *ds, double *x)
{
    as[n - 1] = a[n - 1] / b[n - 1];
    ds[n - 1] = d[n - 1] / b[n - 1];
    for (int i = n - 2; i >= 0; --i) {
        double inv = 1.0 / (b[i] - c[i] * as[i + 1]);
        as[i] = a[i] * inv;
        ds[i] = (d[i] - c[i] * ds[i + 1]) * inv;
    }
    x[0] = ds[0];
    for (int i = 1; i < n; ++i) x[i] = ds[i] - as[i] * x[i - 1];
}

/* setup_tri_diag, fluxes.py:139-183, for one wavelength column. */
static void build_tridiag(int L, const double *cpu, const double *cmu, const double *cpd,
                          const double *cmd, double b_top, double b_surface, double r,
                          const double *gam, const double *ep, const double *em,
                          double *A, double *B, double *C, double *D)
{
    int n = 2 * L;
    A[0] = 0.0;
    B[0] = gam[0] + 1.0;
    C[0] = gam[0] - 1.0;
    D[0] = b_top - cmu[0];
    for (int l = 0; l < L - 1; ++l) {
        double e1 = ep[l] + gam[l] * em[l];
        double e2 = ep[l] - gam[l] * em[l];
        double e3 = gam[l] * ep[l] + em[l];
        double e4 = gam[l] * ep[l] - em[l];
        double gn = gam[l + 1];
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
        double e1 = ep[l] + gam[l] * em[l];
        double e2 = ep[l] - gam[l] * em[l];
        double e3 = gam[l] * ep[l] + em[l];
        double e4 = gam[l] * ep[l] - em[l];
        A[n - 1] = e1 - r * e3;
        B[n - 1] = e2 - r * e4;
        C[n - 1] = 0.0;
        D[n - 1] = b_surface - cpd[l] + r * cmd[l];
    }
}

static inline double hg_down(double g, double cos_theta)
{
    /* fluxes.py:1310: Henyey-Greenstein in the frame of the downward beam (+ sign) */
    double t = 1.0 + g * g + 2.0 * g * cos_theta;
    return (1.0 - g * g) / sqrt(t * t * t);
}

void orc_get_reflected_1d(
    int nlevel, int nwno, int numg, int numt,
    const double *dtau, const double *tau, const double *w0, const double *cosb,
    const double *gcos2, const double *ftau_cld, const double *ftau_ray,
    const double *dtau_og, const double *tau_og, const double *w0_og, const double *cosb_og,
    const double *surf_reflect, const double *ubar0, const double *ubar1,
    double cos_theta, const double *F0PI,
    int single_phase, int multi_phase,
    double frac_a, double frac_b, double frac_c, double constant_back, double constant_forward,
    int get_toa_intensity, int get_lvl_flux, int toon_coefficients,
    const double *b_top,
    double *xint_at_top, double *flux_minus, double *flux_plus,
    double *flux_minus_mdpt, double *flux_plus_mdpt, int nthreads)
{
    const int L = nlevel - 1, W = nwno, G = numg * numt;
    const double sq3 = sqrt(3.0);
    (void)nthreads;
    memset(xint_at_top, 0, sizeof(double) * (size_t)G * W);
    if (flux_minus) {
        size_t nb = sizeof(double) * (size_t)G * nlevel * W;
        memset(flux_minus, 0, nb); memset(flux_plus, 0, nb);
        memset(flux_minus_mdpt, 0, nb); memset(flux_plus_mdpt, 0, nb);
    }
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        /* per-thread work arrays: 12 layer vectors + 7 row vectors of 2L */
        double *buf = (double *)malloc(sizeof(double) * (size_t)(14 * L + 14 * L + 2 * nlevel));
        double *g1 = buf, *g2 = g1 + L, *lam = g2 + L, *gam = lam + L, *cpu = gam + L,
               *cmu = cpu + L, *cpd = cmu + L, *cmd = cpd + L, *ep = cmd + L, *em = ep + L,
               *pos = em + L, *neg = pos + L, *ex = neg + L, *apl = ex + L;
        double *A = apl + L, *B = A + 2 * L, *C = B + 2 * L, *D = C + 2 * L, *AS = D + 2 * L,
               *DS = AS + 2 * L, *X = DS + 2 * L;
        double *xint = X + 2 * L; /* nlevel */
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
#define LW(a, l) ((a)[(size_t)(l) * W + w])
            /* angle independent, fluxes.py:1132-1141 */
            for (int l = 0; l < L; ++l) {
                double om = LW(w0, l), fc = LW(ftau_cld, l), g = LW(cosb, l);
                if (toon_coefficients == 1) {
                    g1[l] = (7.0 - om * (4.0 + 3.0 * fc * g)) / 4.0;
                    g2[l] = -(1.0 - om * (4.0 - 3.0 * fc * g)) / 4.0;
                } else {
                    g1[l] = (sq3 * 0.5) * (2.0 - om * (1.0 + fc * g));
                    g2[l] = (sq3 * om * 0.5) * (1.0 - fc * g);
                }
                lam[l] = sqrt(g1[l] * g1[l] - g2[l] * g2[l]);
                gam[l] = (g1[l] - lam[l]) / g2[l];
            }
            const double f0 = F0PI[w], r = surf_reflect[w], bt = b_top ? b_top[w] : 0.0;
            for (int a = 0; a < G; ++a) {
                const double u0 = ubar0[a], u1 = ubar1[a];
                /* fluxes.py:1146-1183 */
                for (int l = 0; l < L; ++l) {
                    double om = LW(w0, l), fc = LW(ftau_cld, l), g = LW(cosb, l);
                    double g3 = (toon_coefficients == 1) ? (2.0 - 3.0 * fc * g * u0) / 4.0
                                                         : 0.5 * (1.0 - sq3 * fc * g * u0);
                    double g4 = 1.0 - g3;
                    double den = lam[l] * lam[l] - 1.0 / (u0 * u0);
                    double a_minus = f0 * om * (g4 * (g1[l] + 1.0 / u0) + g2[l] * g3) / den;
                    double a_plus = f0 * om * (g3 * (g1[l] - 1.0 / u0) + g2[l] * g4) / den;
                    double xu = exp(-LW(tau, l) / u0), xd = exp(-LW(tau, l + 1) / u0);
                    cmu[l] = a_minus * xu; cpu[l] = a_plus * xu;
                    cmd[l] = a_minus * xd; cpd[l] = a_plus * xd;
                    apl[l] = a_plus; ex[l] = a_minus; /* keep a+- for the midpoint terms */
                    double e = lam[l] * LW(dtau, l);
                    if (e > 35.0) e = 35.0;
                    ep[l] = exp(e);
                    em[l] = 1.0 / ep[l];
                }
                double b_surface = 0.0 + r * u0 * f0 * exp(-LW(tau, L) / u0);
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
                        double fm = pos[l] * gam[l] + neg[l] + cmu[l];
                        double fp = pos[l] + gam[l] * neg[l] + cpu[l];
                        fm = fm + u0 * f0 * exp(-LW(tau, l) / u0);
                        flux_minus[base + (size_t)l * W + w] = fm;
                        flux_plus[base + (size_t)l * W + w] = fp;
                        double e = lam[l] * LW(dtau, l);
                        if (e > 35.0) e = 35.0;
                        double epm = exp(0.5 * e), emm = 1.0 / epm;
                        double taumid = LW(tau, l) + 0.5 * LW(dtau, l);
                        double xm = exp(-taumid / u0);
                        double cpm = apl[l] * xm, cmm = ex[l] * xm;
                        double fmm = gam[l] * pos[l] * epm + neg[l] * emm + cmm;
                        double fpm = pos[l] * epm + gam[l] * neg[l] * emm + cpm;
                        fmm = fmm + u0 * f0 * exp(-taumid / u0);
                        flux_minus_mdpt[base + (size_t)l * W + w] = fmm;
                        flux_plus_mdpt[base + (size_t)l * W + w] = fpm;
                    }
                    int l = L - 1;
                    double fzm = gam[l] * pos[l] * ep[l] + neg[l] * em[l] + cmd[l];
                    double fzp = pos[l] * ep[l] + gam[l] * neg[l] * em[l] + cpd[l];
                    fzm = fzm + u0 * f0 * exp(-LW(tau, L) / u0);
                    flux_minus[base + (size_t)L * W + w] = fzm;
                    flux_plus[base + (size_t)L * W + w] = fzp;
                }
                if (get_toa_intensity) {
                    /* fluxes.py:1262-1410 */
                    int lb = L - 1;
                    double flux_zero = pos[lb] * ep[lb] + gam[lb] * neg[lb] * em[lb] + cpd[lb];
                    xint[L] = flux_zero / PI;
                    for (int l = L - 1; l >= 0; --l) {
                        double om = LW(w0, l), fc = LW(ftau_cld, l), g = LW(cosb, l);
                        double mplus, mminus;
                        if (multi_phase == 0) {
                            const double ubar2 = 0.767;
                            double t2 = LW(gcos2, l) * (3.0 * ubar2 * ubar2 * u1 * u1 - 1.0) / 2.0;
                            mplus = 1.0 + 1.5 * fc * g * u1 + t2;
                            mminus = 1.0 - 1.5 * fc * g * u1 + t2;
                        } else {
                            mplus = 1.0 + 1.5 * fc * g * u1;
                            mminus = 1.0 - 1.5 * fc * g * u1;
                        }
                        double Gt = pos[l] * (mplus + gam[l] * mminus) * om * 0.5 / PI;
                        double Ht = neg[l] * (gam[l] * mplus + mminus) * om * 0.5 / PI;
                        double At = (mplus * cpu[l] + mminus * cmu[l]) * om * 0.5 / PI;
                        double go = LW(cosb_og, l), ps;
                        double gf = 0, gb = 0, f = 0;
                        if (single_phase != 1) {
                            gf = constant_forward * go;
                            gb = constant_back * go;
                            f = frac_a + frac_b * pow(gb, frac_c);
                        }
                        if (single_phase == 0)
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
                        double e = lam[l] * LW(dtau, l);
                        if (e > 35.0) e = 35.0;
                        double dt = LW(dtau, l);
                        xint[l] = xint[l + 1] * exp(-dt / u1) +
                                  (LW(w0_og, l) * f0 / (4.0 * PI)) * ps * exp(-LW(tau_og, l) / u0) *
                                      (1.0 - exp(-LW(dtau_og, l) * (u0 + u1) / (u0 * u1))) *
                                      (u0 / (u0 + u1)) +
                                  At * (1.0 - exp(-dt * (u0 + u1) / (u0 * u1))) * (u0 / (u0 + u1)) +
                                  Gt * (exp(e - dt / u1) - 1.0) / (lam[l] * u1 - 1.0) +
                                  Ht * (1.0 - exp(-e - dt / u1)) / (lam[l] * u1 + 1.0);
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
static inline double planck_wavelength(double t, double wno)
{
    const double h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
    double w = 1.0 / wno;
    return ((2.0 * h * c * c) / pow(w, 5.0)) * (1.0 / (exp((h * c) / (t * (w * k))) - 1.0));
}

/* blackbody_integrated, fluxes.py:1632-1656 (nbb = 1: three sub-bins) */
static inline double planck_binned(double t, double wave, double dwave)
{
    const double h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
    const double c1 = 2 * h * c * c, c2 = h * c / k;
    double s = 0.0;
    for (int kk = -1; kk <= 1; ++kk) {
        double wavenum = wave + kk * dwave / 2.0;
        s += c1 * (wavenum * wavenum * wavenum) / (exp(c2 * wavenum / t) - 1.0);
    }
    return s / 3.0;
}

void orc_get_thermal_1d(
    int nlevel, const double *wno, int nwno, int numg, int numt,
    const double *tlevel, const double *dtau, const double *w0, const double *cosb,
    const double *plevel, const double *ubar1, const double *surf_reflect,
    int hard_surface, const double *dwno, int calc_type,
    double *flux_at_top, double *flux_minus, double *flux_plus,
    double *flux_minus_mdpt, double *flux_plus_mdpt, int nthreads)
{
    const int L = nlevel - 1, W = nwno, G = numg * numt, V = nlevel;
    const double mu1 = 0.5;
    (void)nthreads;
    if (flux_minus) {
        size_t nb = sizeof(double) * (size_t)G * V * W;
        memset(flux_minus, 0, nb); memset(flux_plus, 0, nb);
        memset(flux_minus_mdpt, 0, nb); memset(flux_plus_mdpt, 0, nb);
    }
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)(24 * L + 14 * L + 6 * V));
        double *bb = buf;            /* V */
        double *b0 = bb + V, *b1 = b0 + L, *lam = b1 + L, *gam = lam + L, *q = gam + L,
               *cpu = q + L, *cmu = cpu + L, *cpd = cmu + L, *cmd = cpd + L, *ep = cmd + L,
               *em = ep + L, *pos = em + L, *neg = pos + L, *epm = neg + L, *emm = epm + L,
               *Gt = emm + L, *Ht = Gt + L, *Jt = Ht + L, *Kt = Jt + L, *al1 = Kt + L,
               *al2 = al1 + L, *si1 = al2 + L, *si2 = si1 + L, *ext = si2 + L;
        double *A = ext + L, *B = A + 2 * L, *C = B + 2 * L, *D = C + 2 * L, *AS = D + 2 * L,
               *DS = AS + 2 * L, *X = DS + 2 * L;
        double *fm = X + 2 * L, *fp = fm + V, *fmm = fp + V, *fpm = fmm + V;
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
#define LW(a, l) ((a)[(size_t)(l) * W + w])
            for (int i = 0; i < V; ++i)
                bb[i] = (calc_type == 0) ? planck_wavelength(tlevel[i], wno[w])
                                         : planck_binned(tlevel[i], wno[w], dwno[w]);
            /* fluxes.py:1756-1789 */
            for (int l = 0; l < L; ++l) {
                double dt = LW(dtau, l), om = LW(w0, l), g = LW(cosb, l);
                b0[l] = bb[l];
                b1[l] = (bb[l + 1] - b0[l]) / dt;
                double g1 = 2.0 - om * (1 + g), g2 = om * (1 - g);
                lam[l] = sqrt(g1 * g1 - g2 * g2);
                gam[l] = (g1 - lam[l]) / g2;
                q[l] = 1.0 / (g1 + g2);
                cpu[l] = 2 * PI * mu1 * (b0[l] + b1[l] * q[l]);
                cmu[l] = 2 * PI * mu1 * (b0[l] - b1[l] * q[l]);
                cpd[l] = 2 * PI * mu1 * (b0[l] + b1[l] * dt + b1[l] * q[l]);
                cmd[l] = 2 * PI * mu1 * (b0[l] + b1[l] * dt - b1[l] * q[l]);
                double e = lam[l] * dt;
                if (e > 35.0) e = 35.0;
                ep[l] = exp(e); em[l] = 1.0 / ep[l];
                epm[l] = exp(0.5 * e); emm[l] = 1 / epm[l];
            }
            /* fluxes.py:1797-1806 */
            double tau_top = LW(dtau, 0) * plevel[0] / (plevel[1] - plevel[0]);
            double b_top = (1.0 - exp(-tau_top / mu1)) * bb[0] * PI;
            double r = surf_reflect[w];
            double b_surface = hard_surface ? (1.0 - r) * bb[L] * PI
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
                double u = ubar1[a];
                for (int i = 0; i < V; ++i) fm[i] = fp[i] = fmm[i] = fpm[i] = 0.0;
                fp[L] = hard_surface ? (1.0 - r) * bb[L] * 2 * PI
                                     : (bb[L] + b1[L - 1] * u) * 2 * PI;
                fm[0] = (1 - exp(-tau_top / u)) * bb[0] * 2 * PI;
                for (int it = 0; it < L; ++it) {
                    double dt = LW(dtau, it);
                    double xa = exp(-dt / u), xh = exp(-0.5 * dt / u);
                    fm[it + 1] = fm[it] * xa + (Jt[it] / (lam[it] * u + 1.0)) * (ep[it] - xa) +
                                 (Kt[it] / (lam[it] * u - 1.0)) * (xa - em[it]) +
                                 si1[it] * (1. - xa) + si2[it] * (u * xa + dt - u);
                    fmm[it] = fm[it] * xh + (Jt[it] / (lam[it] * u + 1.0)) * (epm[it] - xh) +
                              (Kt[it] / (-lam[it] * u + 1.0)) * (emm[it] - xh) +
                              si1[it] * (1. - xh) + si2[it] * (u * xh + 0.5 * dt - u);
                    int ib = L - 1 - it;
                    dt = LW(dtau, ib);
                    xa = exp(-dt / u); xh = exp(-0.5 * dt / u);
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

void orc_get_transit_1d(
    const double *z, const double *dz, int nlevel, int nwno, double rstar,
    const double *mmw, double k_b, double amu, const double *player, const double *tlayer,
    const double *colden, const double *DTAU, double *F, int nthreads)
{
    const int V = nlevel, L = nlevel - 1, W = nwno;
    (void)nthreads;
    /* path lengths, fluxes.py:2624-2644 */
    double *dl = (double *)calloc((size_t)V * V, sizeof(double));
    for (int i = 0; i < V; ++i)
        for (int j = 0; j < i; ++j) {
            double ref = z[i], inner = z[i - j], outer = z[i - j - 1], seg = 0.0;
            if (inner != ref && outer != ref)
                seg = sqrt(outer * outer - ref * ref) - sqrt(inner * inner - ref * ref);
            else if (inner == ref)
                seg = sqrt(outer * outer - ref * ref);
            dl[(size_t)i * V + j] = seg * player[i - j - 1] / tlayer[i - j - 1] / k_b;
        }
    double zmin = z[0];
    for (int i = 1; i < V; ++i) if (z[i] < zmin) zmin = z[i];
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        double *t = (double *)malloc(sizeof(double) * L);
#pragma omp for schedule(static)
        for (int w = 0; w < W; ++w) {
            /* fluxes.py:2648-2661 */
            for (int k = 0; k < L; ++k) t[k] = DTAU[(size_t)k * W + w] / colden[k] * (mmw[k] * amu);
            double acc = 0.0;
            for (int i = 0; i < V; ++i) {
                double tauall = 0.0;
                for (int j = 0; j < i; ++j) tauall = tauall + 2 * t[i - j - 1] * dl[(size_t)i * V + j];
                acc += (1. - exp(-tauall)) * (z[i] * dz[i]);
            }
            F[w] = (zmin / rstar) * (zmin / rstar) + 2. / (rstar * rstar) * acc;
        }
        free(t);
    }
    free(dl);
}

void orc_compress_disco(int nwno, double cos_theta, const double *xint_at_top,
                        const double *gweight, int ng, const double *tweight, int nt,
                        const double *F0PI, double *albedo)
{
    /* disco.py:138-149 */
    double sym = (nt == 1) ? 2 * PI : 1.0;
    for (int w = 0; w < nwno; ++w) {
        double s = 0.0;
        for (int ig = 0; ig < ng; ++ig)
            for (int it = 0; it < nt; ++it)
                s = s + xint_at_top[((size_t)ig * nt + it) * nwno + w] * gweight[ig] * tweight[it];
        albedo[w] = sym * 0.5 * s / F0PI[w] * (cos_theta + 1.0);
    }
}

void orc_compress_thermal(int n, const double *flux_at_top, const double *gweight, int ng,
                          const double *tweight, int nt, double *flux)
{
    /* disco.py:169-180 */
    double sym = (nt == 1) ? 1.0 : 1 / (2 * PI);
    for (int w = 0; w < n; ++w) {
        double s = 0.0;
        for (int ig = 0; ig < ng; ++ig)
            for (int it = 0; it < nt; ++it)
                s = s + flux_at_top[((size_t)ig * nt + it) * n + w] * gweight[ig] * tweight[it];
        flux[w] = s * sym;
    }
}
