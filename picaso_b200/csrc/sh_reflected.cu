// sh_reflected.cu - spherical-harmonics (P1 "SH2" / P3 "SH4") reflected-light solver, sm_100a.
//
// Replaces picaso/fluxes.py:2675-2976 (get_reflected_SH) with setup_2_stream_fluxes
// (:3189-3333), setup_4_stream_fluxes (:3336-3607), solve_4_stream_banded (:3610-3628,
// scipy.linalg.solve_banded -> LAPACK dgbsv) and legP (:3639).
//
// B200 design.  The reference assembles an (S*L) x (S*L) banded matrix (5 or 11 diagonals,
// entries spanning e^{+-35}) per wavelength and angle and hands it to LAPACK, then runs a
// bottom-up source-function recurrence over the solution.  Unpivoted elimination is
// unusable on this system, so pivoting is kept - but the band is
// never materialised:
//  * the matrix is block-bidiagonal: S unknowns per layer, S continuity rows per
//    interface  Fb_l X_l - T_{l+1} X_{l+1} = Zd_{l+1} - Zu_l,  S/2 boundary rows at the
//    top and S/2 at the surface;
//  * one bottom-up sweep per (wavelength, angle) thread keeps a register window of
//    S/2 carried constraint rows + S interface rows over 2S unknowns and eliminates the
//    lower layer's S unknowns with PARTIAL PIVOTING inside the window (the same
//    candidate set LAPACK's band pivoting has, kl + 1 = S + S/2 rows);
//  * the TOA intensity is a linear functional of the solution, so it rides along as one
//    more row that is never a pivot candidate (adjoint accumulation, as in
//    toon_reflected.cu): when the top boundary rows close the system, the functional's
//    constant term IS xint_at_top.  No LU factors, no back-substitution, no O(L) storage.
// The reference's cumulative in-place scaling of f_deltaM across angles (fluxes.py:2823,
// SURVEY.md Appendix A1) is reproduced: angle k uses f_deltaM * factor^(k+1).
#include <cstdlib>

#include "pb_common.cuh"
#include "pb_math.cuh"

namespace {

struct ShParams {
    int L, W, G, nt;
    int64_t ld, bs_layer, bs_level, bs_wave;
    const double *dtau, *w0, *fcld, *fray, *fdm, *dtau_og, *w0_og, *cosb_og, *tau, *tau_og;
    const double *surf, *f0pi, *btop;
    const double *ubar0, *ubar1, *gweight, *tweight;
    double cos_theta, frac_a, frac_b, frac_c, cback, cfwd;
    int wsf, wmf, psf, wsr, wmr, psr, single_form;
    double *xint, *albedo, *fdm_out;
    int fuse_albedo;
    // flx = 1 (layer fluxes, fluxes.py:2889-2890): wavelengths [w_begin, w_end) of this launch, pivot-row scratch
    // piv [L-1][kPivN][ncols] (column = ((b G + a) wcap + w - w_begin)), output flux [B][G][S V][W]
    int w_begin, w_end, wcap;
    double *piv, *flux;
};

constexpr int kWaves = 32;
// pivot rows kept per eliminated layer for the back-substitution of the flx = 1 path: row k of the S pivot rows keeps
// its entries k .. 2S (upper triangle over [X_{l+1} | X_l | rhs])
template <int S> struct PivRows { static constexpr int N = S * (2 * S + 1) - S * (S - 1) / 2; };

__device__ __forceinline__ double clip35(double x) { return fmin(fmax(x, -35.0), 35.0); }

template <int S>
struct Layer {       // everything the sweep needs from one layer
    double T[S][S];  // "top of layer" rows (no exponentials)
    double cs[S];    // column scaling of the "bottom of layer" rows  Fb = T * diag(cs)
    double Zu[S], Zd[S];
    double sw[S];    // source-function weights on X_l
    double sconst;   // X-independent source term
    double xa;       // exp(-dtau/u1)
};

// powers g^m for m = 1..3 by repeated multiplication
template <int S>
__device__ __forceinline__ void moments(const ShParams &p, double g, double fcld, double fray,
                                        double &fd /* in: f_deltaM seen by this angle's OTHG part;
                                                      out: after this angle's TTHG scaling */,
                                        double ws[4], double wm[4])
{
    // fluxes.py:2805-2840
#pragma unroll
    for (int m = 0; m < 4; ++m) ws[m] = wm[m] = 1.0;
    if (p.wsf == 1 || p.wmf == 1) {
        double gm = 1.0;
#pragma unroll
        for (int m = 1; m < S; ++m) {
            gm *= g;
            const double wv = (2 * m + 1) * gm;
            const double v = (wv - (2 * m + 1) * fd) / (1 - fd);
            if (p.wsf == 1) ws[m] = v;
            if (p.wmf == 1) wm[m] = v;
        }
    }
    if (p.wsf == 0 || p.wmf == 0) {
        const double gf = p.cfwd * g, gb = p.cback * g;
        const double gbc = (p.frac_c == 2.0) ? gb * gb : pow(gb, p.frac_c);
        const double f = p.frac_a + p.frac_b * gbc;
        double cfS = 1.0, cbS = 1.0;
#pragma unroll
        for (int m = 0; m < S; ++m) { cfS *= p.cfwd; cbS *= p.cback; }
        fd *= (f * cfS + (1 - f) * cbS);
        double gfm = 1.0, gbm = 1.0;
#pragma unroll
        for (int m = 1; m < S; ++m) {
            gfm *= gf;
            gbm *= gb;
            const double wv = (2 * m + 1) * (f * gfm + (1 - f) * gbm);
            const double v = (wv - (2 * m + 1) * fd) / (1 - fd);
            if (p.wsf == 0) ws[m] = v;
            if (p.wmf == 0) wm[m] = v;
        }
    }
    if (p.wsr == 1) {
#pragma unroll
        for (int m = 1; m < S; ++m) ws[m] *= fcld;
        if (S == 4) ws[2] += 0.5 * fray;
    }
    if (p.wmr == 1) {
#pragma unroll
        for (int m = 1; m < S; ++m) wm[m] *= fcld;
        if (S == 4) wm[2] += 0.5 * fray;
    }
}

__device__ __forceinline__ double sh_psingle(const ShParams &p, double g, double fcld, double fray)
{
    // fluxes.py:2843-2855
    double ps;
    if (p.psf == 1) {
        const double s = sqrt(1 + g * g + 2 * g * p.cos_theta);
        ps = (1 - g * g) / (s * s * s);
    } else {
        const double gf = p.cfwd * g, gb = p.cback * g;
        const double gbc = (p.frac_c == 2.0) ? gb * gb : pow(gb, p.frac_c);
        const double f = p.frac_a + p.frac_b * gbc;
        const double tf = 1 + gf * gf + 2 * gf * p.cos_theta, tb = 1 + gb * gb + 2 * gb * p.cos_theta;
        ps = f * (1 - gf * gf) / sqrt(tf * tf * tf) + (1 - f) * (1 - gb * gb) / sqrt(tb * tb * tb);
    }
    if (p.psr == 1) ps = fcld * ps + fray * (0.75 * (1 + p.cos_theta * p.cos_theta));
    return ps;
}

// Forward elimination of the first S unknowns of an NR-row window with partial pivoting
// (pivot = largest |entry| among the not-yet-used rows, bubbled into place by conditional
// swaps so that every index stays compile-time), carrying the functional row along.
template <int S, int NR, int NC>
__device__ __forceinline__ void eliminate(double (&R)[NR][NC], double (&F)[NC])
{
#pragma unroll
    for (int k = 0; k < S; ++k) {
#pragma unroll
        for (int r = k + 1; r < NR; ++r) {
            const bool sw = fabs(R[r][k]) > fabs(R[k][k]);
#pragma unroll
            for (int c = k; c < NC; ++c) {
                const double a = R[k][c], b = R[r][c];
                R[k][c] = sw ? b : a;
                R[r][c] = sw ? a : b;
            }
        }
        const double inv = pbm::krcp(R[k][k]);
#pragma unroll
        for (int r = k + 1; r < NR; ++r) {
            const double f = R[r][k] * inv;
#pragma unroll
            for (int c = k + 1; c < NC; ++c) R[r][c] = fma(-f, R[k][c], R[r][c]);
        }
        // functional J = F[0..NC-2] . x + F[NC-1];  row k reads  R[k][.] . x = R[k][NC-1]
        const double f = F[k] * inv;
#pragma unroll
        for (int c = k + 1; c < NC - 1; ++c) F[c] = fma(-f, R[k][c], F[c]);
        F[NC - 1] = fma(f, R[k][NC - 1], F[NC - 1]);
    }
}

template <int S>
__device__ __forceinline__ void sh_layer(const ShParams &p, int64_t il, int64_t iv, double u0, double u1,
                                         double f0, const double (&Pu0)[4], const double (&Pu1)[4],
                                         double mus, int angle_index, double &et /* exp(-tau_l/u0) out */,
                                         double eb /* exp(-tau_{l+1}/u0), unclipped */, Layer<S> &o,
                                         double *fdm_out)
{
    const double TWO_PI = 2 * PB_PI;
    const double om = __ldg(p.w0 + il), dt = __ldg(p.dtau + il);
    const double fcld = __ldg(p.fcld + il), fray = __ldg(p.fray + il);
    const double g = __ldg(p.cosb_og + il);
    double fd = __ldg(p.fdm + il);
    double ws[4], wm[4];
    // Appendix A1: angle k sees the caller's f_deltaM already scaled k times (only when a TTHG
    // form is active); replay those scalings with the same operation order
    if (p.wsf == 0 || p.wmf == 0) {
        const double gb = p.cback * g;
        const double gbc = (p.frac_c == 2.0) ? gb * gb : pow(gb, p.frac_c);
        const double f = p.frac_a + p.frac_b * gbc;
        double cfS = 1.0, cbS = 1.0;
#pragma unroll
        for (int m = 0; m < S; ++m) { cfS *= p.cfwd; cbS *= p.cback; }
        const double fac = (f * cfS + (1 - f) * cbS);
        for (int i = 0; i < angle_index; ++i) fd *= fac;
    }
    moments<S>(p, g, fcld, fray, fd, ws, wm);
    if (fdm_out) *fdm_out = fd;
    double ps;
    if (p.single_form == 0) {
        ps = sh_psingle(p, g, fcld, fray);
    } else {
        ps = 0.0;
#pragma unroll
        for (int m = 0; m < S; ++m) ps = ps + ws[m] * Pu0[m] * Pu1[m];
    }
    double a[4], b[4];
#pragma unroll
    for (int m = 0; m < S; ++m) {
        a[m] = (2 * m + 1) - om * wm[m];
        b[m] = (f0 * (om * ws[m])) * Pu0[m] / (4 * PB_PI);
    }
    const double taul = __ldg(p.tau + iv);
    const double inv_u0 = 1.0 / u0, inv_u1 = 1.0 / u1;
    et = pbm::kexp(-taul * inv_u0);
    // exp(-clip35(x)) for x >= -35 is max(exp(-x), exp(-35))
    const double EXPM35 = 6.305116760146989e-16, EXPP35 = 1586013452313430.8;
    const double et_c = fmin(fmax(et, EXPM35), EXPP35);
    double c[S], wgt[S], Nsum;
    if (S == 2) {
        // setup_2_stream_fluxes, fluxes.py:3239-3265
        const double Del = inv_u0 * inv_u0 - a[0] * a[1];
        const double iD = pbm::krcp(Del);
        const double eta0 = (b[1] * inv_u0 - a[1] * b[0]) * iD;
        const double eta1 = (b[0] * inv_u0 - a[0] * b[1]) * iD;
        const double lam = sqrt(a[0] * a[1]);
        const double e = pbm::kexp(-clip35(lam * dt));
        const double q = lam * pbm::krcp(a[1]);
        const double Q1 = (0.5 + q) * TWO_PI, Q2 = (0.5 - q) * TWO_PI;
        o.T[0][0] = Q1; o.T[0][1] = Q2;
        o.T[1][0] = Q2; o.T[1][1] = Q1;
        o.cs[0] = e; o.cs[1] = pbm::krcp(e);
        const double zmn = (0.5 * eta0 - eta1) * TWO_PI, zpl = (0.5 * eta0 + eta1) * TWO_PI;
        // SH2 uses the UNclipped exp(-tau/u0) in the matrix (fluxes.py:3261)
        o.Zu[0] = zmn * eb; o.Zu[1] = zpl * eb;
        o.Zd[0] = zmn * et; o.Zd[1] = zpl * et;
        c[0] = inv_u1 + lam; c[1] = inv_u1 - lam;
        wgt[0] = wm[0] - wm[1] * Pu1[1] * q;
        wgt[1] = wm[0] + wm[1] * Pu1[1] * q;
        Nsum = wm[0] * eta0 + wm[1] * Pu1[1] * eta1;
    } else {
        // setup_4_stream_fluxes, fluxes.py:3387-3450
        const double beta = a[0] * a[1] + 4 * a[0] * a[3] / 9 + a[2] * a[3] / 9;
        const double gama = a[0] * a[1] * a[2] * a[3] / 9;
        const double disc = sqrt(beta * beta - 4 * gama);
        const double l1 = sqrt((beta + disc) / 2), l2 = sqrt((beta - disc) / 2);
        const double x2 = inv_u0 * inv_u0;
        const double iD = pbm::krcp(9 * (x2 * x2 - beta * x2 + gama));
        const double e0 = ((a[1] * b[0] - b[1] * inv_u0) * (a[2] * a[3] - 9 * x2) +
                           2 * (a[3] * b[2] - 2 * a[3] * b[0] - 3 * b[3] * inv_u0) * x2) * iD;
        const double e1 = ((a[0] * b[1] - b[0] * inv_u0) * (a[2] * a[3] - 9 * x2) -
                           2 * a[0] * (a[3] * b[2] - 3 * b[3] * inv_u0) * inv_u0) * iD;
        const double e2 = ((a[3] * b[2] - 3 * b[3] * inv_u0) * (a[0] * a[1] - x2) -
                           2 * a[3] * (a[0] * b[1] - b[0] * inv_u0) * inv_u0) * iD;
        const double e3 = ((a[2] * b[3] - 3 * b[2] * inv_u0) * (a[0] * a[1] - x2) +
                           2 * (3 * a[0] * b[1] - 2 * a[0] * b[3] - 3 * b[0] * inv_u0) * x2) * iD;
        const double z1pl = (e0 / 2 + e1 + 5 * e2 / 8) * TWO_PI, z1mn = (e0 / 2 - e1 + 5 * e2 / 8) * TWO_PI;
        const double z2pl = (-e0 / 8 + 5 * e2 / 8 + e3) * TWO_PI, z2mn = (-e0 / 8 + 5 * e2 / 8 - e3) * TWO_PI;
        const double x1 = pbm::kexp(-clip35(l1 * dt)), xx2 = pbm::kexp(-clip35(l2 * dt));
        const double il1 = pbm::krcp(l1), il2 = pbm::krcp(l2);
        const double R1 = -a[0] * il1, R2 = -a[0] * il2;
        const double Q1 = 0.5 * (a[0] * a[1] * il1 * il1 - 1), Q2 = 0.5 * (a[0] * a[1] * il2 * il2 - 1);
        const double m3 = -3 * pbm::krcp(2 * a[3]);
        const double S1 = m3 * (a[0] * a[1] * il1 - l1), S2 = m3 * (a[0] * a[1] * il2 - l2);
        const double p1pl = (0.5 + R1 + 5 * Q1 / 8) * TWO_PI, p2pl = (0.5 + R2 + 5 * Q2 / 8) * TWO_PI;
        const double q1pl = (-0.125 + 5 * Q1 / 8 + S1) * TWO_PI, q2pl = (-0.125 + 5 * Q2 / 8 + S2) * TWO_PI;
        const double p1mn = (0.5 - R1 + 5 * Q1 / 8) * TWO_PI, p2mn = (0.5 - R2 + 5 * Q2 / 8) * TWO_PI;
        const double q1mn = (-0.125 + 5 * Q1 / 8 - S1) * TWO_PI, q2mn = (-0.125 + 5 * Q2 / 8 - S2) * TWO_PI;
        // rows in matrix order (z1mn, z2mn, z1pl, z2pl): fluxes.py:3470-3543
        o.T[0][0] = p1mn; o.T[0][1] = p1pl; o.T[0][2] = p2mn; o.T[0][3] = p2pl;
        o.T[1][0] = q1mn; o.T[1][1] = q1pl; o.T[1][2] = q2mn; o.T[1][3] = q2pl;
        o.T[2][0] = p1pl; o.T[2][1] = p1mn; o.T[2][2] = p2pl; o.T[2][3] = p2mn;
        o.T[3][0] = q1pl; o.T[3][1] = q1mn; o.T[3][2] = q2pl; o.T[3][3] = q2mn;
        o.cs[0] = x1; o.cs[1] = pbm::krcp(x1); o.cs[2] = xx2; o.cs[3] = pbm::krcp(xx2);
        // SH4 clips tau/u0 at +-35 in the matrix (fluxes.py:3442)
        const double eb_c = fmin(fmax(eb, EXPM35), EXPP35);
        o.Zu[0] = z1mn * eb_c; o.Zu[1] = z2mn * eb_c; o.Zu[2] = z1pl * eb_c; o.Zu[3] = z2pl * eb_c;
        o.Zd[0] = z1mn * et_c; o.Zd[1] = z2mn * et_c; o.Zd[2] = z1pl * et_c; o.Zd[3] = z2pl * et_c;
        c[0] = inv_u1 + l1; c[1] = inv_u1 - l1; c[2] = inv_u1 + l2; c[3] = inv_u1 - l2;
        // sum_j w_multi_j P_j(u1) A_jk, A rows (1, +-R, Q, +-S): fluxes.py:2940-2942, :3601-3605
        const double w1 = wm[1] * Pu1[1], w2 = wm[2] * Pu1[2], w3 = wm[3] * Pu1[3], w0_ = wm[0] * Pu1[0];
        wgt[0] = w0_ + w1 * R1 + w2 * Q1 + w3 * S1;
        wgt[1] = w0_ - w1 * R1 + w2 * Q1 - w3 * S1;
        wgt[2] = w0_ + w1 * R2 + w2 * Q2 + w3 * S2;
        wgt[3] = w0_ - w1 * R2 + w2 * Q2 - w3 * S2;
        Nsum = w0_ * e0 + w1 * e1 + w2 * e2 + w3 * e3;
    }
    // source-function integration, fluxes.py:2900-2970
    const double imus = 1.0 / mus;
    const double expon1 = (1 - pbm::kexp(-clip35(mus * dt))) * imus * et_c;
#pragma unroll
    for (int k = 0; k < S; ++k)
        o.sw[k] = om * wgt[k] * ((1 - pbm::kexp(-clip35(c[k] * dt))) * pbm::krcp(c[k])) * inv_u1;
    const double e1m = pbm::kexp(-clip35(mus * __ldg(p.dtau_og + il)));
    const double single = __ldg(p.w0_og + il) * f0 / (4 * PB_PI) * ps * (1 - e1m) *
                          pbm::kexp(-__ldg(p.tau_og + iv) * inv_u0) * imus;
    o.sconst = (om * (Nsum * expon1) + single) * inv_u1;
    o.xa = pbm::kexp(-dt * inv_u1);
}

// One out-of-line instance of sh_layer for the flx = 1 kernel, which evaluates every layer twice (elimination sweep,
// substitution pass).  Inlined twice, the two copies get different FMA contractions from the compiler; in
// near-resonant columns (1/u0^2 ~ a0 a1: the particular solution is ~1e7 x the flux it leaves after cancelling against
// the homogeneous part) a last-bit difference in Del = 1/u0^2 - a0 a1 moves Z by 1e-12 relative, and fluxes formed
// with one copy's Z from an X solved against the other's came out 1e-9 off (measured: 240 x the level-flux tolerance).
// Sharing the code makes the second evaluation bit-identical to the first.
template <int S>
__device__ __noinline__ void sh_layer_shared(const ShParams &p, int64_t il, int64_t iv, double u0, double u1, double f0,
                                             const double (&Pu0)[4], const double (&Pu1)[4], double mus, int angle_index,
                                             double &et, double eb, Layer<S> &o, double *fdm_out)
{
    sh_layer<S>(p, il, iv, u0, u1, f0, Pu0, Pu1, mus, angle_index, et, eb, o, fdm_out);
}

// FLX: also return the layer fluxes calculate_flux(F, G, X) (fluxes.py:2889-2890, F and G of :3311-3331 / :3551-3598).
// They need the whole solution X, which the TOA-only sweep never forms: the S pivot rows of every eliminated layer go
// to a scratch array in HBM (wavelength fastest, so the stores coalesce), the closed top system gives X_0, and a
// second, top-down pass substitutes X_{l+1} from X_l and the stored rows (the substitution LAPACK's dgbtrs does with
// U) and evaluates level 0 = T_0 X_0 + Zd_0, level l+1 = Fb_l X_l + Zu_l.  Not a hot path (calculate_fluxes is 'off'
// by default, justdoit.py:4638): the layer quantities are simply recomputed in the second pass.
template <int S, bool FLX = false>
__global__ void __launch_bounds__(128) sh_reflected_kernel(ShParams p)
{
    constexpr int H = S / 2, NR = H + S, NC = 2 * S + 1;
    extern __shared__ double s_int[];
    const int lane = threadIdx.x;
    const int w = (FLX ? p.w_begin : 0) + blockIdx.x * kWaves + lane;
    const int a = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    const bool active = (w < (FLX ? p.w_end : p.W)) && (a < p.G);
    // FLX: this thread's column of the pivot-row scratch
    const int64_t pcols = FLX ? (int64_t)gridDim.z * p.G * p.wcap : 0;
    double *pcol = FLX ? p.piv + ((int64_t)b * p.G + a) * p.wcap + (w - p.w_begin) : nullptr;
    double result = 0.0;
    if (active) {
        const int L = p.L;
        const int64_t ld = p.ld;
        const int64_t ol = (int64_t)b * p.bs_layer + w, ov = (int64_t)b * p.bs_level + w;
        const int64_t ow = (int64_t)b * p.bs_wave + w;
        const double u0 = p.ubar0[a], u1 = p.ubar1[a];
        const double f0 = p.f0pi ? p.f0pi[ow] : 1.0;
        const double r = p.surf ? p.surf[ow] : 0.0;
        const double bt = p.btop ? p.btop[ow] : 0.0;
        double Pu0[4], Pu1[4];
        {
            const double m0 = -u0;  // legP(-u0), legP(u1): fluxes.py:2800-2801, :3643
            Pu0[0] = 1; Pu0[1] = m0; Pu0[2] = (3 * m0 * m0 - 1) / 2; Pu0[3] = (5 * m0 * m0 * m0 - 3 * m0) / 2;
            Pu1[0] = 1; Pu1[1] = u1; Pu1[2] = (3 * u1 * u1 - 1) / 2; Pu1[3] = (5 * u1 * u1 * u1 - 3 * u1) / 2;
        }
        const double mus = (u1 + u0) / (u1 * u0);
        double eb = pbm::kexp(-__ldg(p.tau + ov + (int64_t)L * ld) / u0);  // exp(-tau_L/u0)
        const double b_surface = (0. + r * u0 * f0 * eb);

        double C[H][S + 1];  // carried constraints on the current layer's unknowns
        double J[S + 1];     // intensity at the top of the processed stack: J[0..S-1].X_l + J[S]
        double Tn[S][S], Zdn[S];
        double *fdm_out = (p.fdm_out && a == p.G - 1) ? p.fdm_out : nullptr;
        for (int l = L - 1; l >= 0; --l) {
#ifndef PB_SH_NO_PREFETCH
            // ncu (profiles/r1_sh4.summary.json): the sweep stalls on its own global loads (long_scoreboard
            // 3.8 of 7.4 cycles per instruction at 10 % occupancy, no registers left for a software pipeline):
            // pull the next layer's rows towards the SM while this layer is eliminated
            if (l > 0) {
                const int64_t il = ol + (int64_t)(l - 1) * ld, iv = ov + (int64_t)(l - 1) * ld;
                const double *rows[10] = {p.w0 + il, p.dtau + il, p.fcld + il, p.fray + il, p.cosb_og + il, p.fdm + il,
                                          p.dtau_og + il, p.w0_og + il, p.tau + iv, p.tau_og + iv};
#pragma unroll
                for (int i = 0; i < 10; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(rows[i]));
            }
#endif
            Layer<S> y;
            double et;
            if (FLX)
                sh_layer_shared<S>(p, ol + (int64_t)l * ld, ov + (int64_t)l * ld, u0, u1, f0, Pu0, Pu1, mus, a, et, eb, y,
                                   fdm_out ? fdm_out + ((int64_t)b * L + l) * p.W + w : nullptr);
            else
                sh_layer<S>(p, ol + (int64_t)l * ld, ov + (int64_t)l * ld, u0, u1, f0, Pu0, Pu1, mus, a, et, eb, y,
                            fdm_out ? fdm_out + ((int64_t)b * L + l) * p.W + w : nullptr);
            if (l == L - 1) {
                // surface rows (fluxes.py:3286-3289 | :3483-3494) and I_L = flux_bot/pi (:2891, :2967)
#pragma unroll
                for (int h = 0; h < H; ++h) {
#pragma unroll
                    for (int c = 0; c < S; ++c)
                        C[h][c] = y.T[H + h][c] * y.cs[c] - r * (y.T[h][c] * y.cs[c]);
                    const double bs = (h == 0) ? b_surface : -b_surface / 4;
                    C[h][S] = bs - y.Zu[H + h] + r * y.Zu[h];
                }
#pragma unroll
                for (int c = 0; c < S; ++c) J[c] = (y.T[H][c] * y.cs[c]) / PB_PI;
                J[S] = y.Zu[H] / PB_PI;
            } else {
                // window over [X_{l+1} | X_l | rhs]: carried rows, then the S interface rows
                double R[NR][NC], F[NC];
#pragma unroll
                for (int h = 0; h < H; ++h) {
#pragma unroll
                    for (int c = 0; c < S; ++c) { R[h][c] = C[h][c]; R[h][S + c] = 0.0; }
                    R[h][2 * S] = C[h][S];
                }
#pragma unroll
                for (int i = 0; i < S; ++i) {
#pragma unroll
                    for (int c = 0; c < S; ++c) { R[H + i][c] = -Tn[i][c]; R[H + i][S + c] = y.T[i][c] * y.cs[c]; }
                    R[H + i][2 * S] = Zdn[i] - y.Zu[i];
                }
#pragma unroll
                for (int c = 0; c < S; ++c) { F[c] = J[c]; F[S + c] = 0.0; }
                F[2 * S] = J[S];
                eliminate<S, NR, NC>(R, F);
                if (FLX) {
                    // pivot row k: sum_{c >= k} R[k][c] x_c = R[k][2S] over x = [X_{l+1} | X_l]
                    double *o = pcol + (int64_t)l * PivRows<S>::N * pcols;
                    int e = 0;
#pragma unroll
                    for (int k = 0; k < S; ++k)
#pragma unroll
                        for (int c = k; c < NC; ++c, ++e) o[(int64_t)e * pcols] = R[k][c];
                }
#pragma unroll
                for (int h = 0; h < H; ++h) {
#pragma unroll
                    for (int c = 0; c < S; ++c) C[h][c] = R[S + h][S + c];
                    C[h][S] = R[S + h][2 * S];
                }
#pragma unroll
                for (int c = 0; c < S; ++c) J[c] = F[S + c];
                J[S] = F[2 * S];
            }
            // xint[l] = xint[l+1] exp(-dtau/u1) + intgrl_per_layer / u1   (fluxes.py:2968-2970)
#pragma unroll
            for (int c = 0; c < S; ++c) J[c] = fma(y.xa, J[c], y.sw[c]);
            J[S] = fma(y.xa, J[S], y.sconst);
#pragma unroll
            for (int i = 0; i < S; ++i) {
#pragma unroll
                for (int c = 0; c < S; ++c) Tn[i][c] = y.T[i][c];
                Zdn[i] = y.Zd[i];
            }
            eb = et;
        }
        // top boundary rows (fluxes.py:3280-3283 | :3469-3480) close the system
        {
            double R[S][S + 1], F[S + 1];
#pragma unroll
            for (int h = 0; h < H; ++h) {
#pragma unroll
                for (int c = 0; c <= S; ++c) R[h][c] = C[h][c];
#pragma unroll
                for (int c = 0; c < S; ++c) R[H + h][c] = Tn[h][c];
                R[H + h][S] = ((h == 0) ? bt : -bt / 4) - Zdn[h];
            }
#pragma unroll
            for (int c = 0; c <= S; ++c) F[c] = J[c];
            eliminate<S, S, S + 1>(R, F);
            result = F[S];
            if (FLX) {
                // X_0 from the triangular top system, then the top-down substitution pass
                double X[S];
#pragma unroll
                for (int k = S - 1; k >= 0; --k) {
                    double sacc = R[k][S];
#pragma unroll
                    for (int c = k + 1; c < S; ++c) sacc = fma(-R[k][c], X[c], sacc);
                    X[k] = sacc / R[k][k];
                }
                const int V = L + 1;
                double *fo = p.flux + ((int64_t)b * p.G + a) * S * V * p.W + w;
                for (int l = 0; l < L; ++l) {
                    Layer<S> y;
                    double et;
                    // exp(-tau_{l+1}/u0) exactly as the sweep above saw it (division at the surface, 1/u0 elsewhere)
                    const double tb = __ldg(p.tau + ov + (int64_t)(l + 1) * ld);
                    const double ebl = (l == L - 1) ? pbm::kexp(-tb / u0) : pbm::kexp(-tb * (1.0 / u0));
                    sh_layer_shared<S>(p, ol + (int64_t)l * ld, ov + (int64_t)l * ld, u0, u1, f0, Pu0, Pu1, mus, a, et, ebl,
                                       y, nullptr);
                    if (l == 0) {
#pragma unroll
                        for (int i = 0; i < S; ++i) {
                            double acc = 0.0;
#pragma unroll
                            for (int c = 0; c < S; ++c) acc = fma(y.T[i][c], X[c], acc);
                            fo[(int64_t)i * p.W] = acc + y.Zd[i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        double acc = 0.0;
#pragma unroll
                        for (int c = 0; c < S; ++c) acc = fma(y.T[i][c] * y.cs[c], X[c], acc);
                        fo[(int64_t)(S * (l + 1) + i) * p.W] = acc + y.Zu[i];
                    }
                    if (l < L - 1) {
                        // X_{l+1} from the pivot rows of the window [X_{l+1} | X_l]
                        const double *o = pcol + (int64_t)l * PivRows<S>::N * pcols;
                        double Rk[S][NC];
                        int e = 0;
#pragma unroll
                        for (int k = 0; k < S; ++k)
#pragma unroll
                            for (int c = k; c < NC; ++c, ++e) Rk[k][c] = o[(int64_t)e * pcols];
                        double Xn[S];
#pragma unroll
                        for (int k = S - 1; k >= 0; --k) {
                            double sacc = Rk[k][2 * S];
#pragma unroll
                            for (int c = 0; c < S; ++c) sacc = fma(-Rk[k][S + c], X[c], sacc);
#pragma unroll
                            for (int c = k + 1; c < S; ++c) sacc = fma(-Rk[k][c], Xn[c], sacc);
                            Xn[k] = sacc / Rk[k][k];
                        }
#pragma unroll
                        for (int c = 0; c < S; ++c) X[c] = Xn[c];
                    }
                }
            }
        }
        if (p.xint) p.xint[((int64_t)b * p.G + a) * p.W + w] = result;
    }
    if (p.fuse_albedo) {
        s_int[threadIdx.y * kWaves + lane] = result;
        __syncthreads();
        if (threadIdx.y == 0 && w < p.W) {
            double acc = 0.0;
            for (int aa = 0; aa < p.G; ++aa) {
                const int ig = aa / p.nt, it = aa - ig * p.nt;
                acc = acc + s_int[aa * kWaves + lane] * p.gweight[ig] * p.tweight[it];
            }
            const double sym = (p.nt == 1) ? 2.0 * PB_PI : 1.0;
            const double f0 = p.f0pi ? p.f0pi[(int64_t)b * p.bs_wave + w] : 1.0;
            p.albedo[(int64_t)b * p.W + w] = sym * 0.5 * acc / f0 * (p.cos_theta + 1.0);
        }
    }
}

#include "sh_reflected_tile.cuh"

// ---------------------------------------------------------------------------------------
// Thermal SH (get_thermal_SH, fluxes.py:2979-3186): same block structure with the linear-in-tau
// Planck particular solution (calculation == 1 branches, :3266-3270 / :3451-3459); the upward
// intensity starts from the surface emission instead of flux_bot/pi.
// ---------------------------------------------------------------------------------------
struct ShThermParams {
    int L, W, G, nt;
    int64_t ld, bs_layer, bs_wave;
    const double *dtau, *w0, *cosb_og;
    const double *wno, *surf;
    const double *tlevel, *plevel;  // [B][V]
    const double *ubar1, *gweight, *tweight;
    const int *ff_zero;  // device flag: 1 when cosb == cosb_og everywhere (ff = 0, fluxes.py:3044)
    int hard_surface;
    double *xint, *thermal;
    int fuse;
};

__global__ void arrays_equal_kernel(int64_t n, const double *a, const double *b, int *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(a[i] == b[i])) *flag = 0;
}

template <int S>
__device__ __forceinline__ void sh_thermal_layer(double om, double dt, double g, bool ff_zero, double B0,
                                                 double B1, double u1, const double (&Pu1)[4], Layer<S> &o)
{
    const double TWO_PI = 2 * PB_PI;
    double wm[4], a[4];
    {
        double ff = 0. * g;
        if (!ff_zero) {
            ff = 1.0;
#pragma unroll
            for (int m = 0; m < S; ++m) ff *= g;
        }
        double gm = 1.0;
#pragma unroll
        for (int m = 0; m < S; ++m) {
            wm[m] = (2 * m + 1) * (gm - ff) / (1 - ff);
            a[m] = (2 * m + 1) - om * wm[m];
            gm *= g;
        }
    }
    const double inv_u1 = 1.0 / u1;
    const double src = (1 - om) * pbm::krcp(a[0]);  // (1 - w0)/a[0]
    double c[S], wgt[S];
    if (S == 2) {
        const double lam = sqrt(a[0] * a[1]);
        const double e = pbm::kexp(-clip35(lam * dt));
        const double q = lam * pbm::krcp(a[1]);
        const double Q1 = (0.5 + q) * TWO_PI, Q2 = (0.5 - q) * TWO_PI;
        o.T[0][0] = Q1; o.T[0][1] = Q2;
        o.T[1][0] = Q2; o.T[1][1] = Q1;
        o.cs[0] = e; o.cs[1] = pbm::krcp(e);
        const double ia1 = pbm::krcp(a[1]);
        o.Zd[0] = (src * (B0 / 2 - B1 * ia1)) * TWO_PI;
        o.Zu[0] = (src * (B0 / 2 - B1 * ia1 + B1 * dt / 2)) * TWO_PI;
        o.Zd[1] = (src * (B0 / 2 + B1 * ia1)) * TWO_PI;
        o.Zu[1] = (src * (B0 / 2 + B1 * ia1 + B1 * dt / 2)) * TWO_PI;
        c[0] = inv_u1 + lam; c[1] = inv_u1 - lam;
        wgt[0] = wm[0] - wm[1] * Pu1[1] * q;
        wgt[1] = wm[0] + wm[1] * Pu1[1] * q;
    } else {
        const double beta = a[0] * a[1] + 4 * a[0] * a[3] / 9 + a[2] * a[3] / 9;
        const double gama = a[0] * a[1] * a[2] * a[3] / 9;
        const double disc = sqrt(beta * beta - 4 * gama);
        const double l1 = sqrt((beta + disc) / 2), l2 = sqrt((beta - disc) / 2);
        const double x1 = pbm::kexp(-clip35(l1 * dt)), xx2 = pbm::kexp(-clip35(l2 * dt));
        const double il1 = pbm::krcp(l1), il2 = pbm::krcp(l2);
        const double R1 = -a[0] * il1, R2 = -a[0] * il2;
        const double Q1 = 0.5 * (a[0] * a[1] * il1 * il1 - 1), Q2 = 0.5 * (a[0] * a[1] * il2 * il2 - 1);
        const double m3 = -3 * pbm::krcp(2 * a[3]);
        const double S1 = m3 * (a[0] * a[1] * il1 - l1), S2 = m3 * (a[0] * a[1] * il2 - l2);
        const double p1pl = (0.5 + R1 + 5 * Q1 / 8) * TWO_PI, p2pl = (0.5 + R2 + 5 * Q2 / 8) * TWO_PI;
        const double q1pl = (-0.125 + 5 * Q1 / 8 + S1) * TWO_PI, q2pl = (-0.125 + 5 * Q2 / 8 + S2) * TWO_PI;
        const double p1mn = (0.5 - R1 + 5 * Q1 / 8) * TWO_PI, p2mn = (0.5 - R2 + 5 * Q2 / 8) * TWO_PI;
        const double q1mn = (-0.125 + 5 * Q1 / 8 - S1) * TWO_PI, q2mn = (-0.125 + 5 * Q2 / 8 - S2) * TWO_PI;
        o.T[0][0] = p1mn; o.T[0][1] = p1pl; o.T[0][2] = p2mn; o.T[0][3] = p2pl;
        o.T[1][0] = q1mn; o.T[1][1] = q1pl; o.T[1][2] = q2mn; o.T[1][3] = q2pl;
        o.T[2][0] = p1pl; o.T[2][1] = p1mn; o.T[2][2] = p2pl; o.T[2][3] = p2mn;
        o.T[3][0] = q1pl; o.T[3][1] = q1mn; o.T[3][2] = q2pl; o.T[3][3] = q2mn;
        o.cs[0] = x1; o.cs[1] = pbm::krcp(x1); o.cs[2] = xx2; o.cs[3] = pbm::krcp(xx2);
        const double ia1 = pbm::krcp(a[1]);
        const double s4 = -0.5 * (1 - om) / (4 * a[0]);
        o.Zu[0] = src * (B0 / 2 - B1 * ia1 + B1 * dt / 2) * TWO_PI;
        o.Zu[1] = s4 * (B0 + B1 * dt) * TWO_PI;
        o.Zu[2] = src * (B0 / 2 + B1 * ia1 + B1 * dt / 2) * TWO_PI;
        o.Zu[3] = o.Zu[1];
        o.Zd[0] = src * (B0 / 2 - B1 * ia1) * TWO_PI;
        o.Zd[1] = s4 * (B0)*TWO_PI;
        o.Zd[2] = src * (B0 / 2 + B1 * ia1) * TWO_PI;
        o.Zd[3] = o.Zd[1];
        c[0] = inv_u1 + l1; c[1] = inv_u1 - l1; c[2] = inv_u1 + l2; c[3] = inv_u1 - l2;
        const double w1 = wm[1] * Pu1[1], w2 = wm[2] * Pu1[2], w3 = wm[3] * Pu1[3], w0_ = wm[0] * Pu1[0];
        wgt[0] = w0_ + w1 * R1 + w2 * Q1 + w3 * S1;
        wgt[1] = w0_ - w1 * R1 + w2 * Q1 - w3 * S1;
        wgt[2] = w0_ + w1 * R2 + w2 * Q2 + w3 * S2;
        wgt[3] = w0_ - w1 * R2 + w2 * Q2 - w3 * S2;
    }
    // source function, fluxes.py:3105-3182
    const double ed = pbm::kexp(-dt * inv_u1);                  // exp(-dtau/u1)
    const double ed_n = (S == 4) ? fmax(ed, 6.305116760146989e-16) : ed;  // SH4 clips dtau/u1 at 35 in Nint
    const double pl = B0 * (1 - ed) + B1 * (u1 - (dt + u1) * ed);
    const double pl_n = B0 * (1 - ed_n) + B1 * (u1 - (dt + u1) * ed_n);
    const double N0 = wm[0] * (src * u1 * pl_n);
    const double N1 = wm[1] * Pu1[1] * (src * u1 * (B1 * (1 - ed_n) * pbm::krcp(a[1])));
#pragma unroll
    for (int k = 0; k < S; ++k)
        o.sw[k] = om * TWO_PI * wgt[k] * ((1 - pbm::kexp(-clip35(c[k] * dt))) * pbm::krcp(c[k])) * inv_u1;
    o.sconst = (om * (N0 + N1) * TWO_PI + TWO_PI * (1 - om) * u1 * pl) * inv_u1;
    o.xa = ed;
}

template <int S>
__global__ void __launch_bounds__(128) sh_thermal_kernel(ShThermParams p)
{
    constexpr int H = S / 2, NR = H + S, NC = 2 * S + 1;
    extern __shared__ double s_int[];
    const int lane = threadIdx.x;
    const int w = blockIdx.x * kWaves + lane;
    const int a = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    const bool active = (w < p.W) && (a < p.G);
    double result = 0.0;
    if (active) {
        const int L = p.L, V = p.L + 1;
        const int64_t ld = p.ld;
        const int64_t ol = (int64_t)b * p.bs_layer + w;
        const double *tl = p.tlevel + (int64_t)b * V, *pl = p.plevel + (int64_t)b * V;
        const double u1 = p.ubar1[a];
        const double r = p.surf ? p.surf[(int64_t)b * p.bs_wave + w] : 0.0;
        const bool ff_zero = *p.ff_zero != 0;
        const double kMu = 0.5;
        double Pu1[4];
        Pu1[0] = 1; Pu1[1] = u1; Pu1[2] = (3 * u1 * u1 - 1) / 2; Pu1[3] = (5 * u1 * u1 * u1 - 3 * u1) / 2;
        // blackbody(t, 1/wno), fluxes.py:1676-1680
        const double h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
        const double wl = 1.0 / p.wno[w];
        const double c1w = (2.0 * h * c * c) / pow(wl, 5.0), c2w = (h * c) / (wl * k);
        double Bbot = c1w * (1.0 / (exp(c2w / tl[L]) - 1.0));
        const double BL = Bbot;
        double C[H][S + 1], J[S + 1], Tn[S][S], Zdn[S];
        for (int l = L - 1; l >= 0; --l) {
            const int64_t il = ol + (int64_t)l * ld;
            const double dt = __ldg(p.dtau + il);
            const double Btop = c1w * (1.0 / (exp(c2w / tl[l]) - 1.0));
            const double B1 = (Bbot - Btop) / dt;
            Layer<S> y;
            sh_thermal_layer<S>(__ldg(p.w0 + il), dt, __ldg(p.cosb_og + il), ff_zero, Btop, B1, u1, Pu1, y);
            if (l == L - 1) {
                const double b_surface = p.hard_surface ? PB_PI * BL : PB_PI * (BL + B1 * kMu);
                const double b_surface4 = (-PB_PI * BL / 4);
#pragma unroll
                for (int hh = 0; hh < H; ++hh) {
#pragma unroll
                    for (int cc = 0; cc < S; ++cc)
                        C[hh][cc] = y.T[H + hh][cc] * y.cs[cc] - r * (y.T[hh][cc] * y.cs[cc]);
                    C[hh][S] = ((hh == 0) ? b_surface : b_surface4) - y.Zu[H + hh] + r * y.Zu[hh];
                }
#pragma unroll
                for (int cc = 0; cc < S; ++cc) J[cc] = 0.0;
                J[S] = p.hard_surface ? BL * 2 * PB_PI : (BL + B1 * u1) * 2 * PB_PI;  // fluxes.py:3171-3174
            } else {
                double R[NR][NC], F[NC];
#pragma unroll
                for (int hh = 0; hh < H; ++hh) {
#pragma unroll
                    for (int cc = 0; cc < S; ++cc) { R[hh][cc] = C[hh][cc]; R[hh][S + cc] = 0.0; }
                    R[hh][2 * S] = C[hh][S];
                }
#pragma unroll
                for (int i = 0; i < S; ++i) {
#pragma unroll
                    for (int cc = 0; cc < S; ++cc) { R[H + i][cc] = -Tn[i][cc]; R[H + i][S + cc] = y.T[i][cc] * y.cs[cc]; }
                    R[H + i][2 * S] = Zdn[i] - y.Zu[i];
                }
#pragma unroll
                for (int cc = 0; cc < S; ++cc) { F[cc] = J[cc]; F[S + cc] = 0.0; }
                F[2 * S] = J[S];
                eliminate<S, NR, NC>(R, F);
#pragma unroll
                for (int hh = 0; hh < H; ++hh) {
#pragma unroll
                    for (int cc = 0; cc < S; ++cc) C[hh][cc] = R[S + hh][S + cc];
                    C[hh][S] = R[S + hh][2 * S];
                }
#pragma unroll
                for (int cc = 0; cc < S; ++cc) J[cc] = F[S + cc];
                J[S] = F[2 * S];
            }
#pragma unroll
            for (int cc = 0; cc < S; ++cc) J[cc] = fma(y.xa, J[cc], y.sw[cc]);
            J[S] = fma(y.xa, J[S], y.sconst);
#pragma unroll
            for (int i = 0; i < S; ++i) {
#pragma unroll
                for (int cc = 0; cc < S; ++cc) Tn[i][cc] = y.T[i][cc];
                Zdn[i] = y.Zd[i];
            }
            Bbot = Btop;
        }
        {
            // top boundary (fluxes.py:3034-3035): isothermal overburden above the model top
            const double tau_top = __ldg(p.dtau + ol) * pl[0] / (pl[1] - pl[0]);
            const double b_top = PB_PI * (1.0 - exp(-tau_top / kMu)) * Bbot;
            double R[S][S + 1], F[S + 1];
#pragma unroll
            for (int hh = 0; hh < H; ++hh) {
#pragma unroll
                for (int cc = 0; cc <= S; ++cc) R[hh][cc] = C[hh][cc];
#pragma unroll
                for (int cc = 0; cc < S; ++cc) R[H + hh][cc] = Tn[hh][cc];
                R[H + hh][S] = ((hh == 0) ? b_top : -b_top / 4) - Zdn[hh];
            }
#pragma unroll
            for (int cc = 0; cc <= S; ++cc) F[cc] = J[cc];
            eliminate<S, S, S + 1>(R, F);
            result = F[S];
        }
        if (p.xint) p.xint[((int64_t)b * p.G + a) * p.W + w] = result;
    }
    if (p.fuse) {
        s_int[threadIdx.y * kWaves + lane] = result;
        __syncthreads();
        if (threadIdx.y == 0 && w < p.W) {
            double acc = 0.0;
            for (int aa = 0; aa < p.G; ++aa) {
                const int ig = aa / p.nt, it = aa - ig * p.nt;
                acc = acc + s_int[aa * kWaves + lane] * p.gweight[ig] * p.tweight[it];
            }
            const double sym = (p.nt == 1) ? 1.0 : 1 / (2 * PB_PI);
            p.thermal[(int64_t)b * p.W + w] = acc * sym;
        }
    }
}

__global__ void sh_compress_thermal_kernel(int W, int G, int nt, const double *flux, const double *gweight,
                                           const double *tweight, double *out)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (w >= W) return;
    double acc = 0.0;
    for (int a = 0; a < G; ++a) {
        const int ig = a / nt, it = a - ig * nt;
        acc = acc + flux[((int64_t)b * G + a) * W + w] * gweight[ig] * tweight[it];
    }
    out[(int64_t)b * W + w] = acc * ((nt == 1) ? 1.0 : 1 / (2 * PB_PI));
}

__global__ void sh_compress_kernel(int W, int G, int nt, double cos_theta, const double *xint,
                                   const double *gweight, const double *tweight, const double *f0pi,
                                   int64_t bs_wave, double *albedo)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (w >= W) return;
    double acc = 0.0;
    for (int a = 0; a < G; ++a) {
        const int ig = a / nt, it = a - ig * nt;
        acc = acc + xint[((int64_t)b * G + a) * W + w] * gweight[ig] * tweight[it];
    }
    const double sym = (nt == 1) ? 2.0 * PB_PI : 1.0;
    const double f0 = f0pi ? f0pi[(int64_t)b * bs_wave + w] : 1.0;
    albedo[(int64_t)b * W + w] = sym * 0.5 * acc / f0 * (cos_theta + 1.0);
}

} // namespace

extern "C" int pb_reflected_sh(pb_ctx *ctx, const pb_sh_args *a, int memspace)
{
    if (!ctx || !a) return PB_ERR_ARG;
    const int L = a->nlayer, W = a->nwno, G = a->numg * a->numt, V = L + 1;
    const int B = a->nbatch > 0 ? a->nbatch : 1;
    if (L < 1 || W < 0 || G < 1) return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: bad sizes L=%d W=%d G=%d", L, W, G);
    if (a->stream != 2 && a->stream != 4) return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: stream must be 2 or 4");
    if (a->flx != 0 && a->flx != 1) return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: flx must be 0 or 1");
    const bool flx = a->flx == 1;
    if (flx && !a->flux) return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: flx=1 needs the flux array [nbatch][numg*numt][stream*nlevel][nwno]");
    if (W == 0) return PB_OK;
    if (a->ld < W) return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: ld < nwno");
    if (!a->dtau || !a->tau || !a->w0 || !a->ftau_cld || !a->ftau_ray || !a->f_deltaM || !a->dtau_og ||
        !a->tau_og || !a->w0_og || !a->cosb_og || !a->ubar0 || !a->ubar1)
        return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: NULL input array");
    if (a->albedo && (!a->gweight || !a->tweight)) return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: albedo needs gweight/tweight");
    const int forms[7] = {a->w_single_form, a->w_multi_form, a->psingle_form, a->w_single_rayleigh,
                          a->w_multi_rayleigh, a->psingle_rayleigh, a->single_form};
    for (int i = 0; i < 7; ++i)
        if (forms[i] != 0 && forms[i] != 1) return pb_fail(ctx, PB_ERR_ARG, "reflected_sh: form flags must be 0 or 1");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const size_t nW = (size_t)W * sizeof(double);
    // SH4 with drift-free (OTHG) moment forms and the explicit single-scattering phase function: the matrix work is
    // angle-independent -> sh4_tile_kernel (sh_reflected_tile.cuh); PB_SH_TILE=0 forces the per-angle kernel
    const char *tile_e = getenv("PB_SH_TILE");
    const int tile_env = tile_e ? atoi(tile_e) : 1;
    const bool tile = tile_env != 0 && a->stream == 4 && a->w_single_form == 1 && a->w_multi_form == 1 &&
                      a->single_form == 0 && !a->f_deltaM_out && !flx;
    const bool fuse = a->albedo && G <= (tile ? 8 : 4);
    const bool need_xint = a->xint_at_top || (a->albedo && !fuse);
    size_t need = 16 * 256 + 4 * pb_align((size_t)G * 8);
    if (host) {
        need += 8 * pb_align((size_t)B * L * nW) + 2 * pb_align((size_t)B * V * nW) + 3 * pb_align(B * nW);
        need += pb_align((size_t)B * G * nW) + pb_align(B * nW) + pb_align((size_t)B * L * nW);
    } else {
        need += pb_align((size_t)B * G * nW);
    }
    // flx = 1: pivot-row scratch for a chunk of `wcap` wavelengths (<= ~1 GB; multiple of the 32-wavelength tile)
    const int S = a->stream;
    const size_t piv_per_wave = (size_t)(L - 1) * (S == 2 ? PivRows<2>::N : PivRows<4>::N) * B * G * sizeof(double);
    int wcap = (W + kWaves - 1) / kWaves * kWaves;
    if (flx) {
        if (piv_per_wave > 0) {
            const size_t fit = ((size_t)1 << 30) / piv_per_wave / kWaves * kWaves;
            if (fit < (size_t)wcap) wcap = fit < (size_t)kWaves ? kWaves : (int)fit;
        }
        if (const char *e = getenv("PB_SH_FLX_WCAP")) {   // tests: force several chunks on a small case
            const int v = atoi(e) / kWaves * kWaves;
            if (v >= kWaves && v < wcap) wcap = v;
        }
        need += pb_align(piv_per_wave * wcap + 256);
        if (host) need += pb_align((size_t)B * G * S * V * nW);
    }
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, need));
    PB_TRY(pb_pinned_reserve(ctx, 4 * ((size_t)G + 16) * sizeof(double)));
    ShParams p;
    memset(&p, 0, sizeof(p));
    p.L = L; p.W = W; p.G = G; p.nt = a->numt;
    int64_t ldo;
    const int64_t rowsL = (int64_t)B * L, rowsV = (int64_t)B * V;
    PB_TRY(pb_stage_in(ctx, a->dtau, memspace, rowsL, W, a->ld, &p.dtau, &ldo));
    PB_TRY(pb_stage_in(ctx, a->w0, memspace, rowsL, W, a->ld, &p.w0, &ldo));
    PB_TRY(pb_stage_in(ctx, a->ftau_cld, memspace, rowsL, W, a->ld, &p.fcld, &ldo));
    PB_TRY(pb_stage_in(ctx, a->ftau_ray, memspace, rowsL, W, a->ld, &p.fray, &ldo));
    PB_TRY(pb_stage_in(ctx, a->f_deltaM, memspace, rowsL, W, a->ld, &p.fdm, &ldo));
    if (a->dtau_og == a->dtau) p.dtau_og = p.dtau; else PB_TRY(pb_stage_in(ctx, a->dtau_og, memspace, rowsL, W, a->ld, &p.dtau_og, &ldo));
    if (a->w0_og == a->w0) p.w0_og = p.w0; else PB_TRY(pb_stage_in(ctx, a->w0_og, memspace, rowsL, W, a->ld, &p.w0_og, &ldo));
    PB_TRY(pb_stage_in(ctx, a->cosb_og, memspace, rowsL, W, a->ld, &p.cosb_og, &ldo));
    PB_TRY(pb_stage_in(ctx, a->tau, memspace, rowsV, W, a->ld, &p.tau, &ldo));
    if (a->tau_og == a->tau) p.tau_og = p.tau; else PB_TRY(pb_stage_in(ctx, a->tau_og, memspace, rowsV, W, a->ld, &p.tau_og, &ldo));
    PB_TRY(pb_stage_in(ctx, a->surf_reflect, memspace, B, W, W, &p.surf, &ldo));
    PB_TRY(pb_stage_in(ctx, a->F0PI, memspace, B, W, W, &p.f0pi, &ldo));
    PB_TRY(pb_stage_in(ctx, a->b_top, memspace, B, W, W, &p.btop, &ldo));
    p.ld = host ? W : a->ld;
    p.bs_layer = (int64_t)L * p.ld; p.bs_level = (int64_t)V * p.ld; p.bs_wave = W;
    PB_TRY(pb_upload_small(ctx, a->ubar0, G, &p.ubar0));
    PB_TRY(pb_upload_small(ctx, a->ubar1, G, &p.ubar1));
    if (a->gweight) PB_TRY(pb_upload_small(ctx, a->gweight, a->numg, &p.gweight));
    if (a->tweight) PB_TRY(pb_upload_small(ctx, a->tweight, a->numt, &p.tweight));
    p.cos_theta = a->cos_theta; p.frac_a = a->frac_a; p.frac_b = a->frac_b; p.frac_c = a->frac_c;
    p.cback = a->constant_back; p.cfwd = a->constant_forward;
    p.wsf = a->w_single_form; p.wmf = a->w_multi_form; p.psf = a->psingle_form;
    p.wsr = a->w_single_rayleigh; p.wmr = a->w_multi_rayleigh; p.psr = a->psingle_rayleigh;
    p.single_form = a->single_form;
    double *d_xint = nullptr, *d_alb = nullptr, *d_fdm = nullptr;
    if (host) {
        if (need_xint) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * nW, (void **)&d_xint));
        if (a->albedo) PB_TRY(pb_arena_alloc(ctx, B * nW, (void **)&d_alb));
        if (a->f_deltaM_out) PB_TRY(pb_arena_alloc(ctx, (size_t)B * L * nW, (void **)&d_fdm));
    } else {
        d_xint = a->xint_at_top;
        if (!d_xint && need_xint) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * nW, (void **)&d_xint));
        d_alb = a->albedo;
        d_fdm = a->f_deltaM_out;
    }
    p.xint = d_xint; p.albedo = d_alb; p.fdm_out = d_fdm; p.fuse_albedo = fuse ? 1 : 0;
    double *d_flux = nullptr;
    if (flx) {
        if (host) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * S * V * nW, (void **)&d_flux));
        else d_flux = a->flux;
        PB_TRY(pb_arena_alloc(ctx, piv_per_wave * wcap + 256, (void **)&p.piv));
        p.flux = d_flux; p.wcap = wcap;
    }
    PB_TRY(pb_upload_flush(ctx));
    const int ay = G < 4 ? G : 4;
    dim3 block(kWaves, ay, 1);
    dim3 grid((W + kWaves - 1) / kWaves, (G + ay - 1) / ay, B);
    const size_t smem = fuse ? (size_t)ay * kWaves * sizeof(double) : 0;
    if (flx) {
        // one launch per wavelength chunk; launches of a stream run in order, so they can share the scratch
        for (int wb = 0; wb < W; wb += wcap) {
            p.w_begin = wb;
            p.w_end = wb + wcap < W ? wb + wcap : W;
            dim3 cgrid((p.w_end - p.w_begin + kWaves - 1) / kWaves, (G + ay - 1) / ay, B);
            if (S == 2) sh_reflected_kernel<2, true><<<cgrid, block, smem, ctx->stream>>>(p);
            else sh_reflected_kernel<4, true><<<cgrid, block, smem, ctx->stream>>>(p);
            PB_CHECK_LAUNCH(ctx);
        }
    } else if (tile) {
        const int nwa = G < 8 ? G : 8;
        const size_t tsmem = ((size_t)pbm::kExpTabDoubles + (size_t)2 * TS_N * 32) * sizeof(double);
        dim3 tgrid((W + 31) / 32, (G + nwa - 1) / nwa, B);
        PB_CUDA(ctx, pb_ensure_smem(ctx, sh4_tile_kernel, tsmem));
        sh4_tile_kernel<<<tgrid, (nwa + 1) * 32, tsmem, ctx->stream>>>(p);
    } else if (a->stream == 2) sh_reflected_kernel<2><<<grid, block, smem, ctx->stream>>>(p);
    else sh_reflected_kernel<4><<<grid, block, smem, ctx->stream>>>(p);
    if (!flx) PB_CHECK_LAUNCH(ctx);
    if (a->albedo && !fuse) {
        dim3 g2((W + 127) / 128, B);
        sh_compress_kernel<<<g2, 128, 0, ctx->stream>>>(W, G, a->numt, a->cos_theta, d_xint, p.gweight, p.tweight,
                                                        p.f0pi, p.bs_wave, d_alb);
        PB_CHECK_LAUNCH(ctx);
    }
    if (host) {
        if (a->xint_at_top) PB_CUDA(ctx, cudaMemcpyAsync(a->xint_at_top, d_xint, (size_t)B * G * nW, cudaMemcpyDeviceToHost, ctx->stream));
        if (a->albedo) PB_CUDA(ctx, cudaMemcpyAsync(a->albedo, d_alb, B * nW, cudaMemcpyDeviceToHost, ctx->stream));
        if (a->f_deltaM_out) PB_CUDA(ctx, cudaMemcpyAsync(a->f_deltaM_out, d_fdm, (size_t)B * L * nW, cudaMemcpyDeviceToHost, ctx->stream));
        if (flx) PB_CUDA(ctx, cudaMemcpyAsync(a->flux, d_flux, (size_t)B * G * S * V * nW, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}

extern "C" int pb_thermal_sh(pb_ctx *ctx, const pb_thermal_sh_args *a, int memspace)
{
    if (!ctx || !a) return PB_ERR_ARG;
    const int L = a->nlayer, W = a->nwno, G = a->numg * a->numt, V = L + 1;
    const int B = a->nbatch > 0 ? a->nbatch : 1;
    if (L < 1 || W < 0 || G < 1) return pb_fail(ctx, PB_ERR_ARG, "thermal_sh: bad sizes L=%d W=%d G=%d", L, W, G);
    if (a->stream != 2 && a->stream != 4) return pb_fail(ctx, PB_ERR_ARG, "thermal_sh: stream must be 2 or 4");
    if (a->flx != 0) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "thermal_sh: flx=1 is not implemented (it raises in the reference too)");
    if (W == 0) return PB_OK;
    if (a->ld < W) return pb_fail(ctx, PB_ERR_ARG, "thermal_sh: ld < nwno");
    if (!a->dtau || !a->w0 || !a->cosb || !a->cosb_og || !a->wno || !a->tlevel || !a->plevel || !a->ubar1)
        return pb_fail(ctx, PB_ERR_ARG, "thermal_sh: NULL input array");
    if (a->thermal && (!a->gweight || !a->tweight)) return pb_fail(ctx, PB_ERR_ARG, "thermal_sh: thermal output needs gweight/tweight");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const size_t nW = (size_t)W * sizeof(double);
    const bool fuse = a->thermal && G <= 4;
    const bool need_x = a->xint_at_top || (a->thermal && !fuse);
    size_t need = 32 * 256 + 2 * pb_align((size_t)B * V * 8) + 3 * pb_align((size_t)G * 8);
    if (host) need += 4 * pb_align((size_t)B * L * nW) + 2 * pb_align(nW) + pb_align(B * nW) + pb_align((size_t)B * G * nW) + pb_align(B * nW);
    else need += pb_align((size_t)B * G * nW);
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, need));
    PB_TRY(pb_pinned_reserve(ctx, (2 * (size_t)B * V + 3 * (size_t)G + 64) * sizeof(double)));
    ShThermParams p;
    memset(&p, 0, sizeof(p));
    p.L = L; p.W = W; p.G = G; p.nt = a->numt;
    int64_t ldo;
    const double *d_cosb;
    PB_TRY(pb_stage_in(ctx, a->dtau, memspace, (int64_t)B * L, W, a->ld, &p.dtau, &ldo));
    PB_TRY(pb_stage_in(ctx, a->w0, memspace, (int64_t)B * L, W, a->ld, &p.w0, &ldo));
    PB_TRY(pb_stage_in(ctx, a->cosb_og, memspace, (int64_t)B * L, W, a->ld, &p.cosb_og, &ldo));
    if (a->cosb == a->cosb_og) d_cosb = p.cosb_og;
    else PB_TRY(pb_stage_in(ctx, a->cosb, memspace, (int64_t)B * L, W, a->ld, &d_cosb, &ldo));
    PB_TRY(pb_stage_in(ctx, a->wno, memspace, 1, W, W, &p.wno, &ldo));
    PB_TRY(pb_stage_in(ctx, a->surf_reflect, memspace, B, W, W, &p.surf, &ldo));
    p.ld = host ? W : a->ld;
    p.bs_layer = (int64_t)L * p.ld; p.bs_wave = W;
    PB_TRY(pb_upload_small(ctx, a->tlevel, (size_t)B * V, &p.tlevel));
    PB_TRY(pb_upload_small(ctx, a->plevel, (size_t)B * V, &p.plevel));
    PB_TRY(pb_upload_small(ctx, a->ubar1, G, &p.ubar1));
    if (a->gweight) PB_TRY(pb_upload_small(ctx, a->gweight, a->numg, &p.gweight));
    if (a->tweight) PB_TRY(pb_upload_small(ctx, a->tweight, a->numt, &p.tweight));
    p.hard_surface = a->hard_surface;
    PB_TRY(pb_upload_flush(ctx));
    // ff = 0 iff cosb == cosb_og everywhere (np.array_equal, fluxes.py:3044): decided on the device
    int *d_flag;
    PB_TRY(pb_arena_alloc(ctx, sizeof(int), (void **)&d_flag));
    {
        const int one = 1;
        const double *tmp;
        double pad = 0;
        memcpy(&pad, &one, sizeof(int));
        PB_TRY(pb_upload_small(ctx, &pad, 1, &tmp));
        PB_TRY(pb_upload_flush(ctx));
        PB_CUDA(ctx, cudaMemcpyAsync(d_flag, tmp, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (d_cosb != p.cosb_og) {
        if (p.ld != W && !host) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "thermal_sh: padded device arrays need cosb == cosb_og aliasing");
        const int64_t n = (int64_t)B * L * W;
        arrays_equal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, d_cosb, p.cosb_og, d_flag);
        PB_CHECK_LAUNCH(ctx);
    }
    p.ff_zero = d_flag;
    double *d_x = nullptr, *d_th = nullptr;
    if (host) {
        if (need_x) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * nW, (void **)&d_x));
        if (a->thermal) PB_TRY(pb_arena_alloc(ctx, B * nW, (void **)&d_th));
    } else {
        d_x = a->xint_at_top;
        if (!d_x && need_x) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * nW, (void **)&d_x));
        d_th = a->thermal;
    }
    p.xint = d_x; p.thermal = d_th; p.fuse = fuse ? 1 : 0;
    const int ay = G < 4 ? G : 4;
    dim3 block(kWaves, ay, 1);
    dim3 grid((W + kWaves - 1) / kWaves, (G + ay - 1) / ay, B);
    const size_t smem = fuse ? (size_t)ay * kWaves * sizeof(double) : 0;
    if (a->stream == 2) sh_thermal_kernel<2><<<grid, block, smem, ctx->stream>>>(p);
    else sh_thermal_kernel<4><<<grid, block, smem, ctx->stream>>>(p);
    PB_CHECK_LAUNCH(ctx);
    if (a->thermal && !fuse) {
        dim3 g2((W + 127) / 128, B);
        sh_compress_thermal_kernel<<<g2, 128, 0, ctx->stream>>>(W, G, a->numt, d_x, p.gweight, p.tweight, d_th);
        PB_CHECK_LAUNCH(ctx);
    }
    if (host) {
        if (a->xint_at_top) PB_CUDA(ctx, cudaMemcpyAsync(a->xint_at_top, d_x, (size_t)B * G * nW, cudaMemcpyDeviceToHost, ctx->stream));
        if (a->thermal) PB_CUDA(ctx, cudaMemcpyAsync(a->thermal, d_th, B * nW, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
