// toon_thermal.cu - thermal-emission Toon89 two-stream + source-function solver, sm_100a.
//
// Replaces picaso/fluxes.py:1683-1912 (get_thermal_1d; blackbody :1661-1680,
// blackbody_integrated :1609-1658) and, optionally fused, disco.compress_thermal
// (disco.py:152-180).
//
// Same single-sweep idea as toon_reflected.cu: the TOA output flux_plus_mdpt[0] is
// the end of an upward recurrence whose per-layer source terms are linear in the
// tridiagonal solution, so one bottom-up sweep (Thomas elimination in the direction
// of tri_diag_solve + symbolic carry of the upward flux) yields it in registers.
// The Planck function is evaluated in the kernel from tlevel (no [nlevel, nwno]
// blackbody matrix is ever materialised).
#include <cstdlib>

#include "pb_common.cuh"
#include "pb_math.cuh"

namespace {

struct ThermParams {
    int L, W, G, nt;
    int64_t ld, bs_layer, bs_wave;
    const double *dtau, *w0, *cosb;
    const double *wno, *dwno, *surf;
    const double *tlevel, *plevel;  // [B][V] device
    const double *ubar1, *gweight, *tweight;
    int hard_surface, calc_type;
    double *ftop, *thermal;
    double *fm, *fp, *fmm, *fpm;
    int fuse;
    int variant;  // 1: get_thermal_3d per-facet semantics (fluxes.py:2148-2352)
    int wt, ay;   // therm_toa_kernel<GEN = true>, therm_toa_chain_kernel: wavelengths / angles per CTA
    int ch;       // therm_toa_chain_kernel: layers per chunk (= producing warps)
    int ob_period; // > 0: batch entry b reads opacity / surf block b % ob_period (pb_thermal_args.opacity_period)
};

constexpr int kWavesPerCta = 32;
constexpr double kMu1 = 0.5;  // fluxes.py:1748
#ifndef PB_THERM_CHAIN
#define PB_THERM_CHAIN 1
#endif
// PB_THERM_CHAIN_TAB 1: the consumers' exp(-dtau/u) through the 64-entry table of pb_math.cuh (9 fp64 instructions, 8 KB of
// shared memory: chunks of 3 layers instead of 4 keep three CTAs per SM)
#ifndef PB_THERM_CHAIN_TAB
#define PB_THERM_CHAIN_TAB 0
#endif
#ifndef PB_THERM_CHAIN_UNROLL
#define PB_THERM_CHAIN_UNROLL 1
#endif
constexpr int kThermChainUnroll = PB_THERM_CHAIN_UNROLL;
constexpr int kThermChainCh = PB_THERM_CHAIN_TAB ? 3 : 4;
constexpr int kThermChainTab = PB_THERM_CHAIN_TAB ? pbm::kExpTabDoubles : 0;
constexpr bool kThermChainDefault = PB_THERM_CHAIN != 0;  // therm_toa_chain_kernel for angle-parallel launches

struct Planck {
    double c1w, c2w;      // calc_type 0: B = c1w / (exp(c2w / T) - 1)
    double wn, dw;        // calc_type 1
    int type;
    __device__ __forceinline__ void init(int calc_type, double wno, double dwno)
    {
        const double h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
        type = calc_type;
        if (calc_type == 0) {
            // blackbody(t, 1/wno), fluxes.py:1676-1680
            const double wl = 1.0 / wno;
            c1w = (2.0 * h * c * c) / pow(wl, 5.0);
            c2w = (h * c) / (wl * k);
        } else {
            wn = wno;
            dw = dwno;
        }
    }
    __device__ __forceinline__ double operator()(double t) const
    {
        if (type == 0) return c1w * (1.0 / (exp(c2w / t) - 1.0));
        // blackbody_integrated, fluxes.py:1632-1656, nbb = 1
        const double h = 6.62607004e-27, c = 2.99792458e+10, k = 1.38064852e-16;
        const double c1 = 2 * h * c * c, c2 = h * c / k;
        double s = 0.0;
#pragma unroll
        for (int kk = -1; kk <= 1; ++kk) {
            const double wv = wn + kk * dw / 2.0;
            s += c1 * (wv * wv * wv) / (exp(c2 * wv / t) - 1.0);
        }
        return s / 3.0;
    }
};

// per-layer two-stream quantities, fluxes.py:1756-1789
struct TLayer {
    double b0, b1, lam, gam, q, cpu, cmu, cpd, cmd, E, EP, EM, e1, e2, e3, e4;
};

__device__ __forceinline__ void thermal_layer(double dt, double om, double g, double Btop,
                                              double Bbot, TLayer &t)
{
    t.b0 = Btop;
    t.b1 = (Bbot - Btop) / dt;
    const double g1 = 2.0 - om * (1 + g), g2 = om * (1 - g);
    t.lam = sqrt(g1 * g1 - g2 * g2);
    t.gam = (g1 - t.lam) / g2;
    t.q = 1.0 / (g1 + g2);
    const double tp = 2 * PB_PI * kMu1;
    t.cpu = tp * (t.b0 + t.b1 * t.q);
    t.cmu = tp * (t.b0 - t.b1 * t.q);
    t.cpd = tp * (t.b0 + t.b1 * dt + t.b1 * t.q);
    t.cmd = tp * (t.b0 + t.b1 * dt - t.b1 * t.q);
    t.E = fmin(t.lam * dt, 35.0);
    t.EP = exp(t.E);
    t.EM = 1.0 / t.EP;
    t.e1 = t.EP + t.gam * t.EM;
    t.e2 = t.EP - t.gam * t.EM;
    t.e3 = t.gam * t.EP + t.EM;
    t.e4 = t.gam * t.EP - t.EM;
}

// ---------------------------------------------------------------------------------------
// TOA flux.  CTA = 32 wavelengths (lane) x NW angle-warps, same chunked producer/consumer
// scheme as refl_toa_kernel: the Planck function of every level is evaluated once per CTA
// into shared memory, then layers are walked bottom-up in chunks of NW; each warp produces
// the angle-independent two-stream quantities of one layer of the next chunk (sqrt, 4
// reciprocals, exp) and consumes the current chunk for its own angle (1 exp, 3
// reciprocals per layer).
// ---------------------------------------------------------------------------------------
enum { T_LAM = 0, T_GAM, T_EP, T_EM, T_CPU, T_CMU, T_CPD, T_CMD, T_DT, T_AL1, T_AL2, T_E, T_B1, TNQ };

__device__ __forceinline__ void therm_produce(double dt, double om, double g, double Btop, double Bbot,
                                              double *q /* [TNQ][32] column of this lane */)
{
    // fluxes.py:1756-1789, :1846-1847
    const double b0 = Btop;
    const double b1 = (Bbot - Btop) * pbm::krcp(dt);
    const double g1 = 2.0 - om * (1 + g), g2 = om * (1 - g);
    const double lam = sqrt(g1 * g1 - g2 * g2);
    const double gam = (g1 - lam) * pbm::krcp(g2);
    const double qq = pbm::krcp(g1 + g2);
    const double tp = 2 * PB_PI * kMu1;
    const double E = fmin(lam * dt, 35.0);
    const double EP = pbm::kexp(E);
    q[T_LAM * 32] = lam;
    q[T_GAM * 32] = gam;
    q[T_EP * 32] = EP;
    q[T_EM * 32] = pbm::krcp(EP);
    q[T_CPU * 32] = tp * (b0 + b1 * qq);
    q[T_CMU * 32] = tp * (b0 - b1 * qq);
    q[T_CPD * 32] = tp * (b0 + b1 * dt + b1 * qq);
    q[T_CMD * 32] = tp * (b0 + b1 * dt - b1 * qq);
    q[T_DT * 32] = dt;
    q[T_AL1 * 32] = 2 * PB_PI * (b0 + b1 * (qq - kMu1));
    q[T_AL2 * 32] = 2 * PB_PI * b1;
    q[T_E * 32] = E;
    q[T_B1 * 32] = b1;
}

// GEN: wt (< 32) wavelengths x ay angles flattened over a 1-D block (thread t -> wave t % wt, angle t / wt)
// so that the grid can be sized to the SM count - see refl_toa_kernel4 in toon_reflected_toa4.cuh.
template <bool GEN>
__global__ void __launch_bounds__(256) therm_toa_kernel(ThermParams p)
{
    extern __shared__ double smem[];  // [V][32] Planck | [2][NW][TNQ][32] tiles
    const int tid = GEN ? (int)threadIdx.x : (int)(threadIdx.y * 32 + threadIdx.x);
    const int lane = tid & 31, wy = tid >> 5;  // producer identity
    const int NW = GEN ? (int)(blockDim.x >> 5) : (int)blockDim.y;
    const int WT = GEN ? p.wt : kWavesPerCta, AY = GEN ? p.ay : NW;
    const int cw = GEN ? tid % WT : lane, ca = GEN ? tid / WT : wy;  // consumer identity
    const int w = blockIdx.x * WT + cw;
    const int wp = blockIdx.x * WT + lane;
    const int wc = wp < p.W ? wp : p.W - 1;          // producer column (clamped)
    const int wcc = w < p.W ? w : p.W - 1;           // consumer column (clamped)
    const int a = blockIdx.y * AY + ca;
    const int ac = a < p.G ? a : p.G - 1;
    const int b = blockIdx.z;
    const int L = p.L, V = p.L + 1;
    const int64_t ld = p.ld;
    const int64_t ol = (int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_layer + wc;
    const double *tl = p.tlevel + (int64_t)b * V;
    const double *pl = p.plevel + (int64_t)b * V;
    const double u = p.variant ? p.ubar1[b] : p.ubar1[ac];
    const double inv_u = 1.0 / u;
    const double r = p.surf ? p.surf[(int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_wave + wcc] : 0.0;
    double *sB = smem;
    double *tiles = smem + (size_t)V * 32;
    const int tile = NW * TNQ * 32;
    const int nchunks = (L + NW - 1) / NW;

    {
        Planck planck;
        planck.init(p.calc_type, p.wno[wc], p.dwno ? p.dwno[wc] : 0.0);
        for (int v = wy; v < V; v += NW) sB[v * 32 + lane] = planck(tl[v]);
    }
    __syncthreads();
    const double BL = sB[L * 32 + cw], B0 = sB[cw];
    {
        const int l = L - 1 - wy;
        if (l >= 0) {
            const int64_t il = ol + (int64_t)l * ld;
            therm_produce(__ldg(p.dtau + il), __ldg(p.w0 + il), __ldg(p.cosb + il), sB[l * 32 + lane],
                          sB[(l + 1) * 32 + lane], tiles + wy * TNQ * 32 + lane);
        }
    }
    __syncthreads();
    double AS = 0.0, DS = 0.0, Pp = 0.0, Rp = 0.0;
    double gam_n = 0.0, cpu_n = 0.0, cmu_n = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        const int ln = L - 1 - ((c + 1) * NW + wy);
        const bool have_next = (c + 1 < nchunks) && (ln >= 0);
        double ndt = 0.0, nom = 0.0, ncb = 0.0;
        if (have_next) {
            const int64_t il = ol + (int64_t)ln * ld;
            ndt = __ldg(p.dtau + il);
            nom = __ldg(p.w0 + il);
            ncb = __ldg(p.cosb + il);
        }
        const double *buf = tiles + (c & 1) * tile + cw;
        const int lbase = L - 1 - c * NW;
        const int nk = lbase + 1 < NW ? lbase + 1 : NW;
        for (int k = 0; k < nk; ++k) {
            const int l = lbase - k;
            const double *q = buf + k * TNQ * 32;
            const double lam = q[T_LAM * 32], gam = q[T_GAM * 32], EP = q[T_EP * 32], EM = q[T_EM * 32];
            const double cpu = q[T_CPU * 32], cmu = q[T_CMU * 32], cpd = q[T_CPD * 32], cmd = q[T_CMD * 32];
            const double dt = q[T_DT * 32], al1 = q[T_AL1 * 32], al2 = q[T_AL2 * 32];
            const double e1 = EP + gam * EM, e2 = EP - gam * EM;
            const double e3 = gam * EP + EM, e4 = gam * EP - EM;
            // Table 3 of Toon89 (fluxes.py:1842-1849): G = (1/mu1 - lam) Y+, H = gam (lam + 1/mu1) Y-
            const double lu = lam * u;
            const double inv_l = pbm::krcp(lu * lu - 1.0);
            const double kG = (1 / kMu1 - lam) * ((lu + 1.0) * inv_l);
            const double kH = gam * (lam + 1 / kMu1) * ((lu - 1.0) * inv_l);
            double x, cG, cH, K;
            if (l > 0) {
                // flux_plus recurrence, fluxes.py:1897-1901
                x = pbm::kexp(-dt * inv_u);
                cG = kG * (EP * x - 1.0);
                cH = kH * (1.0 - EM * x);
                K = al1 * (1. - x) + al2 * (u - (dt + u) * x);
            } else {
                // flux_plus_mdpt[0], fluxes.py:1903-1910
                x = pbm::kexp(-0.5 * dt * inv_u);
                const double EPh = pbm::kexp(0.5 * q[T_E * 32]), EMh = pbm::krcp(EPh);
                cG = kG * (EP * x - EPh);
                cH = -kH * (EM * x - EMh);
                K = al1 * (1. - x) + al2 * (u + 0.5 * dt - (dt + u) * x);
            }
            double alpha, beta;
            if (l == L - 1) {
                // surface boundary, fluxes.py:1802-1806, :1869-1873, last row :178-181
                const double b1 = q[T_B1 * 32];
                // 3-D variant: b_surface = pi B_L without emissivity, int_plus[L] = pi b_surface | pi (B_L + b1 u)
                const double b_surface = p.hard_surface ? (p.variant ? PB_PI * BL : (1.0 - r) * BL * PB_PI)
                                                        : (BL + b1 * kMu1) * PB_PI;
                const double a_ = e1 - r * e3, b_ = e2 - r * e4;
                const double d_ = b_surface - cpd + r * cmd;
                const double ib = pbm::krcp(b_);
                AS = a_ * ib;
                DS = d_ * ib;
                if (p.variant) alpha = p.hard_surface ? PB_PI * b_surface : PB_PI * (BL + b1 * u);
                else alpha = p.hard_surface ? (1.0 - r) * BL * 2 * PB_PI : (BL + b1 * u) * 2 * PB_PI;
                beta = 0.0;
            } else {
                const double gm1 = gam_n - 1.0;
                const double e13 = (e1 + e3) * gm1;
                double a_ = 2.0 * (1.0 - gam * gam);
                double b_ = (e1 - e3) * (gam_n + 1.0);
                double d_ = e3 * (cpu_n - cpd) + e1 * (cmd - cmu_n);
                double xi = pbm::krcp(b_ - e13 * AS);
                const double ASe = a_ * xi, DSe = (d_ - e13 * DS) * xi;
                alpha = Rp + Pp * DSe;
                beta = -Pp * ASe;
                b_ = (e2 + e4) * gm1;
                const double c_ = 2.0 * (1.0 - gam_n * gam_n);
                d_ = gm1 * (cpu_n - cpd) - gm1 * (cmd - cmu_n);
                xi = pbm::krcp(b_ - c_ * ASe);
                AS = e13 * xi;
                DS = (d_ - c_ * DSe) * xi;
            }
            // F+_l = x (alpha + beta X[2l+1]) + cG (X0 + X1) + cH (X0 - X1) + K
            const double P = cG + cH;
            const double Q = x * beta + (cG - cH);
            const double R = x * alpha + K;
            Pp = P - Q * AS;
            Rp = R + Q * DS;
            gam_n = gam;
            cpu_n = cpu;
            cmu_n = cmu;
        }
        if (have_next)
            therm_produce(ndt, nom, ncb, sB[ln * 32 + lane], sB[(ln + 1) * 32 + lane],
                          tiles + ((c + 1) & 1) * tile + wy * TNQ * 32 + lane);
        __syncthreads();
    }
    double result;
    {
        // top boundary: fake isothermal overburden, fluxes.py:1797-1800; row 0 :155-158
        const double tau_top = __ldg(p.dtau + (int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_layer + wcc) * pl[0] / (pl[1] - pl[0]);
        const double b_top = (1.0 - exp(-tau_top / kMu1)) * B0 * PB_PI;
        const double b_ = gam_n + 1.0, c_ = gam_n - 1.0, d_ = b_top - cmu_n;
        const double xi = pbm::krcp(b_ - c_ * AS);
        const double X0 = (d_ - c_ * DS) * xi;
        result = Rp + Pp * X0;
    }
    const bool active = (w < p.W) && (a < p.G) && (ca < AY);
    if (active && p.ftop) p.ftop[((int64_t)b * p.G + a) * p.W + w] = result;
    if (p.fuse) {
        double *s_f = tiles;
        if (ca < AY) s_f[ca * kWavesPerCta + cw] = result;
        __syncthreads();
        if (ca == 0 && w < p.W) {
            double acc = 0.0;
            for (int aa = 0; aa < p.G; ++aa) {
                const int ig = aa / p.nt, it = aa - ig * p.nt;
                acc = acc + s_f[aa * kWavesPerCta + cw] * p.gweight[ig] * p.tweight[it];
            }
            const double sym = (p.nt == 1) ? 1.0 : 1 / (2 * PB_PI);
            p.thermal[(int64_t)b * p.W + w] = acc * sym;
        }
    }
}

// ---------------------------------------------------------------------------------------
// TOA flux with the elimination taken out of the angle threads (round 2; the recipe of refl_toa_kernel5).
//
// In get_thermal_1d the tridiagonal system - matrix AND right-hand side - is the same for every viewing angle
// (fluxes.py:1812-1831); therm_toa_kernel repeats its bottom-up elimination (two reciprocals and ~30 fp64
// instructions per layer) in all G angle threads of a wavelength.  Here one extra warp per CTA, the CHAIN WARP
// (lane = wavelength), walks the elimination once, one chunk ahead of the consumers, and publishes four numbers per
// layer (AS, DS of the odd and of the even interface row); an angle thread keeps the angle-dependent Table-3 terms
// and the two affine updates of its functional  flux_at_top = Rp + Pp X[2l].  The chain warp also closes the system
// (top boundary row -> X[0]).  Pipeline per chunk of CH layers, one barrier: consumers integrate chunk c
// (P[c % 3], C[c & 1]), the chain warp eliminates chunk c + 1 (P[(c+1) % 3] -> C[(c+1) & 1]), the producing warps
// write chunk c + 2 (P[(c+2) % 3]) from inputs loaded one iteration earlier.  Same expressions as therm_toa_kernel.
// ---------------------------------------------------------------------------------------
enum { TC_ASE = 0, TC_DSE, TC_AS, TC_DS, TNC };

__global__ void __maxnreg__(112) therm_toa_chain_kernel(ThermParams p)
{
    extern __shared__ double smem[];  // Planck [V][32] | P tiles [3][CH][TNQ][32] | C tiles [2][CH][TNC][32] | X0 [32]
    const int tid = (int)threadIdx.x;
    const int lane = tid & 31, wy = tid >> 5;
    const int NW = (int)(blockDim.x >> 5), NWC = NW - 1;  // the last warp is the chain warp
    const int CH = p.ch, WT = p.wt, AY = p.ay;
    const bool is_chain = wy == NWC;
    const bool is_cons = tid < WT * AY;
    const int cw = is_cons ? tid % WT : 0, ca = is_cons ? tid / WT : 0;  // consumer identity
    const int w = blockIdx.x * WT + cw;
    const int wp = blockIdx.x * WT + (lane < WT ? lane : WT - 1);        // producer / chain column
    const int wc = wp < p.W ? wp : p.W - 1;
    const int wcc = w < p.W ? w : p.W - 1;
    const int a = blockIdx.y * AY + ca;
    const int ac = a < p.G ? a : p.G - 1;
    const int b = blockIdx.z;
    const int L = p.L, V = p.L + 1;
    const int64_t ld = p.ld;
    const int ob = p.ob_period ? b % p.ob_period : b;
    const int64_t ol = (int64_t)ob * p.bs_layer + wc;
    const double *tl = p.tlevel + (int64_t)b * V;
    const double *pl = p.plevel + (int64_t)b * V;
    const double u = p.ubar1[ac];
    const double inv_u = 1.0 / u;
    double *sB = smem + kThermChainTab;
    double *ptile = sB + (size_t)V * 32;
    const int psz = CH * TNQ * 32, csz = CH * TNC * 32;
    double *ctile = ptile + 3 * psz;
    double *sX0 = ctile + 2 * csz;
    const int nchunks = (L + CH - 1) / CH;
#if PB_THERM_CHAIN_TAB
    pbm::exp_tab_fill(smem, tid, (int)blockDim.x);
    const double *tab = smem + (lane & 15);
#define PB_TC_EXP(x) pbm::exp_tab((x), tab)
#else
#define PB_TC_EXP(x) pbm::kexp(x)
#endif
    {
        Planck planck;
        planck.init(p.calc_type, p.wno[wc], p.dwno ? p.dwno[wc] : 0.0);
        for (int v = wy; v < V; v += NW) sB[v * 32 + lane] = planck(tl[v]);
    }
    __syncthreads();

    // producers: warp wy < CH owns layer L - 1 - (c CH + wy) of chunk c
    auto load_in = [&](int c, double &dt, double &om, double &cb) -> bool {
        const int l = L - 1 - (c * CH + wy);
        if (wy >= CH || c >= nchunks || l < 0) return false;
        const int64_t il = ol + (int64_t)l * ld;
        dt = __ldg(p.dtau + il);
        om = __ldg(p.w0 + il);
        cb = __ldg(p.cosb + il);
        return true;
    };
    auto produce = [&](int c, double dt, double om, double cb) {
        const int l = L - 1 - (c * CH + wy);
        therm_produce(dt, om, cb, sB[l * 32 + lane], sB[(l + 1) * 32 + lane], ptile + (c % 3) * psz + wy * TNQ * 32 + lane);
    };

    // chain-warp state (fluxes.py:289-323 bottom-up: X[n] = DS - AS X[n-1])
    double cAS = 0.0, cDS = 0.0, c_gam = 0.0, c_cpu = 0.0, c_cmu = 0.0;
    auto chain = [&](int c) {
        const double *pt = ptile + (c % 3) * psz + lane;
        double *ct = ctile + (c & 1) * csz + lane;
        const int lbase = L - 1 - c * CH;
        const int nk = lbase + 1 < CH ? lbase + 1 : CH;
        for (int k = 0; k < nk; ++k) {
            const int l = lbase - k;
            const double *q = pt + k * TNQ * 32;
            double *o = ct + k * TNC * 32;
            const double gam = q[T_GAM * 32], EP = q[T_EP * 32], EM = q[T_EM * 32];
            const double cpu = q[T_CPU * 32], cmu = q[T_CMU * 32], cpd = q[T_CPD * 32], cmd = q[T_CMD * 32];
            const double e1 = EP + gam * EM, e2 = EP - gam * EM;
            const double e3 = gam * EP + EM, e4 = gam * EP - EM;
            if (l == L - 1) {
                // surface boundary, fluxes.py:1802-1806, last row :178-181
                const double b1 = q[T_B1 * 32];
                const double BL = sB[L * 32 + lane];
                const double r = p.surf ? p.surf[(int64_t)ob * p.bs_wave + wc] : 0.0;
                const double b_surface = p.hard_surface ? (1.0 - r) * BL * PB_PI : (BL + b1 * kMu1) * PB_PI;
                const double a_ = e1 - r * e3, b_ = e2 - r * e4;
                const double d_ = b_surface - cpd + r * cmd;
                const double ib = pbm::krcp(b_);
                cAS = a_ * ib;
                cDS = d_ * ib;
                o[TC_ASE * 32] = 0.0;
                o[TC_DSE * 32] = 0.0;
            } else {
                const double gm1 = c_gam - 1.0;
                const double e13 = (e1 + e3) * gm1;
                double a_ = 2.0 * (1.0 - gam * gam);
                double b_ = (e1 - e3) * (c_gam + 1.0);
                double d_ = e3 * (c_cpu - cpd) + e1 * (cmd - c_cmu);
                double xi = pbm::krcp(b_ - e13 * cAS);
                const double ASe = a_ * xi, DSe = (d_ - e13 * cDS) * xi;
                o[TC_ASE * 32] = ASe;
                o[TC_DSE * 32] = DSe;
                b_ = (e2 + e4) * gm1;
                const double c_ = 2.0 * (1.0 - c_gam * c_gam);
                d_ = gm1 * (c_cpu - cpd) - gm1 * (cmd - c_cmu);
                xi = pbm::krcp(b_ - c_ * ASe);
                cAS = e13 * xi;
                cDS = (d_ - c_ * DSe) * xi;
            }
            o[TC_AS * 32] = cAS;
            o[TC_DS * 32] = cDS;
            c_gam = gam;
            c_cpu = cpu;
            c_cmu = cmu;
        }
    };

    // consumer state: flux_at_top = Rp + Pp X[2l] over the first not-yet-eliminated unknown
    double Pp = 0.0, Rp = 0.0;
    const double r_c = p.surf ? p.surf[(int64_t)ob * p.bs_wave + wcc] : 0.0;
    const double BLc = sB[L * 32 + cw];
    auto consume = [&](int c) {
        const double *pt = ptile + (c % 3) * psz + cw;
        const double *ct = ctile + (c & 1) * csz + cw;
        const int lbase = L - 1 - c * CH;
        const int nk = lbase + 1 < CH ? lbase + 1 : CH;
#pragma unroll kThermChainUnroll
        for (int k = 0; k < nk; ++k) {
            const int l = lbase - k;
            const double *q = pt + k * TNQ * 32;
            const double *o = ct + k * TNC * 32;
            const double lam = q[T_LAM * 32], gam = q[T_GAM * 32], EP = q[T_EP * 32], EM = q[T_EM * 32];
            const double dt = q[T_DT * 32], al1 = q[T_AL1 * 32], al2 = q[T_AL2 * 32];
            // Table 3 of Toon89 (fluxes.py:1842-1849): G = (1/mu1 - lam) Y+, H = gam (lam + 1/mu1) Y-
            const double lu = lam * u;
            const double inv_l = pbm::krcp(lu * lu - 1.0);
            const double kG = (1 / kMu1 - lam) * ((lu + 1.0) * inv_l);
            const double kH = gam * (lam + 1 / kMu1) * ((lu - 1.0) * inv_l);
            double x, cG, cH, K;
            if (l > 0) {
                // flux_plus recurrence, fluxes.py:1897-1901
                x = PB_TC_EXP(-dt * inv_u);
                cG = kG * (EP * x - 1.0);
                cH = kH * (1.0 - EM * x);
                K = al1 * (1. - x) + al2 * (u - (dt + u) * x);
            } else {
                // flux_plus_mdpt[0], fluxes.py:1903-1910
                x = PB_TC_EXP(-0.5 * dt * inv_u);
                const double EPh = PB_TC_EXP(0.5 * q[T_E * 32]), EMh = pbm::krcp(EPh);
                cG = kG * (EP * x - EPh);
                cH = -kH * (EM * x - EMh);
                K = al1 * (1. - x) + al2 * (u + 0.5 * dt - (dt + u) * x);
            }
            double alpha, beta;
            if (l == L - 1) {
                // fluxes.py:1869-1873
                const double b1 = q[T_B1 * 32];
                alpha = p.hard_surface ? (1.0 - r_c) * BLc * 2 * PB_PI : (BLc + b1 * u) * 2 * PB_PI;
                beta = 0.0;
            } else {
                alpha = Rp + Pp * o[TC_DSE * 32];
                beta = -Pp * o[TC_ASE * 32];
            }
            // F+_l = x (alpha + beta X[2l+1]) + cG (X0 + X1) + cH (X0 - X1) + K
            const double P = cG + cH;
            const double Q = x * beta + (cG - cH);
            const double R = x * alpha + K;
            Pp = P - Q * o[TC_AS * 32];
            Rp = R + Q * o[TC_DS * 32];
        }
    };

    double ndt = 0.0, nom = 0.0, ncb = 0.0;
    if (load_in(0, ndt, nom, ncb)) produce(0, ndt, nom, ncb);
    bool have = load_in(1, ndt, nom, ncb);
    __syncthreads();
    if (is_chain) chain(0);
    if (have) produce(1, ndt, nom, ncb);
    have = load_in(2, ndt, nom, ncb);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        if (is_chain) {
            if (c + 1 < nchunks) chain(c + 1);
        } else if (is_cons) {
            consume(c);
        }
        if (have) produce(c + 2, ndt, nom, ncb);
        have = load_in(c + 3, ndt, nom, ncb);
        __syncthreads();
    }
    if (is_chain) {
        // top boundary: fake isothermal overburden, fluxes.py:1797-1800; row 0 :155-158
        const double tau_top = __ldg(p.dtau + (int64_t)ob * p.bs_layer + wc) * pl[0] / (pl[1] - pl[0]);
        const double b_top = (1.0 - exp(-tau_top / kMu1)) * sB[lane] * PB_PI;
        const double b_ = c_gam + 1.0, c_ = c_gam - 1.0, d_ = b_top - c_cmu;
        const double xi = pbm::krcp(b_ - c_ * cAS);
        sX0[lane] = (d_ - c_ * cDS) * xi;
    }
    __syncthreads();
    const double result = Rp + Pp * sX0[cw];
    const bool active = is_cons && (w < p.W) && (a < p.G);
    if (active && p.ftop) p.ftop[((int64_t)b * p.G + a) * p.W + w] = result;
    if (p.fuse) {
        double *s_f = ptile;
        if (is_cons) s_f[ca * kWavesPerCta + cw] = result;
        __syncthreads();
        if (is_cons && ca == 0 && w < p.W) {
            double acc = 0.0;
            for (int aa = 0; aa < p.G; ++aa) {
                const int ig = aa / p.nt, it = aa - ig * p.nt;
                acc = acc + s_f[aa * kWavesPerCta + cw] * p.gweight[ig] * p.tweight[it];
            }
            const double sym = (p.nt == 1) ? 1.0 : 1 / (2 * PB_PI);
            p.thermal[(int64_t)b * p.W + w] = acc * sym;
        }
    }
}

// ---------------------------------------------------------------------------------------
// TOA flux, one thread per (atmosphere, wavelength) - used when the launch has enough wavelengths
// to fill the machine without the angle axis (batched retrievals, R ~ 1e5 grids).
//
// In get_thermal_1d the tridiagonal system - matrix AND right-hand side - does not depend on the
// viewing angle (fluxes.py:1812-1831: one solve per wavelength); only the upward source-function
// recurrence does (:1864-1907).  therm_toa_kernel nevertheless repeats the elimination in every
// angle warp (it needs the angle axis for parallelism at W = 10 000).  Here a thread eliminates
// once per layer and carries the G linear functionals (Rp_a + Pp_a X[2l]) of all angles in
// registers: per (layer, wavelength) the two pivots / reciprocals, sqrt, exp(lam dtau) and Planck
// are paid once instead of G times, nothing goes through shared memory and there is no barrier.
// ---------------------------------------------------------------------------------------
#ifndef PB_THERM_WAVE_MINB
#define PB_THERM_WAVE_MINB 4
#endif
template <int G>
__global__ void __launch_bounds__(128, PB_THERM_WAVE_MINB) therm_toa_wave_kernel(ThermParams p)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (w >= p.W) return;
    const int L = p.L, V = p.L + 1;
    const int64_t ld = p.ld;
    const int64_t ol = (int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_layer + w;
    const double *tl = p.tlevel + (int64_t)b * V;
    const double *pl = p.plevel + (int64_t)b * V;
    const double r = p.surf ? p.surf[(int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_wave + w] : 0.0;
    double u[G], inv_u[G], Pp[G], Rp[G];
#pragma unroll
    for (int a = 0; a < G; ++a) {
        u[a] = p.ubar1[a];
        inv_u[a] = 1.0 / u[a];
        Pp[a] = 0.0;
        Rp[a] = 0.0;
    }
    Planck planck;
    planck.init(p.calc_type, p.wno[w], p.dwno ? p.dwno[w] : 0.0);
    const double BL = planck(tl[L]);
    double Bbot = BL;
    double AS = 0.0, DS = 0.0, gam_n = 0.0, cpu_n = 0.0, cmu_n = 0.0;
    const double tp = 2 * PB_PI * kMu1;
    // software prefetch of the next layer's three inputs
    double ndt = __ldg(p.dtau + ol + (int64_t)(L - 1) * ld), nom = __ldg(p.w0 + ol + (int64_t)(L - 1) * ld),
           ncb = __ldg(p.cosb + ol + (int64_t)(L - 1) * ld);
    for (int l = L - 1; l >= 0; --l) {
        const double dt = ndt, om = nom, g = ncb;
        if (l > 0) {
            const int64_t il = ol + (int64_t)(l - 1) * ld;
            ndt = __ldg(p.dtau + il);
            nom = __ldg(p.w0 + il);
            ncb = __ldg(p.cosb + il);
        }
        const double Btop = planck(tl[l]);
        // fluxes.py:1756-1789, :1846-1847 (same arithmetic as therm_produce)
        const double b0 = Btop;
        const double b1 = (Bbot - Btop) * pbm::krcp(dt);
        const double g1 = 2.0 - om * (1 + g), g2 = om * (1 - g);
        const double lam = sqrt(g1 * g1 - g2 * g2);
        const double gam = (g1 - lam) * pbm::krcp(g2);
        const double qq = pbm::krcp(g1 + g2);
        const double E = fmin(lam * dt, 35.0);
        const double EP = pbm::kexp(E), EM = pbm::krcp(EP);
        const double cpu = tp * (b0 + b1 * qq), cmu = tp * (b0 - b1 * qq);
        const double cpd = tp * (b0 + b1 * dt + b1 * qq), cmd = tp * (b0 + b1 * dt - b1 * qq);
        const double al1 = 2 * PB_PI * (b0 + b1 * (qq - kMu1)), al2 = 2 * PB_PI * b1;
        const double e1 = EP + gam * EM, e2 = EP - gam * EM;
        const double e3 = gam * EP + EM, e4 = gam * EP - EM;
        // ---- angle-independent elimination step ----
        double ASe = 0.0, DSe = 0.0;
        const bool bottom = (l == L - 1);
        double b_surface = 0.0;
        if (bottom) {
            // surface boundary, fluxes.py:1802-1806, last row :178-181
            b_surface = p.hard_surface ? (1.0 - r) * BL * PB_PI : (BL + b1 * kMu1) * PB_PI;
            const double a_ = e1 - r * e3, b_ = e2 - r * e4;
            const double d_ = b_surface - cpd + r * cmd;
            const double ib = pbm::krcp(b_);
            AS = a_ * ib;
            DS = d_ * ib;
        } else {
            const double gm1 = gam_n - 1.0;
            const double e13 = (e1 + e3) * gm1;
            double a_ = 2.0 * (1.0 - gam * gam);
            double b_ = (e1 - e3) * (gam_n + 1.0);
            double d_ = e3 * (cpu_n - cpd) + e1 * (cmd - cmu_n);
            double xi = pbm::krcp(b_ - e13 * AS);
            ASe = a_ * xi;
            DSe = (d_ - e13 * DS) * xi;
            b_ = (e2 + e4) * gm1;
            const double c_ = 2.0 * (1.0 - gam_n * gam_n);
            d_ = gm1 * (cpu_n - cpd) - gm1 * (cmd - cmu_n);
            xi = pbm::krcp(b_ - c_ * ASe);
            AS = e13 * xi;
            DS = (d_ - c_ * DSe) * xi;
        }
        double EPh = 0.0, EMh = 0.0;
        if (l == 0) {
            EPh = pbm::kexp(0.5 * E);
            EMh = pbm::krcp(EPh);
        }
        // ---- per angle: Table 3 of Toon89 (fluxes.py:1842-1849) and the flux_plus recurrence (:1897-1910) ----
#pragma unroll
        for (int a = 0; a < G; ++a) {
            const double lu = lam * u[a];
            const double inv_l = pbm::krcp(lu * lu - 1.0);
            const double kG = (1 / kMu1 - lam) * ((lu + 1.0) * inv_l);
            const double kH = gam * (lam + 1 / kMu1) * ((lu - 1.0) * inv_l);
            double x, cG, cH, K;
            if (l > 0) {
                x = pbm::kexp(-dt * inv_u[a]);
                cG = kG * (EP * x - 1.0);
                cH = kH * (1.0 - EM * x);
                K = al1 * (1. - x) + al2 * (u[a] - (dt + u[a]) * x);
            } else {
                x = pbm::kexp(-0.5 * dt * inv_u[a]);
                cG = kG * (EP * x - EPh);
                cH = -kH * (EM * x - EMh);
                K = al1 * (1. - x) + al2 * (u[a] + 0.5 * dt - (dt + u[a]) * x);
            }
            double alpha, beta;
            if (bottom) {
                alpha = p.hard_surface ? (1.0 - r) * BL * 2 * PB_PI : (BL + b1 * u[a]) * 2 * PB_PI;
                beta = 0.0;
            } else {
                alpha = Rp[a] + Pp[a] * DSe;
                beta = -Pp[a] * ASe;
            }
            const double P = cG + cH;
            const double Q = x * beta + (cG - cH);
            const double R = x * alpha + K;
            Pp[a] = P - Q * AS;
            Rp[a] = R + Q * DS;
        }
        gam_n = gam;
        cpu_n = cpu;
        cmu_n = cmu;
        Bbot = Btop;
    }
    // top boundary: fake isothermal overburden, fluxes.py:1797-1800; row 0 :155-158
    const double B0 = Bbot;
    const double tau_top = __ldg(p.dtau + ol) * pl[0] / (pl[1] - pl[0]);
    const double b_top = (1.0 - exp(-tau_top / kMu1)) * B0 * PB_PI;
    const double bb = gam_n + 1.0, cc = gam_n - 1.0, dd = b_top - cmu_n;
    const double xi = pbm::krcp(bb - cc * AS);
    const double X0 = (dd - cc * DS) * xi;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < G; ++a) {
        const double result = Rp[a] + Pp[a] * X0;
        if (p.ftop) p.ftop[((int64_t)b * G + a) * p.W + w] = result;
        if (p.fuse) {
            const int ig = a / p.nt, it = a - ig * p.nt;
            acc = acc + result * p.gweight[ig] * p.tweight[it];
        }
    }
    if (p.fuse) {
        const double sym = (p.nt == 1) ? 1.0 : 1 / (2 * PB_PI);
        p.thermal[(int64_t)b * p.W + w] = acc * sym;
    }
}

template <int G>
static void launch_therm_wave(const ThermParams &p, int B, cudaStream_t st)
{
    dim3 grid((p.W + 127) / 128, B);
    therm_toa_wave_kernel<G><<<grid, 128, 0, st>>>(p);
}

// ---------------------------------------------------------------------------------------
// All four level arrays (fluxes.py:1864-1907).  Three sweeps per thread, with the
// caller's output arrays doubling as O(L) scratch:
//   1. bottom-up elimination, (AS,DS) of rows 2l | 2l+1 parked in (fm,fp) | (fmm,fpm)
//   2. top-down substitution -> Y+/-; downward recurrences written to fm/fmm;
//      Y+/- parked in fp/fpm
//   3. bottom-up: upward recurrences overwrite fp/fpm
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) therm_levels_kernel(ThermParams p)
{
    const int lane = threadIdx.x;
    const int w = blockIdx.x * kWavesPerCta + lane;
    const int a = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W || a >= p.G) return;
    const int L = p.L, V = p.L + 1;
    const int64_t ld = p.ld;
    const int64_t ol = (int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_layer + w;
    const int64_t oo = (((int64_t)b * p.G + a) * V) * p.W + w;
    const double *tl = p.tlevel + (int64_t)b * V;
    const double *pl = p.plevel + (int64_t)b * V;
    const double u = p.ubar1[a];
    const double r = p.surf ? p.surf[(int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_wave + w] : 0.0;
    Planck planck;
    planck.init(p.calc_type, p.wno[w], p.dwno ? p.dwno[w] : 0.0);

    // ---- sweep 1 ----
    double AS = 0.0, DS = 0.0, gam_n = 0.0, cpu_n = 0.0, cmu_n = 0.0;
    double Bbot = planck(tl[L]);
    const double BL = Bbot;
    double b1_last = 0.0;
    for (int l = L - 1; l >= 0; --l) {
        const int64_t il = ol + (int64_t)l * ld;
        const double Btop = planck(tl[l]);
        TLayer t;
        thermal_layer(p.dtau[il], p.w0[il], p.cosb[il], Btop, Bbot, t);
        if (l == L - 1) {
            b1_last = t.b1;
            const double b_surface = p.hard_surface ? (1.0 - r) * BL * PB_PI
                                                    : (BL + t.b1 * kMu1) * PB_PI;
            const double a_ = t.e1 - r * t.e3, b_ = t.e2 - r * t.e4;
            const double d_ = b_surface - t.cpd + r * t.cmd;
            AS = a_ / b_;
            DS = d_ / b_;
        } else {
            double a_ = 2.0 * (1.0 - t.gam * t.gam);
            double b_ = (t.e1 - t.e3) * (gam_n + 1.0);
            double c_ = (t.e1 + t.e3) * (gam_n - 1.0);
            double d_ = t.e3 * (cpu_n - t.cpd) + t.e1 * (t.cmd - cmu_n);
            double xi = 1.0 / (b_ - c_ * AS);
            const double ASe = a_ * xi, DSe = (d_ - c_ * DS) * xi;
            const int64_t o1 = oo + (int64_t)(l + 1) * p.W;
            p.fm[o1] = ASe;
            p.fp[o1] = DSe;
            a_ = (t.e1 + t.e3) * (gam_n - 1.0);
            b_ = (t.e2 + t.e4) * (gam_n - 1.0);
            c_ = 2.0 * (1.0 - gam_n * gam_n);
            d_ = (gam_n - 1.0) * (cpu_n - t.cpd) + (1.0 - gam_n) * (t.cmd - cmu_n);
            xi = 1.0 / (b_ - c_ * ASe);
            AS = a_ * xi;
            DS = (d_ - c_ * DSe) * xi;
        }
        const int64_t o0 = oo + (int64_t)l * p.W;
        p.fmm[o0] = AS;
        p.fpm[o0] = DS;
        gam_n = t.gam;
        cpu_n = t.cpu;
        cmu_n = t.cmu;
        Bbot = Btop;
    }
    const double B0 = Bbot;
    const double tau_top = p.dtau[ol] * pl[0] / (pl[1] - pl[0]);
    {
        const double b_top = (1.0 - exp(-tau_top / kMu1)) * B0 * PB_PI;
        const double b_ = gam_n + 1.0, c_ = gam_n - 1.0, d_ = b_top - cmu_n;
        const double xi = 1.0 / (b_ - c_ * AS);
        p.fm[oo] = 0.0;
        p.fp[oo] = (d_ - c_ * DS) * xi;
    }
    // ---- sweep 2: top-down ----
    double Xprev = 0.0;
    double fminus = (1 - exp(-tau_top / u)) * B0 * 2 * PB_PI;  // fluxes.py:1875
    double Btop = B0;
    for (int l = 0; l < L; ++l) {
        const int64_t il = ol + (int64_t)l * ld;
        const int64_t o0 = oo + (int64_t)l * p.W;
        const double X0 = p.fp[o0] - p.fm[o0] * Xprev;
        const double X1 = p.fpm[o0] - p.fmm[o0] * X0;
        Xprev = X1;
        const double pos = X0 + X1, neg = X0 - X1;
        const double dt = p.dtau[il];
        const double Bb = planck(tl[l + 1]);
        TLayer t;
        thermal_layer(dt, p.w0[il], p.cosb[il], Btop, Bb, t);
        const double J = t.gam * (t.lam + 1 / kMu1) * pos;
        const double K = (1 / kMu1 - t.lam) * neg;
        const double si1 = 2 * PB_PI * (t.b0 - t.b1 * (t.q - kMu1));
        const double si2 = 2 * PB_PI * t.b1;
        const double xa = exp(-dt / u), xh = exp(-0.5 * dt / u);
        const double EPh = exp(0.5 * t.E), EMh = 1 / EPh;
        const double lu = t.lam * u;
        // fluxes.py:1883-1893
        const double fnext = fminus * xa + (J / (lu + 1.0)) * (t.EP - xa) +
                             (K / (lu - 1.0)) * (xa - t.EM) + si1 * (1. - xa) +
                             si2 * (u * xa + dt - u);
        const double fmid = fminus * xh + (J / (lu + 1.0)) * (EPh - xh) +
                            (K / (-lu + 1.0)) * (EMh - xh) + si1 * (1. - xh) +
                            si2 * (u * xh + 0.5 * dt - u);
        p.fm[o0] = fminus;
        p.fmm[o0] = fmid;
        p.fp[o0] = pos;   // parked for sweep 3
        p.fpm[o0] = neg;
        fminus = fnext;
        Btop = Bb;
    }
    const int64_t oL = oo + (int64_t)L * p.W;
    p.fm[oL] = fminus;
    p.fmm[oL] = 0.0;
    // ---- sweep 3: bottom-up ----
    double fplus = p.hard_surface ? (1.0 - r) * BL * 2 * PB_PI : (BL + b1_last * u) * 2 * PB_PI;
    p.fp[oL] = fplus;
    p.fpm[oL] = 0.0;
    Bbot = BL;
    for (int l = L - 1; l >= 0; --l) {
        const int64_t il = ol + (int64_t)l * ld;
        const int64_t o0 = oo + (int64_t)l * p.W;
        const double pos = p.fp[o0], neg = p.fpm[o0];
        const double dt = p.dtau[il];
        const double Bt = planck(tl[l]);
        TLayer t;
        thermal_layer(dt, p.w0[il], p.cosb[il], Bt, Bbot, t);
        const double Gt = (1 / kMu1 - t.lam) * pos;
        const double Ht = t.gam * (t.lam + 1 / kMu1) * neg;
        const double al1 = 2 * PB_PI * (t.b0 + t.b1 * (t.q - kMu1));
        const double al2 = 2 * PB_PI * t.b1;
        const double xa = exp(-dt / u), xh = exp(-0.5 * dt / u);
        const double EPh = exp(0.5 * t.E), EMh = 1 / EPh;
        const double lu = t.lam * u;
        // fluxes.py:1897-1907
        const double fmid = fplus * xh + (Gt / (lu - 1.0)) * (t.EP * xh - EPh) -
                            (Ht / (lu + 1.0)) * (t.EM * xh - EMh) + al1 * (1. - xh) +
                            al2 * (u + 0.5 * dt - (dt + u) * xh);
        fplus = fplus * xa + (Gt / (lu - 1.0)) * (t.EP * xa - 1.0) +
                (Ht / (lu + 1.0)) * (1.0 - t.EM * xa) + al1 * (1. - xa) +
                al2 * (u - (dt + u) * xa);
        p.fp[o0] = fplus;
        p.fpm[o0] = fmid;
        Bbot = Bt;
    }
    if (p.ftop) p.ftop[((int64_t)b * p.G + a) * p.W + w] = p.fpm[oo];
}

// ---------------------------------------------------------------------------------------
// Level fluxes with precomputed layer records.  therm_levels_kernel is pure single-warp latency when the
// wavelength axis is short (climate solver: 661 waves x 8 gauss points = 166 warps on 148 SMs, 270 serial
// layer steps of ~300 dependent instructions: sqrt, 6 exponentials, 6 divisions, Planck).  Everything that
// does not depend on the solution is therefore evaluated first by a fully parallel kernel (one thread per
// (atmosphere, layer, wavelength[, angle])) into a record array in HBM, with exactly the expressions of the
// one-kernel version (bit-identical results); the three serial sweeps then only load, multiply and add.
// ---------------------------------------------------------------------------------------
enum { LR_LAM = 0, LR_GAM, LR_Q, LR_B0, LR_B1, LR_EP, LR_EM, LR_EPH, LR_EMH, LR_DT, LR_N };

__global__ void __launch_bounds__(128) therm_layer_records_kernel(ThermParams p, double *rec /* [B][L][LR_N][W] */,
                                                                  double *xrec /* [B][G][L][2][W] */)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y, b = blockIdx.z;
    if (w >= p.W) return;
    const int L = p.L, V = L + 1;
    const double *tl = p.tlevel + (int64_t)b * V;
    Planck planck;
    planck.init(p.calc_type, p.wno[w], p.dwno ? p.dwno[w] : 0.0);
    const int64_t il = (int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_layer + (int64_t)l * p.ld + w;
    const double dt = p.dtau[il];
    TLayer t;
    thermal_layer(dt, p.w0[il], p.cosb[il], planck(tl[l]), planck(tl[l + 1]), t);
    double *r = rec + (((int64_t)b * L + l) * LR_N) * p.W + w;
    r[LR_LAM * (int64_t)p.W] = t.lam; r[LR_GAM * (int64_t)p.W] = t.gam; r[LR_Q * (int64_t)p.W] = t.q;
    r[LR_B0 * (int64_t)p.W] = t.b0; r[LR_B1 * (int64_t)p.W] = t.b1;
    r[LR_EP * (int64_t)p.W] = t.EP; r[LR_EM * (int64_t)p.W] = t.EM;
    const double EPh = exp(0.5 * t.E);
    r[LR_EPH * (int64_t)p.W] = EPh; r[LR_EMH * (int64_t)p.W] = 1 / EPh;
    r[LR_DT * (int64_t)p.W] = dt;
    for (int a = 0; a < p.G; ++a) {
        const double u = p.ubar1[a];
        double *x = xrec + ((((int64_t)b * p.G + a) * L + l) * 2) * p.W + w;
        x[0] = exp(-dt / u);
        x[p.W] = exp(-0.5 * dt / u);
    }
}

// same three sweeps as therm_levels_kernel, fed from the records
__global__ void __launch_bounds__(256) therm_levels_rec_kernel(ThermParams p, const double *__restrict__ rec,
                                                               const double *__restrict__ xrec)
{
    const int lane = threadIdx.x;
    const int w = blockIdx.x * kWavesPerCta + lane;
    const int a = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W || a >= p.G) return;
    const int L = p.L, V = p.L + 1;
    const int64_t W = p.W;
    const int64_t oo = (((int64_t)b * p.G + a) * V) * W + w;
    const double *tl = p.tlevel + (int64_t)b * V;
    const double *pl = p.plevel + (int64_t)b * V;
    const double u = p.ubar1[a];
    const double r = p.surf ? p.surf[(int64_t)(p.ob_period ? b % p.ob_period : b) * p.bs_wave + w] : 0.0;
    const double *R = rec + ((int64_t)b * L * LR_N) * W + w;           // + (l * LR_N + field) * W
    const double *X = xrec + ((((int64_t)b * p.G + a) * L) * 2) * W + w;  // + (l * 2 + {0,1}) * W
    Planck planck;
    planck.init(p.calc_type, p.wno[w], p.dwno ? p.dwno[w] : 0.0);
    const double tp = 2 * PB_PI * kMu1;

    // ---- sweep 1: bottom-up elimination ----
    double AS = 0.0, DS = 0.0, gam_n = 0.0, cpu_n = 0.0, cmu_n = 0.0;
    const double BL = planck(tl[L]);
    double b1_last = 0.0;
    for (int l = L - 1; l >= 0; --l) {
        const double *q = R + (int64_t)l * LR_N * W;
        const double gam = q[LR_GAM * W], qq = q[LR_Q * W], b0 = q[LR_B0 * W], b1 = q[LR_B1 * W];
        const double EP = q[LR_EP * W], EM = q[LR_EM * W], dt = q[LR_DT * W];
        // thermal_layer(): fluxes.py:1772-1789
        const double cpu = tp * (b0 + b1 * qq), cmu = tp * (b0 - b1 * qq);
        const double cpd = tp * (b0 + b1 * dt + b1 * qq), cmd = tp * (b0 + b1 * dt - b1 * qq);
        const double e1 = EP + gam * EM, e2 = EP - gam * EM, e3 = gam * EP + EM, e4 = gam * EP - EM;
        if (l == L - 1) {
            b1_last = b1;
            const double b_surface = p.hard_surface ? (1.0 - r) * BL * PB_PI : (BL + b1 * kMu1) * PB_PI;
            const double a_ = e1 - r * e3, b_ = e2 - r * e4;
            const double d_ = b_surface - cpd + r * cmd;
            AS = a_ / b_;
            DS = d_ / b_;
        } else {
            double a_ = 2.0 * (1.0 - gam * gam);
            double b_ = (e1 - e3) * (gam_n + 1.0);
            double c_ = (e1 + e3) * (gam_n - 1.0);
            double d_ = e3 * (cpu_n - cpd) + e1 * (cmd - cmu_n);
            double xi = 1.0 / (b_ - c_ * AS);
            const double ASe = a_ * xi, DSe = (d_ - c_ * DS) * xi;
            const int64_t o1 = oo + (int64_t)(l + 1) * W;
            p.fm[o1] = ASe;
            p.fp[o1] = DSe;
            a_ = (e1 + e3) * (gam_n - 1.0);
            b_ = (e2 + e4) * (gam_n - 1.0);
            c_ = 2.0 * (1.0 - gam_n * gam_n);
            d_ = (gam_n - 1.0) * (cpu_n - cpd) + (1.0 - gam_n) * (cmd - cmu_n);
            xi = 1.0 / (b_ - c_ * ASe);
            AS = a_ * xi;
            DS = (d_ - c_ * DSe) * xi;
        }
        const int64_t o0 = oo + (int64_t)l * W;
        p.fmm[o0] = AS;
        p.fpm[o0] = DS;
        gam_n = gam;
        cpu_n = cpu;
        cmu_n = cmu;
    }
    const double B0 = R[LR_B0 * W];  // b0 of layer 0 = Planck at level 0
    const double tau_top = R[LR_DT * W] * pl[0] / (pl[1] - pl[0]);
    {
        const double b_top = (1.0 - exp(-tau_top / kMu1)) * B0 * PB_PI;
        const double b_ = gam_n + 1.0, c_ = gam_n - 1.0, d_ = b_top - cmu_n;
        const double xi = 1.0 / (b_ - c_ * AS);
        p.fm[oo] = 0.0;
        p.fp[oo] = (d_ - c_ * DS) * xi;
    }
    // ---- sweep 2: top-down substitution + downward recurrence (fluxes.py:1875-1893) ----
    double Xprev = 0.0;
    double fminus = (1 - exp(-tau_top / u)) * B0 * 2 * PB_PI;
    for (int l = 0; l < L; ++l) {
        const double *q = R + (int64_t)l * LR_N * W;
        const int64_t o0 = oo + (int64_t)l * W;
        const double X0 = p.fp[o0] - p.fm[o0] * Xprev;
        const double X1 = p.fpm[o0] - p.fmm[o0] * X0;
        Xprev = X1;
        const double pos = X0 + X1, neg = X0 - X1;
        const double lam = q[LR_LAM * W], gam = q[LR_GAM * W], qq = q[LR_Q * W], b0 = q[LR_B0 * W], b1 = q[LR_B1 * W];
        const double EP = q[LR_EP * W], EM = q[LR_EM * W], EPh = q[LR_EPH * W], EMh = q[LR_EMH * W], dt = q[LR_DT * W];
        const double xa = X[(int64_t)l * 2 * W], xh = X[((int64_t)l * 2 + 1) * W];
        const double J = gam * (lam + 1 / kMu1) * pos;
        const double K = (1 / kMu1 - lam) * neg;
        const double si1 = 2 * PB_PI * (b0 - b1 * (qq - kMu1));
        const double si2 = 2 * PB_PI * b1;
        const double lu = lam * u;
        const double fnext = fminus * xa + (J / (lu + 1.0)) * (EP - xa) + (K / (lu - 1.0)) * (xa - EM) + si1 * (1. - xa) +
                             si2 * (u * xa + dt - u);
        const double fmid = fminus * xh + (J / (lu + 1.0)) * (EPh - xh) + (K / (-lu + 1.0)) * (EMh - xh) + si1 * (1. - xh) +
                            si2 * (u * xh + 0.5 * dt - u);
        p.fm[o0] = fminus;
        p.fmm[o0] = fmid;
        p.fp[o0] = pos;   // parked for sweep 3
        p.fpm[o0] = neg;
        fminus = fnext;
    }
    const int64_t oL = oo + (int64_t)L * W;
    p.fm[oL] = fminus;
    p.fmm[oL] = 0.0;
    // ---- sweep 3: bottom-up upward recurrence (fluxes.py:1897-1907) ----
    double fplus = p.hard_surface ? (1.0 - r) * BL * 2 * PB_PI : (BL + b1_last * u) * 2 * PB_PI;
    p.fp[oL] = fplus;
    p.fpm[oL] = 0.0;
    for (int l = L - 1; l >= 0; --l) {
        const double *q = R + (int64_t)l * LR_N * W;
        const int64_t o0 = oo + (int64_t)l * W;
        const double pos = p.fp[o0], neg = p.fpm[o0];
        const double lam = q[LR_LAM * W], gam = q[LR_GAM * W], qq = q[LR_Q * W], b0 = q[LR_B0 * W], b1 = q[LR_B1 * W];
        const double EP = q[LR_EP * W], EM = q[LR_EM * W], EPh = q[LR_EPH * W], EMh = q[LR_EMH * W], dt = q[LR_DT * W];
        const double xa = X[(int64_t)l * 2 * W], xh = X[((int64_t)l * 2 + 1) * W];
        const double Gt = (1 / kMu1 - lam) * pos;
        const double Ht = gam * (lam + 1 / kMu1) * neg;
        const double al1 = 2 * PB_PI * (b0 + b1 * (qq - kMu1));
        const double al2 = 2 * PB_PI * b1;
        const double lu = lam * u;
        const double fmid = fplus * xh + (Gt / (lu - 1.0)) * (EP * xh - EPh) - (Ht / (lu + 1.0)) * (EM * xh - EMh) +
                            al1 * (1. - xh) + al2 * (u + 0.5 * dt - (dt + u) * xh);
        fplus = fplus * xa + (Gt / (lu - 1.0)) * (EP * xa - 1.0) + (Ht / (lu + 1.0)) * (1.0 - EM * xa) + al1 * (1. - xa) +
                al2 * (u - (dt + u) * xa);
        p.fp[o0] = fplus;
        p.fpm[o0] = fmid;
    }
    if (p.ftop) p.ftop[((int64_t)b * p.G + a) * W + w] = p.fpm[oo];
}

__global__ void compress_thermal_kernel(int64_t n, int G, int nt, const double *flux,
                                        const double *gweight, const double *tweight, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n) return;
    double acc = 0.0;
    for (int a = 0; a < G; ++a) {
        const int ig = a / nt, it = a - ig * nt;
        acc = acc + flux[((int64_t)b * G + a) * n + i] * gweight[ig] * tweight[it];
    }
    const double sym = (nt == 1) ? 1.0 : 1 / (2 * PB_PI);
    out[(int64_t)b * n + i] = acc * sym;
}

} // namespace

extern "C" int pb_thermal_toon_1d(pb_ctx *ctx, const pb_thermal_args *a, int memspace)
{
    if (!ctx || !a) return PB_ERR_ARG;
    const int L = a->nlayer, W = a->nwno, G = a->numg * a->numt;
    const int B = a->nbatch > 0 ? a->nbatch : 1;
    const int V = L + 1;
    if (L < 1 || W < 0 || G < 1) return pb_fail(ctx, PB_ERR_ARG, "thermal: bad sizes L=%d W=%d G=%d", L, W, G);
    if (B > 65535) return pb_fail(ctx, PB_ERR_ARG, "thermal: nbatch = %d exceeds 65535 (batch entries map to gridDim.y/z); split the batch", B);
    if (W == 0) return PB_OK;
    if (a->ld < W) return pb_fail(ctx, PB_ERR_ARG, "thermal: ld < nwno");
    if (!a->dtau || !a->w0 || !a->cosb || !a->wno || !a->tlevel || !a->plevel || !a->ubar1)
        return pb_fail(ctx, PB_ERR_ARG, "thermal: NULL input array");
    if (a->calc_type != 0 && a->calc_type != 1) return pb_fail(ctx, PB_ERR_ARG, "thermal: calc_type must be 0 or 1");
    if (a->calc_type == 1 && !a->dwno) return pb_fail(ctx, PB_ERR_ARG, "thermal: calc_type=1 needs dwno");
    if (a->thermal && (!a->gweight || !a->tweight)) return pb_fail(ctx, PB_ERR_ARG, "thermal: thermal output needs gweight/tweight");
    const bool want_lvl = a->flux_minus || a->flux_plus || a->flux_minus_mdpt || a->flux_plus_mdpt;
    if (a->variant != 0 && a->variant != 1) return pb_fail(ctx, PB_ERR_ARG, "thermal: variant must be 0 or 1");
    if (a->variant == 1 && (G != 1 || want_lvl || a->thermal || a->calc_type != 0))
        return pb_fail(ctx, PB_ERR_ARG, "thermal: variant 1 (3-D facets) needs numg=numt=1 per facet, calc_type 0, TOA only");
    if (want_lvl && (!a->flux_minus || !a->flux_plus || !a->flux_minus_mdpt || !a->flux_plus_mdpt))
        return pb_fail(ctx, PB_ERR_ARG, "thermal: level fluxes need all four arrays");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    if (a->opacity_period > 0 && (memspace != PB_DEVICE || a->variant || a->opacity_period > B))
        return pb_fail(ctx, PB_ERR_ARG, "thermal: opacity_period needs PB_DEVICE arrays, variant 0 and period <= nbatch");
    const size_t nW = (size_t)W * sizeof(double);
    const bool fuse = a->thermal && !want_lvl && G <= 8;
    const bool need_ftop = a->flux_at_top || (a->thermal && !fuse);

    size_t need = 16 * 256 + 2 * pb_align((size_t)B * V * 8) + 3 * pb_align((size_t)G * 8);
    if (host) {
        need += 3 * pb_align((size_t)B * L * nW) + 2 * pb_align(nW) + pb_align(B * nW);
        need += pb_align((size_t)B * G * nW) + pb_align(B * nW);
        if (want_lvl) need += 4 * pb_align((size_t)B * G * V * nW);
    } else {
        need += pb_align((size_t)B * G * nW);
    }
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, need));
    PB_TRY(pb_pinned_reserve(ctx, (2 * (size_t)B * V + 3 * (size_t)G + (size_t)B + 64) * sizeof(double)));

    ThermParams p;
    memset(&p, 0, sizeof(p));
    p.L = L; p.W = W; p.G = G; p.nt = a->numt;
    int64_t ldo;
    PB_TRY(pb_stage_in(ctx, a->dtau, memspace, (int64_t)B * L, W, a->ld, &p.dtau, &ldo));
    PB_TRY(pb_stage_in(ctx, a->w0, memspace, (int64_t)B * L, W, a->ld, &p.w0, &ldo));
    PB_TRY(pb_stage_in(ctx, a->cosb, memspace, (int64_t)B * L, W, a->ld, &p.cosb, &ldo));
    PB_TRY(pb_stage_in(ctx, a->wno, memspace, 1, W, W, &p.wno, &ldo));
    PB_TRY(pb_stage_in(ctx, a->dwno, memspace, 1, W, W, &p.dwno, &ldo));
    PB_TRY(pb_stage_in(ctx, a->surf_reflect, memspace, a->variant ? 1 : B, W, W, &p.surf, &ldo));
    p.ld = host ? W : a->ld;
    p.bs_layer = (int64_t)L * p.ld; p.bs_wave = W;
    p.ob_period = a->opacity_period > 0 ? a->opacity_period : 0;
    PB_TRY(pb_upload_small(ctx, a->tlevel, (size_t)B * V, &p.tlevel));
    PB_TRY(pb_upload_small(ctx, a->plevel, (size_t)B * V, &p.plevel));
    PB_TRY(pb_upload_small(ctx, a->ubar1, a->variant ? B : G, &p.ubar1));
    if (a->gweight) PB_TRY(pb_upload_small(ctx, a->gweight, a->numg, &p.gweight));
    if (a->tweight) PB_TRY(pb_upload_small(ctx, a->tweight, a->numt, &p.tweight));
    p.hard_surface = a->hard_surface; p.calc_type = a->calc_type; p.variant = a->variant;
    if (a->variant) p.bs_wave = 0;  // facets share surf_reflect

    double *d_ftop = nullptr, *d_th = nullptr, *d_lv[4] = {nullptr, nullptr, nullptr, nullptr};
    double *h_lv[4] = {a->flux_minus, a->flux_plus, a->flux_minus_mdpt, a->flux_plus_mdpt};
    if (host) {
        if (need_ftop) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * nW, (void **)&d_ftop));
        if (a->thermal) PB_TRY(pb_arena_alloc(ctx, B * nW, (void **)&d_th));
        if (want_lvl) for (int k = 0; k < 4; ++k) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * V * nW, (void **)&d_lv[k]));
    } else {
        d_ftop = a->flux_at_top;
        if (!d_ftop && need_ftop) PB_TRY(pb_arena_alloc(ctx, (size_t)B * G * nW, (void **)&d_ftop));
        d_th = a->thermal;
        for (int k = 0; k < 4; ++k) d_lv[k] = h_lv[k];
    }
    PB_TRY(pb_upload_flush(ctx));
    const int ay = G < 8 ? G : 8;
    dim3 block(kWavesPerCta, ay, 1);
    dim3 grid((W + kWavesPerCta - 1) / kWavesPerCta, (G + ay - 1) / ay, B);
    p.ftop = d_ftop; p.thermal = d_th;
    // one thread per wavelength (all angles in registers) once the wavelength axis alone fills the
    // machine: >= PB_THERM_WAVE_MIN (default 32 768) wavelengths x atmospheres.  PB_THERM_KERNEL=angle|wave forces one.
    static const long wave_min = []() { const char *e = getenv("PB_THERM_WAVE_MIN"); return e ? atol(e) : 32768L; }();
    const char *force = getenv("PB_THERM_KERNEL");
    bool use_wave_kernel = !want_lvl && a->variant == 0 && G <= 8 && (!a->thermal || fuse) && (long)W * B >= wave_min;
    if (force && force[0] == 'a') use_wave_kernel = false;
    if (force && force[0] == 'w' && !want_lvl && a->variant == 0 && G <= 8 && (!a->thermal || fuse)) use_wave_kernel = true;
    // angle-parallel launches: the chain-warp kernel shares the elimination between the angles of a wavelength
    // (PB_THERM_KERNEL=chain opts in, =angle keeps therm_toa_kernel); it needs all angles of a wavelength in one CTA
    // and its tiles (3 CTAs per SM) in shared memory
    const size_t chain_smem = ((size_t)kThermChainTab + (size_t)V * 32 + (size_t)kThermChainCh * (3 * TNQ + 2 * TNC) * 32 + 32) * sizeof(double);
    bool use_chain_kernel = kThermChainDefault && !want_lvl && !use_wave_kernel && a->variant == 0 && G >= 2 && G <= 8 && chain_smem <= 75 * 1024;
    if (force && force[0] == 'c' && !want_lvl && a->variant == 0 && G >= 2 && G <= 8 && chain_smem <= 200 * 1024) {
        use_chain_kernel = true;
        use_wave_kernel = false;
    }
    if (force && force[0] == 'a') use_chain_kernel = false;
    if (want_lvl) {
        p.fm = d_lv[0]; p.fp = d_lv[1]; p.fmm = d_lv[2]; p.fpm = d_lv[3];
        // Short wavelength axes (the climate solver) are serial-latency bound: precompute the layer records with
        // a fully parallel kernel when the launch cannot fill the machine anyway.  PB_THERM_LEVELS=rec|fused forces.
        const size_t rec_bytes = ((size_t)B * L * LR_N + (size_t)B * G * L * 2) * nW;
        const char *lf = getenv("PB_THERM_LEVELS");
        bool use_rec = (long)W * B * G <= 64L * 1024 && rec_bytes <= ((size_t)1 << 31);
        if (lf && lf[0] == 'r' && rec_bytes <= ((size_t)1 << 31)) use_rec = true;
        if (lf && lf[0] == 'f') use_rec = false;
        if (use_rec) {
            if (rec_bytes > ctx->rec_cap) {
                PB_CUDA(ctx, cudaDeviceSynchronize());
                if (ctx->rec) PB_CUDA(ctx, cudaFree(ctx->rec));
                ctx->rec = nullptr; ctx->rec_cap = 0;
                const size_t cap = pb_align(rec_bytes + rec_bytes / 8, 1 << 20);
                cudaError_t e = cudaMalloc((void **)&ctx->rec, cap);
                if (e != cudaSuccess) return pb_fail(ctx, PB_ERR_NOMEM, "thermal: layer-record cudaMalloc(%zu) -> %s", cap, cudaGetErrorString(e));
                ctx->rec_cap = cap;
            }
            double *rec = (double *)ctx->rec, *xrec = rec + (size_t)B * L * LR_N * W;
            dim3 g1((W + 127) / 128, L, B);
            therm_layer_records_kernel<<<g1, 128, 0, ctx->stream>>>(p, rec, xrec);
            PB_CHECK_LAUNCH(ctx);
            therm_levels_rec_kernel<<<grid, block, 0, ctx->stream>>>(p, rec, xrec);
            PB_CHECK_LAUNCH(ctx);
        } else {
            therm_levels_kernel<<<grid, block, 0, ctx->stream>>>(p);
            PB_CHECK_LAUNCH(ctx);
        }
    } else if (use_wave_kernel) {
        p.fuse = fuse ? 1 : 0;
        switch (G) {
        case 1: launch_therm_wave<1>(p, B, ctx->stream); break;
        case 2: launch_therm_wave<2>(p, B, ctx->stream); break;
        case 3: launch_therm_wave<3>(p, B, ctx->stream); break;
        case 4: launch_therm_wave<4>(p, B, ctx->stream); break;
        case 5: launch_therm_wave<5>(p, B, ctx->stream); break;
        case 6: launch_therm_wave<6>(p, B, ctx->stream); break;
        case 7: launch_therm_wave<7>(p, B, ctx->stream); break;
        default: launch_therm_wave<8>(p, B, ctx->stream); break;
        }
        PB_CHECK_LAUNCH(ctx);
    } else if (use_chain_kernel) {
        p.fuse = fuse ? 1 : 0;
        const int nsm = ctx->sm_count > 0 ? ctx->sm_count : 148;
        const char *wte = getenv("PB_THERM_WT");
        int wt = wte ? atoi(wte) : 0;
        const int cap = 160 / ay < 32 ? 160 / ay : 32;   // at most five consumer warps
        if (wt <= 0 || wt > cap) {
            wt = cap;
            // narrowed, when the whole launch is a single residency wave (3 CTAs per SM), to an even CTA count per SM
            const int std_ctas = (W + cap - 1) / cap;
            const int per_sm = (std_ctas + nsm - 1) / nsm;
            if (B == 1 && per_sm <= 3 && (double)per_sm * nsm > 1.05 * std_ctas) {
                const int cand = (W + nsm * per_sm - 1) / (nsm * per_sm);
                if (cand >= 12 && cand < cap) wt = cand;
            }
        }
        p.wt = wt; p.ay = ay;
        const int nw = (wt * ay + 31) / 32 + 1;
        p.ch = nw < kThermChainCh ? nw : kThermChainCh;
        const size_t smem = ((size_t)kThermChainTab + (size_t)V * 32 + (size_t)p.ch * (3 * TNQ + 2 * TNC) * 32 + 32) * sizeof(double);
        if (smem > 48 * 1024) PB_CUDA(ctx, pb_ensure_smem(ctx, therm_toa_chain_kernel, smem));
        dim3 cgrid((W + wt - 1) / wt, 1, B);
        therm_toa_chain_kernel<<<cgrid, nw * 32, smem, ctx->stream>>>(p);
        PB_CHECK_LAUNCH(ctx);
    } else {
        p.fuse = fuse ? 1 : 0;
        // tile width from the SM count when 32-wide tiles give an uneven single residency wave (see toon_reflected.cu)
        const char *wte = getenv("PB_THERM_WT");
        int wt = wte ? atoi(wte) : 0;
        if (wt <= 0 || wt > 32) {
            wt = 32;
            const int nsm = ctx->sm_count > 0 ? ctx->sm_count : 148;
            const int std_ctas = (W + 31) / 32;
            const int per_sm = (std_ctas + nsm - 1) / nsm;
            const int cap = 16 / ay;
            if (B == 1 && G <= 8 && a->variant == 0 && per_sm >= 2 && per_sm <= cap && (double)per_sm * nsm > 1.15 * std_ctas) {
                const int cand = (W + nsm * per_sm - 1) / (nsm * per_sm);
                if (cand >= 16 && cand < 32) wt = cand;
            }
        }
        if (wt == 32) {
            const size_t smem = ((size_t)V * 32 + (size_t)2 * ay * TNQ * 32) * sizeof(double);
            if (smem > 200 * 1024) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "thermal: nlevel=%d exceeds the shared-memory Planck tile", V);
            if (smem > 48 * 1024)
                PB_CUDA(ctx, pb_ensure_smem(ctx, therm_toa_kernel<false>, smem));
            therm_toa_kernel<false><<<grid, block, smem, ctx->stream>>>(p);
        } else {
            p.wt = wt; p.ay = ay;
            const int nthreads = (wt * ay + 31) / 32 * 32, nwarp = nthreads / 32;
            const size_t smem = ((size_t)V * 32 + (size_t)2 * nwarp * TNQ * 32) * sizeof(double);
            if (smem > 200 * 1024) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "thermal: nlevel=%d exceeds the shared-memory Planck tile", V);
            if (smem > 48 * 1024)
                PB_CUDA(ctx, pb_ensure_smem(ctx, therm_toa_kernel<true>, smem));
            dim3 ggrid((W + wt - 1) / wt, (G + ay - 1) / ay, B);
            therm_toa_kernel<true><<<ggrid, nthreads, smem, ctx->stream>>>(p);
        }
        PB_CHECK_LAUNCH(ctx);
    }
    if (a->thermal && !(fuse && !want_lvl)) {
        dim3 g2((W + 127) / 128, B);
        compress_thermal_kernel<<<g2, 128, 0, ctx->stream>>>(W, G, a->numt, d_ftop, p.gweight, p.tweight, d_th);
        PB_CHECK_LAUNCH(ctx);
    }
    if (host) {
        if (a->flux_at_top) PB_CUDA(ctx, cudaMemcpyAsync(a->flux_at_top, d_ftop, (size_t)B * G * nW, cudaMemcpyDeviceToHost, ctx->stream));
        if (a->thermal) PB_CUDA(ctx, cudaMemcpyAsync(a->thermal, d_th, B * nW, cudaMemcpyDeviceToHost, ctx->stream));
        if (want_lvl) for (int k = 0; k < 4; ++k)
            PB_CUDA(ctx, cudaMemcpyAsync(h_lv[k], d_lv[k], (size_t)B * G * V * nW, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}

extern "C" int pb_compress_thermal(pb_ctx *ctx, int64_t n, const double *flux, const double *gweight,
                                   int ng, const double *tweight, int nt, double *out, int memspace)
{
    if (!ctx || !flux || !gweight || !tweight || !out || ng < 1 || nt < 1 || n < 0)
        return pb_fail(ctx, PB_ERR_ARG, "compress_thermal: bad arguments");
    if (n == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int G = ng * nt;
    const size_t nb = (size_t)n * sizeof(double);
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, 8 * 256 + 2 * pb_align((size_t)G * 8) + pb_align((size_t)G * nb) + pb_align(nb)));
    PB_TRY(pb_pinned_reserve(ctx, 2 * ((size_t)G + 16) * sizeof(double)));
    const double *d_x, *d_gw, *d_tw;
    int64_t ldo;
    PB_TRY(pb_stage_in(ctx, flux, memspace, G, n, n, &d_x, &ldo));
    PB_TRY(pb_upload_small(ctx, gweight, ng, &d_gw));
    PB_TRY(pb_upload_small(ctx, tweight, nt, &d_tw));
    double *d_out = out;
    if (memspace == PB_HOST) PB_TRY(pb_arena_alloc(ctx, nb, (void **)&d_out));
    PB_TRY(pb_upload_flush(ctx));
    dim3 grid((unsigned)((n + 127) / 128), 1);
    compress_thermal_kernel<<<grid, 128, 0, ctx->stream>>>(n, G, nt, d_x, d_gw, d_tw, d_out);
    PB_CHECK_LAUNCH(ctx);
    if (memspace == PB_HOST) {
        PB_CUDA(ctx, cudaMemcpyAsync(out, d_out, nb, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
