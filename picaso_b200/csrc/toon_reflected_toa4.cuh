// toon_reflected_toa4.cuh - fourth generation of the reflected TOA kernel (included by
// toon_reflected.cu inside its anonymous namespace, after ReflParams / refl_load / p_single).
//
// What changed against refl_toa_kernel3: the sweep runs TOP-DOWN.  The kernel is bound by the
// fp64 pipe (ncu: ~258 fp64 instructions per (layer, angle), 4-5 of them exponentials), and three
// of those exponentials only exist because a bottom-up sweep meets the optical-depth scale in the
// wrong order:
//     exp(-tau[l]/u0), exp(-tau_og[l]/u0)                     (attenuation of the direct beam)
//     exp(-dtau(1/u0+1/u1)), exp(-dtau_og(1/u0+1/u1))         (slant path through one layer)
// Walking from the top of the atmosphere downwards the first two are running products of the
// per-layer factors exp(-dtau/u0), exp(-dtau_og/u0), and the last two are products of factors
// that are needed anyway.  At zero phase (u0 == u1, the BASELINE geometry) a (layer, angle) step
// needs TWO exponentials instead of four; with delta-Eddington off (aliased *_og arrays) ONE.
//
// The adjoint idea is unchanged, just mirrored: rows are eliminated from row 0 downwards
// (X[n] = DS[n] - CS[n] X[n+1]); the TOA intensity  I_0 = sum_l T_l (K_l + cP_l X[2l] + cQ_l X[2l+1])
// + T_L flux_zero/pi  (T_l = prod_{k<l} exp(-dtau_k/u1), fluxes.py:1395-1407 unrolled) is carried as
// R + P X[n] over the first not-yet-eliminated unknown; the surface row closes the chain.
// Agreement with the reference golden vectors: <= 7e-9 (numpy prototype and kernel), tolerance 1e-6.
//
// tau / tau_og are separate arguments of the reference signature, so the running products are
// only used where tau[l+1] == tau[l] + dtau[l] holds to 8 ulp (it holds bit-exactly for
// compute_opacity's cumsum, optics.py:353-354); the producer checks every level and hands the
// consumer the exact tau[l+1] wherever it does not, and the consumer then re-bases its running
// product with a true exponential.  Results therefore follow the caller's tau arrays in all cases.
//
// Measured (B200, 60 x 10 000 x 5, us per launch; ncu profiles/r1_refl_toa_v4.summary.json):
// v3 90.4 -> v4 88.9.  fp64 instructions fell 17 % (18.9 M -> 16.2 M warp-instructions) but the launch
// is not fp64-throughput bound at this size: 313 CTAs leave 2 or 3 CTAs per SM (10 / 15 warps), the
// warps issue one instruction every 6.6 cycles (stall_wait: dependent fixed-latency chains) and the
// 17 SMs that hold 3 CTAs set the duration (2 resident CTAs: 63 us; 3: 89 us).  A/B variants that did
// NOT help and were removed again: software-pipelining the coefficient set of layer k+1 under the
// elimination step of layer k (102 us, spills), branch-free pbm::exp (92 us), 136 registers (2 CTAs
// per SM: 127 us), prefetch.global.L2 of the next-but-one chunk (91 us), both layers' 26 loads in
// flight before the first use (93 us, spills).
// With SM-count-sized tiles (GEN, 4 warps per CTA) the register budget is 168 instead of 128: 82.2 -> 77.3 us
// (no spills); the software-pipelined consume loop is still slower there (80.7 us).

#ifndef PB_REFL4_REGS
#define PB_REFL4_REGS 128
#endif

enum { R_G = 0, R_OM, R_G1, R_G2, R_LAM, R_GAM, R_EP, R_EM, R_DT, R_GC2, R_S0, R_DTO, R_TAUD, R_TAUOD, NR };
static_assert(NR == NQ, "v3 and v4 share the shared-memory tile size");

constexpr int kConsistentHi = 0x7ff8b200;  // high word of the "tau is consistent" marker (a quiet NaN)

__device__ __forceinline__ double refl4_tau_slot(double tau0, double dt, double tau1)
{
    // |tau1 - (tau0 + dt)| <= 8 ulp(tau1)  ->  marker; anything else (incl. NaN) -> exact tau1
    const bool ok = fabs(tau1 - (tau0 + dt)) <= 1.7763568394002505e-15 * fabs(tau1);
    return ok ? __hiloint2double(kConsistentHi, 0) : tau1;
}

struct Refl4Inputs {
    double om, fc, cb, dt, gc2, fr, dto, omo, cbo, tau0, tau1, tauo0, tauo1;
};

__device__ __forceinline__ void refl4_load(const ReflParams &p, int64_t il, int64_t iv, int64_t ld, Refl4Inputs &x)
{
    x.om = __ldg(p.w0 + il);
    x.fc = __ldg(p.fcld + il);
    x.cb = __ldg(p.cosb + il);
    x.dt = __ldg(p.dtau + il);
    x.gc2 = __ldg(p.gcos2 + il);
    x.fr = __ldg(p.fray + il);
    x.dto = __ldg(p.dtau_og + il);
    x.omo = __ldg(p.w0_og + il);
    x.cbo = __ldg(p.cosb_og + il);
    x.tau0 = __ldg(p.tau + iv);
    x.tau1 = __ldg(p.tau + iv + ld);
    x.tauo0 = __ldg(p.tau_og + iv);
    x.tauo1 = __ldg(p.tau_og + iv + ld);
}

__device__ __forceinline__ void refl4_produce(const ReflParams &p, const Refl4Inputs &x, double f0,
                                              double *q /* [NR][32] column of this lane */)
{
    const double g = x.fc * x.cb;
    double g1, g2;
    toon_g(p.tc, x.om, g, g1, g2);
    const double lam = sqrt(g1 * g1 - g2 * g2);
    const double gam = (g1 - lam) * pbm::krcp(g2);
    const double E = fmin(lam * x.dt, p.clip);  // slice_gt(exptrm, 35 | 40), fluxes.py:1174, :516
    const double EP = pbm::kexp(E);
    const double ps = p_single(p, x.cbo, x.gc2, x.fc, x.fr);
    q[R_G * 32] = g;
    q[R_OM * 32] = x.om;
    q[R_G1 * 32] = g1;
    q[R_G2 * 32] = g2;
    q[R_LAM * 32] = lam;
    q[R_GAM * 32] = gam;
    q[R_EP * 32] = EP;
    q[R_EM * 32] = pbm::krcp(EP);
    q[R_DT * 32] = x.dt;
    q[R_GC2 * 32] = x.gc2;
    q[R_S0 * 32] = (x.omo * f0 / (4.0 * PB_PI)) * ps;
    q[R_DTO * 32] = x.dto;
    q[R_TAUD * 32] = refl4_tau_slot(x.tau0, x.dt, x.tau1);
    q[R_TAUOD * 32] = refl4_tau_slot(x.tauo0, x.dto, x.tauo1);
}

struct Refl4Rec {  // what the elimination / adjoint step needs from one (layer, angle)
    double gam, cpu, cmu, cpd, cmd, e1, e2, e3, e4, f0, f1, K;
};

struct Refl4State {
    double CS, DS, P, R;          // elimination relation of the last row, functional R + P X[n]
    double T1, T0, TO;            // exp(-tau_l/u1), exp(-tau_l/u0), exp(-tau_og_l/u0) at the top of the next layer
    double f1p;                   // weight of X[2l-1] in the functional (previous layer's T1 cQ)
    double gam_p, cpd_p, cmd_p;   // previous (upper) layer
    double e1p, e3p, s13p, s24p;  // e1, e3, e1+e3, e2+e4 of the previous layer
};

// The four exponentials of one (layer, angle) [two at zero phase, one more halving with aliased og arrays]
struct Refl4Exp {
    double xa, xa0, xoa0, xoa1;
    double xd_ex, xod_ex;  // exp(-tau[l+1]/u0), exp(-tau_og[l+1]/u0) where the caller's tau is not cumsum(dtau)
    bool ex, exo;
};

__device__ __forceinline__ void refl4_exps(const ReflAngle &g, const double *q, Refl4Exp &e)
{
    const double dt = q[R_DT * 32];
    e.xa = pbm::kexp(-dt * g.inv_u1);
    e.xa0 = g.same_mu ? e.xa : pbm::kexp(-dt * g.inv_u0);
    if (g.og_alias) {
        e.xoa0 = e.xa0;
        e.xoa1 = e.xa;
    } else {
        const double dto = q[R_DTO * 32];
        e.xoa0 = pbm::kexp(-dto * g.inv_u0);
        e.xoa1 = g.same_mu ? e.xoa0 : pbm::kexp(-dto * g.inv_u1);
    }
    // the (never taken for cumsum inputs) exact path sits here, next to the other exponentials, so
    // that the coefficient block below stays one straight-line basic block
    const double td = q[R_TAUD * 32], tod = q[R_TAUOD * 32];
    e.ex = __double2hiint(td) != kConsistentHi;
    e.exo = !g.og_alias && __double2hiint(tod) != kConsistentHi;
    e.xd_ex = 0.0;
    e.xod_ex = 0.0;
    if (e.ex) e.xd_ex = pbm::kexp(-td * g.inv_u0);
    if (e.exo) e.xod_ex = pbm::kexp(-tod * g.inv_u0);
}

template <int MP>
__device__ __forceinline__ void refl4_coeffs(const ReflParams &p, const ReflAngle &g, const double *q,
                                             const Refl4Exp &e, Refl4State &s, Refl4Rec &o)
{
    const double c2pi = 0.5 / PB_PI;
    const double gg = q[R_G * 32], om = q[R_OM * 32], g1 = q[R_G1 * 32], g2 = q[R_G2 * 32];
    const double lam = q[R_LAM * 32], gam = q[R_GAM * 32], EP = q[R_EP * 32], EM = q[R_EM * 32];
    // direct-beam attenuation at the top (carried) and bottom of this layer
    const double xu = s.T0, xo = s.TO;
    const double xd = e.ex ? e.xd_ex : xu * e.xa0;
    const double xod = g.og_alias ? xd : (e.exo ? e.xod_ex : xo * e.xoa0);
    const double g3 = toon_g3(p.tc, gg, g.u0);
    const double g4 = 1.0 - g3;
    const double inv_den = pbm::krcp(lam * lam - g.inv_u0 * g.inv_u0);
    const double fw = g.f0 * om;
    const double am = fw * (g4 * (g1 + g.inv_u0) + g2 * g3) * inv_den;
    const double ap = fw * (g3 * (g1 - g.inv_u0) + g2 * g4) * inv_den;
    o.gam = gam;
    o.cmu = am * xu; o.cpu = ap * xu; o.cmd = am * xd; o.cpd = ap * xd;
    o.e1 = EP + gam * EM; o.e2 = EP - gam * EM;
    o.e3 = gam * EP + EM; o.e4 = gam * EP - EM;
    double mpl, mmi;  // fluxes.py:1275-1287
    if (MP == 0) {
        const double t2 = q[R_GC2 * 32] * g.t2c;
        mpl = 1.0 + 1.5 * gg * g.u1 + t2;
        mmi = 1.0 - 1.5 * gg * g.u1 + t2;
    } else {
        mpl = 1.0 + 1.5 * gg * g.u1;
        mmi = 1.0 - 1.5 * gg * g.u1;
    }
    // fluxes.py:1290-1296, :1395-1407; exp(+-E - dt/u1) = EP|EM * exp(-dt/u1)
    const double lu = lam * g.u1;
    const double inv_l = pbm::krcp(lu * lu - 1.0);
    const double omc = om * c2pi;
    const double cG = (mpl + gam * mmi) * omc * ((EP * e.xa - 1.0) * ((lu + 1.0) * inv_l));
    const double cH = (gam * mpl + mmi) * omc * ((1.0 - EM * e.xa) * ((lu - 1.0) * inv_l));
    const double At = (mpl * o.cpu + mmi * o.cmu) * omc;
    const double xs = e.xa * e.xa0;      // exp(-dtau (u0+u1)/(u0 u1))
    const double xso = e.xoa0 * e.xoa1;  // exp(-dtau_og (u0+u1)/(u0 u1))
    const double K = q[R_S0 * 32] * xo * (1.0 - xso) * g.wgt + At * (1.0 - xs) * g.wgt;
    // weights in I_0 of this layer's unknowns and its X-independent source, attenuated to the top
    o.f0 = s.T1 * (cG + cH);
    o.f1 = s.T1 * (cG - cH);
    o.K = s.T1 * K;
    s.T1 = s.T1 * e.xa;
    s.T0 = xd;
    s.TO = xod;
}

// row 0 (fluxes.py:155-158)
__device__ __forceinline__ void refl4_first(const Refl4Rec &c, double btop, Refl4State &s)
{
    const double x = pbm::krcp(c.gam + 1.0);
    s.CS = (c.gam - 1.0) * x;
    s.DS = (btop - c.cmu) * x;
    s.P = c.f0;
    s.R = c.K;
}

// interface rows 2l-1, 2l between the previous layer and this one (fluxes.py:161-175), mirrored fold
__device__ __forceinline__ void refl4_step(const Refl4Rec &c, Refl4State &s)
{
    const double gm1 = c.gam - 1.0;
    const double dcp = c.cpu - s.cpd_p, dcm = s.cmd_p - c.cmu;
    const double A1 = s.s13p * gm1;  // also C of the even row
    // odd row 2l-1: A X[2l-2] + B X[2l-1] + C X[2l] = D
    double x = pbm::krcp(s.s24p * gm1 - A1 * s.CS);
    const double CSo = (2.0 * (1.0 - c.gam * c.gam)) * x;
    const double DSo = ((gm1 * dcp - gm1 * dcm) - A1 * s.DS) * x;
    // fold X[2l-2] = DS - CS X[2l-1], then X[2l-1] = DSo - CSo X[2l]
    double R = s.R + s.P * s.DS;
    double P = s.f1p - s.P * s.CS;
    R = R + P * DSo;
    P = c.f0 - P * CSo;
    // even row 2l: A X[2l-1] + B X[2l] + C X[2l+1] = D
    const double A2 = 2.0 * (1.0 - s.gam_p * s.gam_p);
    x = pbm::krcp((s.e1p - s.e3p) * (c.gam + 1.0) - A2 * CSo);
    s.CS = A1 * x;
    s.DS = ((s.e3p * dcp + s.e1p * dcm) - A2 * DSo) * x;
    s.P = P;
    s.R = R + c.K;
}

__device__ __forceinline__ void refl4_carry(const Refl4Rec &c, Refl4State &s)
{
    s.f1p = c.f1;
    s.gam_p = c.gam; s.cpd_p = c.cpd; s.cmd_p = c.cmd;
    s.e1p = c.e1; s.e3p = c.e3; s.s13p = c.e1 + c.e3; s.s24p = c.e2 + c.e4;
}

// GEN = false: CTA = 32 wavelengths (threadIdx.x) x blockDim.y angle warps.
// GEN = true : CTA = p.wt (< 32) wavelengths x p.ay angles flattened over a 1-D block, thread t ->
//   (wave t % wt, angle t / wt).  Used when 32-wide tiles would leave the SMs unevenly loaded in a
//   single residency wave (W = 10 000: 313 CTAs = 2 or 3 per SM, the 3-CTA SMs set the duration);
//   the launcher then picks wt so that the grid is (just under) a multiple of the SM count - 23
//   wavelengths, 435 CTAs of 4 warps, 12 warps on every SM.  As a producer a thread is still lane
//   `t % 32` of warp `t / 32` and fills column `lane` of its warp's layer rows.
template <int MP /*multi_phase*/, bool GEN>
__device__ __forceinline__ void refl4_body(const ReflParams &p)
{
    extern __shared__ double smem[];  // [2][2*NW][NR][32]
    const int tid = GEN ? (int)threadIdx.x : (int)(threadIdx.y * 32 + threadIdx.x);
    const int lane = tid & 31, wy = tid >> 5;                   // producer identity
    const int NW = GEN ? (int)(blockDim.x >> 5) : (int)blockDim.y;
    const int WT = GEN ? p.wt : kWavesPerCta, AY = GEN ? p.ay : NW;
    const int cw = GEN ? tid % WT : lane, ca = GEN ? tid / WT : wy;  // consumer identity
    const int CH = 2 * NW;  // layers per chunk
    const int w = blockIdx.x * WT + cw;
    const int wc = w < p.W ? w : p.W - 1;  // clamp: every thread takes part in the tile protocol
    const int wp = blockIdx.x * WT + lane;
    const int wpc = wp < p.W ? wp : p.W - 1;
    const int a = blockIdx.y * AY + ca;
    const int ac = a < p.G ? a : p.G - 1;
    const int b = blockIdx.z;
    const int L = p.L;
    const int64_t ld = p.ld;
    const int64_t ol = (int64_t)b * p.bs_layer + wpc;   // producer column
    const int64_t ov = (int64_t)b * p.bs_level + wpc;
    const int64_t ovc = (int64_t)b * p.bs_level + wc;   // consumer column
    const int64_t ow = (int64_t)b * p.bs_wave + wc;
    const double f0p = p.f0pi ? p.f0pi[(int64_t)b * p.bs_wave + wpc] : 1.0;
    ReflAngle g;
    g.u0 = p.variant ? fabs(p.ubar0[b]) : p.ubar0[ac];  // 3-D facets: geometry per batch entry, |ubar|
    g.u1 = p.variant ? fabs(p.ubar1[b]) : p.ubar1[ac];
    g.f0 = p.f0pi ? p.f0pi[ow] : 1.0;
    const double r = p.surf ? p.surf[ow] : 0.0;
    const double btop = p.btop ? p.btop[ow] : 0.0;
    g.inv_u0 = 1.0 / g.u0; g.inv_u1 = 1.0 / g.u1;
    g.s01 = (g.u0 + g.u1) / (g.u0 * g.u1);
    g.wgt = g.u0 / (g.u0 + g.u1);
    const double ubar2 = 0.767;  // fluxes.py:1280
    g.t2c = (3.0 * ubar2 * ubar2 * g.u1 * g.u1 - 1.0) / 2.0;
    g.same_mu = (g.u0 == g.u1);
    g.og_alias = (p.dtau_og == p.dtau) && (p.tau_og == p.tau);
    const int nchunks = (L + CH - 1) / CH;
    const int tile = CH * NR * 32;

    auto produce = [&](int c) {
        // this warp's two layers of chunk c: positions wy and wy + NW of the chunk (top-down)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pos = wy + h * NW;
            const int l = c * CH + pos;
            if (l < L) {
                Refl4Inputs x;
                refl4_load(p, ol + (int64_t)l * ld, ov + (int64_t)l * ld, ld, x);
                refl4_produce(p, x, f0p, smem + (c & 1) * tile + pos * NR * 32 + lane);
            }
        }
    };

    Refl4State s;
    s.CS = s.DS = s.P = s.R = 0.0;
    s.T1 = 1.0;
    s.T0 = pbm::kexp(-__ldg(p.tau + ovc) * g.inv_u0);  // exp(-tau[0]/u0): 1 for tau[0] = 0
    s.TO = g.og_alias ? s.T0 : pbm::kexp(-__ldg(p.tau_og + ovc) * g.inv_u0);
    s.f1p = s.gam_p = s.cpd_p = s.cmd_p = s.e1p = s.e3p = s.s13p = s.s24p = 0.0;
    double e2L = 0.0, e4L = 0.0;  // e2, e4 of the last processed layer (surface row)
    produce(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) produce(c + 1);
        const double *buf = smem + (c & 1) * tile + cw;
        const int lbase = c * CH;
        const int nk = L - lbase < CH ? L - lbase : CH;
        {
        int k = 0;
        for (; k + 1 < nk; k += 2) {
            const double *q0 = buf + k * NR * 32, *q1 = q0 + NR * 32;
            // all exponentials of both layers first (independent), then both coefficient sets
            Refl4Exp x0, x1;
            refl4_exps(g, q0, x0);
            refl4_exps(g, q1, x1);
            Refl4Rec c0, c1;
            refl4_coeffs<MP>(p, g, q0, x0, s, c0);
            refl4_coeffs<MP>(p, g, q1, x1, s, c1);
            if (lbase + k == 0) refl4_first(c0, btop, s);
            else refl4_step(c0, s);
            refl4_carry(c0, s);
            refl4_step(c1, s);
            refl4_carry(c1, s);
            e2L = c1.e2; e4L = c1.e4;
        }
        if (k < nk) {
            const double *q0 = buf + k * NR * 32;
            Refl4Exp x0;
            refl4_exps(g, q0, x0);
            Refl4Rec c0;
            refl4_coeffs<MP>(p, g, q0, x0, s, c0);
            if (lbase + k == 0) refl4_first(c0, btop, s);
            else refl4_step(c0, s);
            refl4_carry(c0, s);
            e2L = c0.e2; e4L = c0.e4;
        }
        }
        __syncthreads();
    }
    double result;
    {
        // I_L = flux_zero/pi (fluxes.py:1266-1270) enters with weight T_L; fold X[2L-2], then the
        // surface row 2L-1 (fluxes.py:178-181) closes the chain
        const double ipi = 1.0 / PB_PI;
        double P = s.P + s.T1 * (s.e1p * ipi);
        const double q1 = s.f1p + s.T1 * (e2L * ipi);
        double R = s.R + s.T1 * (s.cpd_p * ipi);
        R = R + P * s.DS;
        P = q1 - P * s.CS;
        const double b_surface = 0.0 + r * g.u0 * g.f0 * s.T0;
        const double A_ = s.e1p - r * s.e3p, B_ = e2L - r * e4L;
        const double D_ = b_surface - s.cpd_p + r * s.cmd_p;
        const double x = pbm::krcp(B_ - A_ * s.CS);
        result = R + P * ((D_ - A_ * s.DS) * x);
    }
    const bool active = (w < p.W) && (a < p.G) && (ca < AY);
    if (active && p.xint) p.xint[((int64_t)b * p.G + a) * p.W + w] = result;
    if (p.fuse_albedo) {
        if (p.g_n > 0 && p.g_wait && wy == 0 && lane == 0) {
            // hold the peer stores until every rank has published wait_step (buffer rotation, see pb_peer_gather)
            const unsigned long long *mine = p.g_flag[p.g_rank];
            const long long t0 = clock64();
            for (int rk = 0; rk < p.g_n; ++rk)
                while (ld_acquire_sys(mine + rk) < p.g_wait)
                    if (clock64() - t0 > kSpinLimit) { atomicExch(p.g_done + 1, 1u); break; }
        }
        // compress_disco (disco.py:138-149): sequential sum over (ig, it) in index order
        if (ca < AY) smem[ca * kWavesPerCta + cw] = result;
        __syncthreads();
        if (ca == 0 && w < p.W) {
            double acc = 0.0;
            for (int aa = 0; aa < p.G; ++aa) {
                const int ig = aa / p.nt, it = aa - ig * p.nt;
                acc = acc + smem[aa * kWavesPerCta + cw] * p.gweight[ig] * p.tweight[it];
            }
            const double sym = (p.nt == 1) ? 2.0 * PB_PI : 1.0;
            const double alb = sym * 0.5 * acc / g.f0 * (p.cos_theta + 1.0);
            p.albedo[(int64_t)b * p.W + w] = alb;
            // fused all-gather: this rank's slab goes to row g_rank of every rank's buffer (NVLink P2P stores)
            for (int rk = 0; rk < p.g_n; ++rk) p.g_alb[rk][(int64_t)p.g_rank * p.W + w] = alb;
        }
        if (p.g_n > 0) {
            // last CTA to finish publishes the step on every rank.  The CTA barrier orders the writers' peer
            // stores before thread 0; its gpu-scope fence + counter increment order them before the last
            // CTA's observation of the full count, and that CTA's (cumulative) system-scope fence orders
            // everything it has observed before the release stores of the flags (PTX causality order is
            // transitive over morally strong edges of different scopes).  A system-scope fence in every
            // CTA costs 8-10 us per launch (measured, self-gather on one GPU).
            __syncthreads();
            if (wy == 0 && lane == 0) {
                __threadfence();
                const unsigned total = gridDim.x * gridDim.y * gridDim.z;
                if (atomicAdd(p.g_done, 1u) == total - 1) {
                    atomicExch(p.g_done, 0u);
                    __threadfence_system();
                    for (int rk = 0; rk < p.g_n; ++rk) st_release_sys(p.g_flag[rk] + p.g_rank, p.g_step);
                }
            }
        }
    }
}

// 32-wide tiles: 5 angle warps per CTA, 3 CTAs per SM -> 128 registers.  SM-count-sized tiles: 4 warps per
// CTA, 3 CTAs per SM -> up to 168 registers (no spills).
#ifndef PB_REFL4_GEN_REGS
#define PB_REFL4_GEN_REGS 168
#endif
template <int MP, bool GEN>
__global__ void __maxnreg__(PB_REFL4_REGS) refl_toa_kernel4(ReflParams p)
{
    refl4_body<MP, GEN>(p);
}
template <int MP>
__global__ void __maxnreg__(PB_REFL4_GEN_REGS) refl_toa_kernel4_gen(ReflParams p)
{
    refl4_body<MP, true>(p);
}
