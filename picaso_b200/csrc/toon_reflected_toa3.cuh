// toon_reflected_toa3.cuh - third generation of the reflected TOA kernel (included by
// toon_reflected.cu inside its anonymous namespace, after ReflParams / refl_load / refl_produce).
//
// Same algorithm as refl_toa_kernel (single-sweep adjoint recurrence, producer/consumer through
// shared memory) with two changes aimed at the limiter ncu showed for v2 (per-warp dependent
// fp64 chains, "stall_wait", fp64 pipe < 50 % busy at 10 warps/SM):
//  * layers are consumed in PAIRS: the angle-dependent coefficient sets of two layers are
//    evaluated back to back (8-10 independent exponentials / reciprocals in one basic block
//    for ptxas to interleave) before the two short, inherently serial elimination steps;
//  * each warp produces TWO layers per chunk (chunk = 2 * NW layers, half as many barriers) and
//    no prefetched inputs are held across the consume phase (fewer live registers).

struct ReflRec {  // what the elimination / adjoint step needs from one (layer, angle)
    double gam, cpu, cmu, cpd, cmd, e1, e2, e3, e4, cP, cQ, xa, K;
};

struct ReflAngle {  // per-thread constants of one viewing geometry
    double u0, u1, inv_u0, inv_u1, s01, wgt, t2c, f0;
    bool same_mu, og_alias;
};

template <int MP>
__device__ __forceinline__ void refl_coeffs(const ReflParams &p, const ReflAngle &g, const double *q,
                                            double xu, double xd, ReflRec &o)
{
    const double c2pi = 0.5 / PB_PI;
    const double gg = q[Q_G * 32], om = q[Q_OM * 32], g1 = q[Q_G1 * 32], g2 = q[Q_G2 * 32];
    const double lam = q[Q_LAM * 32], gam = q[Q_GAM * 32], EP = q[Q_EP * 32], EM = q[Q_EM * 32];
    const double dt = q[Q_DT * 32];
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
        const double t2 = q[Q_GC2 * 32] * g.t2c;
        mpl = 1.0 + 1.5 * gg * g.u1 + t2;
        mmi = 1.0 - 1.5 * gg * g.u1 + t2;
    } else {
        mpl = 1.0 + 1.5 * gg * g.u1;
        mmi = 1.0 - 1.5 * gg * g.u1;
    }
    // fluxes.py:1290-1296, :1395-1407; exp(+-E - dt/u1) = EP|EM * exp(-dt/u1)
    const double xa = pbm::kexp(-dt * g.inv_u1);
    const double lu = lam * g.u1;
    const double inv_l = pbm::krcp(lu * lu - 1.0);
    const double omc = om * c2pi;
    const double cG = (mpl + gam * mmi) * omc * ((EP * xa - 1.0) * ((lu + 1.0) * inv_l));
    const double cH = (gam * mpl + mmi) * omc * ((1.0 - EM * xa) * ((lu - 1.0) * inv_l));
    const double At = (mpl * o.cpu + mmi * o.cmu) * omc;
    const double xs = g.same_mu ? xa * xa : pbm::kexp(-dt * g.s01);
    double xo, xso;
    if (g.og_alias) {
        xo = xu;
        xso = xs;
    } else {
        xo = pbm::kexp(-q[Q_TAUO * 32] * g.inv_u0);
        xso = pbm::kexp(-q[Q_DTO * 32] * g.s01);
    }
    o.cP = cG + cH;
    o.cQ = cG - cH;
    o.xa = xa;
    o.K = q[Q_S0 * 32] * xo * (1.0 - xso) * g.wgt + At * (1.0 - xs) * g.wgt;
}

struct ReflState {
    double AS, DS, Pp, Rp, gam_n, cpu_n, cmu_n;
};

// interface rows between layer l and l+1 (fluxes.py:161-175) + adjoint fold
__device__ __forceinline__ void refl_step(const ReflRec &c, ReflState &s)
{
    const double gm1 = s.gam_n - 1.0;
    const double e13 = (c.e1 + c.e3) * gm1;
    double a_ = 2.0 * (1.0 - c.gam * c.gam);
    double b_ = (c.e1 - c.e3) * (s.gam_n + 1.0);
    double d_ = c.e3 * (s.cpu_n - c.cpd) + c.e1 * (c.cmd - s.cmu_n);
    double x = pbm::krcp(b_ - e13 * s.AS);
    const double ASe = a_ * x, DSe = (d_ - e13 * s.DS) * x;
    const double alpha = s.Rp + s.Pp * DSe;  // I_{l+1} = Rp + Pp X[2l+2], X[2l+2] = DSe - ASe X[2l+1]
    const double beta = -s.Pp * ASe;
    b_ = (c.e2 + c.e4) * gm1;
    const double c_ = 2.0 * (1.0 - s.gam_n * s.gam_n);
    d_ = gm1 * (s.cpu_n - c.cpd) - gm1 * (c.cmd - s.cmu_n);
    x = pbm::krcp(b_ - c_ * ASe);
    s.AS = e13 * x;
    s.DS = (d_ - c_ * DSe) * x;
    const double Q = c.xa * beta + c.cQ;
    const double R = c.xa * alpha + c.K;
    s.Pp = c.cP - Q * s.AS;  // eliminate X[2l+1] = DS - AS X[2l]
    s.Rp = R + Q * s.DS;
    s.gam_n = c.gam; s.cpu_n = c.cpu; s.cmu_n = c.cmu;
}

// surface row 2L-1 (fluxes.py:178-181) and I_L = flux_zero/pi (:1266-1270)
__device__ __forceinline__ void refl_first(const ReflRec &c, double r, double b_surface, ReflState &s)
{
    const double a_ = c.e1 - r * c.e3, b_ = c.e2 - r * c.e4;
    const double d_ = b_surface - c.cpd + r * c.cmd;
    const double ib = pbm::krcp(b_);
    s.AS = a_ * ib;
    s.DS = d_ * ib;
    const double P = c.xa * (c.e1 / PB_PI) + c.cP;
    const double Q = c.xa * (c.e2 / PB_PI) + c.cQ;
    const double R = c.xa * (c.cpd / PB_PI) + c.K;
    s.Pp = P - Q * s.AS;
    s.Rp = R + Q * s.DS;
    s.gam_n = c.gam; s.cpu_n = c.cpu; s.cmu_n = c.cmu;
}

template <int MP /*multi_phase*/>
__global__ void __launch_bounds__(256) refl_toa_kernel3(ReflParams p)
{
    extern __shared__ double smem[];  // [2][2*NW][NQ][32]
    const int lane = threadIdx.x, wy = threadIdx.y, NW = blockDim.y;
    const int CH = 2 * NW;  // layers per chunk
    const int w = blockIdx.x * kWavesPerCta + lane;
    const int wc = w < p.W ? w : p.W - 1;  // clamp: every lane takes part in the tile protocol
    const int a = blockIdx.y * NW + wy;
    const int ac = a < p.G ? a : p.G - 1;
    const int b = blockIdx.z;
    const int L = p.L;
    const int64_t ld = p.ld;
    const int64_t ol = (int64_t)b * p.bs_layer + wc;
    const int64_t ov = (int64_t)b * p.bs_level + wc;
    const int64_t ow = (int64_t)b * p.bs_wave + wc;
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
    const int tile = CH * NQ * 32;

    auto produce = [&](int c) {
        // this warp's two layers of chunk c: positions wy and wy + NW of the chunk (bottom-up)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pos = wy + h * NW;
            const int l = L - 1 - (c * CH + pos);
            if (l >= 0) {
                ReflInputs x;
                refl_load(p, ol + (int64_t)l * ld, ov + (int64_t)l * ld, x);
                refl_produce(p, x, g.f0, smem + (c & 1) * tile + pos * NQ * 32 + lane);
            }
        }
    };

    ReflState s = {0, 0, 0, 0, 0, 0, 0};
    double xd = pbm::kexp(-__ldg(p.tau + ov + (int64_t)L * ld) * g.inv_u0);  // exp(-tau[L]/u0)
    const double b_surface = 0.0 + r * g.u0 * g.f0 * xd;
    produce(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) produce(c + 1);
        const double *buf = smem + (c & 1) * tile + lane;
        const int lbase = L - 1 - c * CH;
        const int nk = lbase + 1 < CH ? lbase + 1 : CH;
        int k = 0;
        for (; k + 1 < nk; k += 2) {
            const double *q0 = buf + k * NQ * 32, *q1 = q0 + NQ * 32;
            // the four exponentials of the optical-depth scale first, then both coefficient sets
            const double xu0 = pbm::kexp(-q0[Q_TAU * 32] * g.inv_u0);
            const double xu1 = pbm::kexp(-q1[Q_TAU * 32] * g.inv_u0);
            ReflRec c0, c1;
            refl_coeffs<MP>(p, g, q0, xu0, xd, c0);
            refl_coeffs<MP>(p, g, q1, xu1, xu0, c1);
            if (lbase - k == L - 1) refl_first(c0, r, b_surface, s);
            else refl_step(c0, s);
            refl_step(c1, s);
            xd = xu1;
        }
        if (k < nk) {
            const double *q0 = buf + k * NQ * 32;
            const double xu0 = pbm::kexp(-q0[Q_TAU * 32] * g.inv_u0);
            ReflRec c0;
            refl_coeffs<MP>(p, g, q0, xu0, xd, c0);
            if (lbase - k == L - 1) refl_first(c0, r, b_surface, s);
            else refl_step(c0, s);
            xd = xu0;
        }
        __syncthreads();
    }
    double result;
    {
        // row 0 (fluxes.py:155-158): X[0] = DS[0]
        const double b_ = s.gam_n + 1.0, c_ = s.gam_n - 1.0, d_ = btop - s.cmu_n;
        const double x = pbm::krcp(b_ - c_ * s.AS);
        const double X0 = (d_ - c_ * s.DS) * x;
        result = s.Rp + s.Pp * X0;
    }
    const bool active = (w < p.W) && (a < p.G);
    if (active && p.xint) p.xint[((int64_t)b * p.G + a) * p.W + w] = result;
    if (p.fuse_albedo) {
        // compress_disco (disco.py:138-149): sequential sum over (ig, it) in index order
        smem[wy * kWavesPerCta + lane] = result;
        __syncthreads();
        if (wy == 0 && w < p.W) {
            double acc = 0.0;
            for (int aa = 0; aa < p.G; ++aa) {
                const int ig = aa / p.nt, it = aa - ig * p.nt;
                acc = acc + smem[aa * kWavesPerCta + lane] * p.gweight[ig] * p.tweight[it];
            }
            const double sym = (p.nt == 1) ? 2.0 * PB_PI : 1.0;
            p.albedo[(int64_t)b * p.W + w] = sym * 0.5 * acc / g.f0 * (p.cos_theta + 1.0);
        }
    }
}
