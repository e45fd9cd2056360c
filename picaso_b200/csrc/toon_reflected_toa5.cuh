// toon_reflected_toa5.cuh - fifth generation of the reflected TOA kernel (included by toon_reflected.cu
// inside its anonymous namespace, after toon_reflected_toa4.cuh whose helpers it reuses).
//
// ncu on v4 (profiles/r1_refl_toa_v4gen.summary.json): fp64 pipe 45 % busy, 146 fp64 instructions per consumed
// (layer, angle), every one of the G angle threads of a wavelength re-deriving quantities that do not depend
// on the angle.  v5 takes the angle out of everything that can lose it (get_reflected_1d, fluxes.py:1132-1208):
//
//  * THE PIVOT CHAIN IS SHARED.  In setup_tri_diag (fluxes.py:89-183) only the right-hand side D depends on
//    the incidence angle; A, B, C do not, so the pivots of the top-down elimination (the two reciprocals per
//    layer, CS of both interface rows) are the same for all angles of a wavelength.  One extra warp per CTA
//    (the "chain warp", lane = wavelength) walks them once, one chunk ahead of the consumers, and publishes
//    7 numbers per layer; an angle thread is left with the two affine updates of DS and the four FMAs of the
//    carried functional R + P X[n].
//  * THE TOON SOURCE COEFFICIENTS ARE A SUM AND A DIFFERENCE.  With g3 = 1/2 - c g, c = kappa u0
//    (kappa = sqrt(3)/2 quadrature | 3/4 Eddington) the numerators of a-/a+ (fluxes.py:1161-1166) are
//        g4 (g1 + 1/u0) + g2 g3 = B0 + v ,  g3 (g1 - 1/u0) + g2 g4 = B0 - v ,
//        B0 = (g1 + g2)/2 + kappa g ,  v = c g (g1 - g2) + 1/(2 u0) ,
//    so the producer ships B0 and g (g1 - g2) and the angle thread needs three instructions for both.
//    Likewise the multiple-scattering weights (fluxes.py:1275-1296) are h +- hv combinations of
//    h = omega/2pi (1 + gcos2 t2(u1)) and hv = 1.5 g omega/2pi u1.
//  * At zero phase (u0 == u1, the BASELINE geometry; template SAME) 1/(lambda^2 u1^2 - 1) is the reciprocal
//    of lambda^2 - 1/u0^2 that a+- already paid for, times 1/u0^2: one reciprocal per (layer, angle).
//  * exp() is a 64-entry table (2^(j/64), 16 interleaved copies in shared memory = conflict-free for any
//    lane pattern) times a degree-5 polynomial: 9 fp64 instructions, no branch (libdevice: ~21 with the
//    slow path).  Relative error <= 1.2e-16 |x| + 1 ulp.
//  * the inputs of the layer a warp will produce in the NEXT chunk are loaded into registers before the
//    consume phase, so no warp ever waits on its own global loads.
//
// Per consumed (layer, angle): 82 fp64 + 22 shared-memory loads (v4: 146 + 14); producer + chain add ~165 fp64
// per (layer, wavelength).  The caller's tau / tau_og arrays are honoured exactly as in v4: producers flag
// any level where tau[l+1] != tau[l] + dtau[l] (8 ulp) in a CTA-wide word and consumers then re-check per
// layer from global memory and re-base their running products with a true exponential.

#ifndef PB_REFL5_UNROLL
#define PB_REFL5_UNROLL 2
#endif
#ifndef PB_REFL5_PREFETCH
#define PB_REFL5_PREFETCH 1
#endif
enum { P5_DT = 0, P5_DTO, P5_LAM, P5_B0, P5_A1, P5_FW, P5_OMC, P5_OGC, P5_OG, P5_GAM, P5_EP, P5_EM, P5_S0, NP5 };
enum { C5_QO = 0, C5_S13, C5_CSO, C5_R1, C5_R2, C5_R3, C5_CS, NC5 };
constexpr int kRefl5Unroll = PB_REFL5_UNROLL;

struct Refl5Angle {  // per-thread constants of one viewing geometry
    double u0, u1, inv_u0, inv_u1, c, hm, t2c, wgt, f0;
    bool og_alias;
};

__device__ __forceinline__ void refl5_produce(const ReflParams &p, const Refl4Inputs &x, double f0, double kappa,
                                              const double *tab, double *q /* [NP5][32] column of this lane */,
                                              int *bad)
{
    const double c2pi = 0.5 / PB_PI;
    const double g = x.fc * x.cb;
    double g1, g2;
    toon_g(p.tc, x.om, g, g1, g2);
    const double lam = sqrt(g1 * g1 - g2 * g2);
    const double gam = (g1 - lam) * pbm::krcp(g2);
    const double E = fmin(lam * x.dt, p.clip);  // slice_gt(exptrm, 35 | 40), fluxes.py:1174, :516
    const double EP = pbm::exp_tab(E, tab);
    const double ps = p_single(p, x.cbo, x.gc2, x.fc, x.fr);
    const double omc = x.om * c2pi;
    q[P5_DT * 32] = x.dt;
    q[P5_DTO * 32] = x.dto;
    q[P5_LAM * 32] = lam;
    q[P5_B0 * 32] = fma(kappa, g, 0.5 * (g1 + g2));
    q[P5_A1 * 32] = g * (g1 - g2);
    q[P5_FW * 32] = f0 * x.om;
    q[P5_OMC * 32] = omc;
    q[P5_OGC * 32] = omc * x.gc2;
    q[P5_OG * 32] = 1.5 * g * omc;
    q[P5_GAM * 32] = gam;
    q[P5_EP * 32] = EP;
    q[P5_EM * 32] = pbm::krcp(EP);
    q[P5_S0 * 32] = (x.omo * f0 / (4.0 * PB_PI)) * ps;
    const bool ok = __double2hiint(refl4_tau_slot(x.tau0, x.dt, x.tau1)) == kConsistentHi &&
                    __double2hiint(refl4_tau_slot(x.tauo0, x.dto, x.tauo1)) == kConsistentHi;
    if (!ok) *bad = 1;
}

// The caller's tau / tau_og are not cumsum(dtau) somewhere in this CTA's tile: re-check level l + 1 of one column and
// return exp(-tau[l+1]/u0), exp(-tau_og[l+1]/u0) where the running product must be re-based (marker NaN otherwise).
// Out of line: it is never taken for compute_opacity's arrays and must not cost the main loop registers.
__device__ __noinline__ double2 refl5_exact_tau(const double *tau, const double *tau_og, const double *dtau_og,
                                                int64_t iv, int64_t il, int64_t ld, double dt, double inv_u0,
                                                bool og_alias)
{
    double2 out;
    const double td = refl4_tau_slot(__ldg(tau + iv), dt, __ldg(tau + iv + ld));
    out.x = __double2hiint(td) != kConsistentHi ? ::exp(-td * inv_u0) : td;
    if (og_alias) {
        out.y = out.x;
    } else {
        const double sd = refl4_tau_slot(__ldg(tau_og + iv), __ldg(dtau_og + il), __ldg(tau_og + iv + ld));
        out.y = __double2hiint(sd) != kConsistentHi ? ::exp(-sd * inv_u0) : sd;
    }
    return out;
}

// SAME: every angle of the launch has u0 == u1 (zero phase).
template <int MP /*multi_phase*/, bool SAME>
__device__ __forceinline__ void refl5_body(const ReflParams &p)
{
    extern __shared__ double smem[];
    // push = 3: the first g_nc CTAs of the x axis (dispatched first) are the previous step's couriers, the tiles follow
    if ((int)blockIdx.x < p.g_nc) {
        if (blockIdx.y == 0 && blockIdx.z == 0) peer_deferred_push(p, (int)blockIdx.x);
        return;
    }
    const int bx = (int)blockIdx.x - p.g_nc;
    // layout: exp table [64][16] | P tiles [3][CH][NP5][32] | C tiles [2][CH][NC5][32] | flag
    const int tid = threadIdx.x;
    const int lane = tid & 31, wy = tid >> 5;
    const int NW = (int)(blockDim.x >> 5);   // the last warp is the chain warp
    const int NWC = NW - 1;
    const int CH = p.ch;                     // layers per chunk = producing warps (wy < CH), CH <= NW
    const int WT = p.wt, AY = p.ay;
    const bool is_chain = wy == NWC;
    const bool is_cons = tid < WT * AY;
    const int cw = is_cons ? tid % WT : 0, ca = is_cons ? tid / WT : 0;  // consumer identity
    const int w = bx * WT + cw;
    const int wc = w < p.W ? w : p.W - 1;
    const int wp = bx * WT + (lane < WT ? lane : WT - 1);        // producer / chain column
    const int wpc = wp < p.W ? wp : p.W - 1;
    const int a = blockIdx.y * AY + ca;
    const int ac = a < p.G ? a : p.G - 1;
    const int b = blockIdx.z;
    const int L = p.L;
    const int64_t ld = p.ld;
    const int64_t ol = (int64_t)b * p.bs_layer + wpc;   // producer column
    const int64_t ov = (int64_t)b * p.bs_level + wpc;
    const int64_t olc = (int64_t)b * p.bs_layer + wc;   // consumer column
    const int64_t ovc = (int64_t)b * p.bs_level + wc;
    const int64_t ow = (int64_t)b * p.bs_wave + wc;
    double *tabw = smem;
    double *ptile = smem + pbm::kExpTabDoubles;
    const int psz = CH * NP5 * 32, csz = CH * NC5 * 32;
    double *ctile = ptile + 3 * psz;
    int *bad = (int *)(ctile + 2 * csz);
    pbm::exp_tab_fill(tabw, tid, blockDim.x);
    if (tid == 0) *bad = 0;
    const double *tab = tabw + (lane & 15);
    const double f0p = p.f0pi ? p.f0pi[(int64_t)b * p.bs_wave + wpc] : 1.0;
    const double kappa = p.tc == 1 ? 0.75 : 0.8660254037844386;
    Refl5Angle g;
    g.u0 = p.variant ? fabs(p.ubar0[b]) : p.ubar0[ac];  // 3-D facets: geometry per batch entry, |ubar|
    g.u1 = p.variant ? fabs(p.ubar1[b]) : p.ubar1[ac];
    g.f0 = p.f0pi ? p.f0pi[ow] : 1.0;
    const double r = p.surf ? p.surf[ow] : 0.0;
    const double btop = p.btop ? p.btop[ow] : 0.0;
    if (SAME) g.u1 = g.u0;
    g.inv_u0 = 1.0 / g.u0; g.inv_u1 = SAME ? g.inv_u0 : 1.0 / g.u1;
    g.c = kappa * g.u0;
    g.hm = 0.5 * g.inv_u0;
    g.wgt = g.u0 / (g.u0 + g.u1);
    const double ubar2 = 0.767;  // fluxes.py:1280
    g.t2c = (3.0 * ubar2 * ubar2 * g.u1 * g.u1 - 1.0) / 2.0;
    g.og_alias = (p.dtau_og == p.dtau) && (p.tau_og == p.tau);
    // opaque to the optimiser: otherwise ptxas re-derives these per-angle constants inside the layer loop
    // (6 fp64 instructions per layer) instead of keeping three registers
    asm volatile("" : "+d"(g.c), "+d"(g.hm), "+d"(g.t2c));
    const int nchunks = (L + CH - 1) / CH;

    // Inputs of the layer this warp produces next.  PB_REFL5_PREFETCH 0: loaded into registers one iteration ahead
    // (26 registers live across the consume phase); 1: pulled into L2 one iteration ahead, loaded when produced.
    Refl4Inputs pre;
    auto prefetch = [&](int c) {
        const int l = c * CH + wy;
        if (wy < CH && l < L) {
#if PB_REFL5_PREFETCH == 0
            refl4_load(p, ol + (int64_t)l * ld, ov + (int64_t)l * ld, ld, pre);
#else
            const int64_t il = ol + (int64_t)l * ld, iv = ov + (int64_t)l * ld;
            if (lane == 0 || lane == 16) {  // one prefetch per 128-byte line of the 256-byte row segment
                const double *rows[12] = {p.w0 + il, p.fcld + il, p.cosb + il, p.dtau + il, p.gcos2 + il, p.fray + il,
                                          p.dtau_og + il, p.w0_og + il, p.cosb_og + il, p.tau + iv + ld, p.tau_og + iv + ld,
                                          p.tau + iv};
#pragma unroll
                for (int i = 0; i < 12; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(rows[i]));
            }
#endif
        }
    };
    auto produce = [&](int c) {
        const int l = c * CH + wy;
        if (wy < CH && l < L) {
#if PB_REFL5_PREFETCH != 0
            refl4_load(p, ol + (int64_t)l * ld, ov + (int64_t)l * ld, ld, pre);
#endif
            refl5_produce(p, pre, f0p, kappa, tab, ptile + (c % 3) * psz + wy * NP5 * 32 + lane, bad);
        }
    };

    // chain-warp state: relation X[2l] = DS - CS X[2l+1] of the last even row, previous layer's e-terms
    double ch_CS = 0.0, ch_e1p = 0.0, ch_e3p = 0.0, ch_s13p = 0.0, ch_s24p = 0.0, ch_gamp = 0.0;
    auto chain = [&](int c) {
        const double *pt = ptile + (c % 3) * psz + lane;
        double *ct = ctile + (c & 1) * csz + lane;
        const int lbase = c * CH;
        const int nk = L - lbase < CH ? L - lbase : CH;
        for (int k = 0; k < nk; ++k) {
            const double *q = pt + k * NP5 * 32;
            double *o = ct + k * NC5 * 32;
            const double gam = q[P5_GAM * 32], EP = q[P5_EP * 32], EM = q[P5_EM * 32];
            if (lbase + k == 0) {
                // row 0 (fluxes.py:155-158)
                const double x = pbm::krcp(gam + 1.0);
                ch_CS = (gam - 1.0) * x;
                o[C5_QO * 32] = x;
                o[C5_CS * 32] = ch_CS;
            } else {
                // interface rows 2l-1, 2l (fluxes.py:161-175)
                const double gm1 = gam - 1.0;
                const double A1 = ch_s13p * gm1;  // A of the odd row, C of the even row
                const double xo = pbm::krcp(ch_s24p * gm1 - A1 * ch_CS);
                const double CSo = (2.0 * (1.0 - gam * gam)) * xo;
                const double A2 = 2.0 * (1.0 - ch_gamp * ch_gamp);
                const double xe = pbm::krcp((ch_e1p - ch_e3p) * (gam + 1.0) - A2 * CSo);
                ch_CS = A1 * xe;
                o[C5_QO * 32] = gm1 * xo;
                o[C5_S13 * 32] = ch_s13p;
                o[C5_CSO * 32] = CSo;
                o[C5_R1 * 32] = ch_e3p * xe;
                o[C5_R2 * 32] = ch_e1p * xe;
                o[C5_R3 * 32] = A2 * xe;
                o[C5_CS * 32] = ch_CS;
            }
            const double e1 = EP + gam * EM, e2 = EP - gam * EM, e3 = gam * EP + EM, e4 = gam * EP - EM;
            ch_e1p = e1; ch_e3p = e3; ch_s13p = e1 + e3; ch_s24p = e2 + e4; ch_gamp = gam;
        }
    };

    // consumer state
    double DS = 0.0, P = 0.0, R = 0.0, CSp = 0.0;
    double T1 = 1.0, T0 = 1.0, TO = 1.0, f1p = 0.0, am_p = 0.0, ap_p = 0.0;
    __syncthreads();  // table
    if (is_cons) {
        T0 = pbm::exp_tab(-__ldg(p.tau + ovc) * g.inv_u0, tab);  // exp(-tau[0]/u0): 1 for tau[0] = 0
        TO = g.og_alias ? T0 : pbm::exp_tab(-__ldg(p.tau_og + ovc) * g.inv_u0, tab);
    }

    // SLOW (std::true_type): some level of this CTA's tile has tau[l+1] != tau[l] + dtau[l] - re-check every level
    auto consume = [&](int c, auto slow_tag) {
        constexpr bool slow = decltype(slow_tag)::value;
        const double *pt = ptile + (c % 3) * psz + cw;
        const double *ct = ctile + (c & 1) * csz + cw;
        const int lbase = c * CH;
        const int nk = L - lbase < CH ? L - lbase : CH;
        const double *q = pt, *o = ct;
#pragma unroll kRefl5Unroll
        for (int k = 0; k < nk; ++k, q += NP5 * 32, o += NC5 * 32) {
            const double dt = q[P5_DT * 32];
            const double xa = pbm::exp_tab(-dt * g.inv_u1, tab);
            const double xa0 = SAME ? xa : pbm::exp_tab(-dt * g.inv_u0, tab);
            double xoa0, xoa1;
            if (g.og_alias) {
                xoa0 = xa0; xoa1 = xa;
            } else {
                const double dto = q[P5_DTO * 32];
                xoa0 = pbm::exp_tab(-dto * g.inv_u0, tab);
                xoa1 = SAME ? xoa0 : pbm::exp_tab(-dto * g.inv_u1, tab);
            }
            const double xu = T0, xo = TO;
            double xd = xu * xa0;
            double xod = g.og_alias ? xd : xo * xoa0;
            if (slow) {
                const int l = lbase + k;
                const double2 ex = refl5_exact_tau(p.tau, p.tau_og, p.dtau_og, ovc + (int64_t)l * ld, olc + (int64_t)l * ld,
                                                   ld, dt, g.inv_u0, g.og_alias);
                if (__double2hiint(ex.x) != kConsistentHi) xd = ex.x;
                if (g.og_alias) xod = xd;
                else if (__double2hiint(ex.y) != kConsistentHi) xod = ex.y;
            }
            // a-, a+ (fluxes.py:1161-1166) as B0 +- v.  lambda^2 - 1/u0^2 is formed as (lambda - 1/u0)(lambda + 1/u0):
            // near lambda u0 = 1 the rounded squares would lose ~9 digits, and at zero phase 1/(lambda u1 + 1) below is
            // derived from this reciprocal - it multiplies the LARGE homogeneous amplitude that cancels the resonant
            // particular solution, so it must carry full relative accuracy (the reference divides by lambda u1 + 1
            // directly); measured: 9e-6 relative error at 5 of 50 000 (wavelength, angle) columns without this.
            const double lam = q[P5_LAM * 32];
            const double lpl = lam + g.inv_u0, lmi = lam - g.inv_u0;
            const double inv_den = pbm::krcp(lpl * lmi);
            const double wq = q[P5_FW * 32] * inv_den;
            const double v = fma(g.c, q[P5_A1 * 32], g.hm);
            const double B0 = q[P5_B0 * 32];
            const double am = (B0 + v) * wq, ap = (B0 - v) * wq;
            const double dcp = (ap - ap_p) * xu;   // c+up_l - c+down_{l-1}
            const double dcm = (am_p - am) * xu;   // c-down_{l-1} - c-up_l
            // multiple scattering (fluxes.py:1275-1296): (mpl, mmi) omega/2pi = h +- hv
            const double h = MP == 0 ? fma(q[P5_OGC * 32], g.t2c, q[P5_OMC * 32]) : q[P5_OMC * 32];
            const double hv = q[P5_OG * 32] * g.u1;
            const double hp = h + hv, hmi = h - hv;
            const double gam = q[P5_GAM * 32], EP = q[P5_EP * 32], EM = q[P5_EM * 32];
            const double mG = fma(gam, hmi, hp), mH = fma(gam, hp, hmi);
            // T1 / (lam u1 - 1), T1 / (lam u1 + 1)
            double rl1, rl2;
            if (SAME) {
                const double imT = inv_den * (g.inv_u0 * T1);
                rl1 = lpl * imT;
                rl2 = lmi * imT;
            } else {
                const double lu = lam * g.u1;
                const double lu1 = lu + 1.0, lu2 = lu - 1.0;
                const double ilT = pbm::krcp(lu1 * lu2) * T1;
                rl1 = lu1 * ilT;
                rl2 = lu2 * ilT;
            }
            const double cG = mG * (fma(EP, xa, -1.0) * rl1);
            const double cH = mH * (fma(-EM, xa, 1.0) * rl2);
            const double f0 = cG + cH, f1 = cG - cH;   // weights of X[2l], X[2l+1] in I_0
            const double At = fma(hp, ap, hmi * am) * xu;
            const double xs = xa * xa0, xso = xoa0 * xoa1;
            const double K = (g.wgt * T1) * fma(q[P5_S0 * 32] * xo, 1.0 - xso, At * (1.0 - xs));
            if (lbase + k == 0) {
                DS = (btop - am * xu) * o[C5_QO * 32];
                P = f0;
                R = K;
            } else {
                const double DSo = o[C5_QO * 32] * ((dcp - dcm) - o[C5_S13 * 32] * DS);
                R = fma(P, DS, R);
                const double Pn = fma(-P, CSp, f1p);
                R = fma(Pn, DSo, R);
                P = fma(-Pn, o[C5_CSO * 32], f0);
                DS = fma(o[C5_R1 * 32], dcp, fma(o[C5_R2 * 32], dcm, -o[C5_R3 * 32] * DSo));
                R = R + K;
            }
            CSp = o[C5_CS * 32];
            f1p = f1; am_p = am; ap_p = ap;
            T1 = T1 * xa; T0 = xd; TO = xod;
        }
    };

    // Pipeline, one barrier per chunk: in iteration c the consumers integrate chunk c (P[c % 3], C[c & 1]), the chain warp
    // eliminates chunk c + 1 (P[(c+1) % 3] -> C[(c+1) & 1]) and every producing warp writes its layer of chunk c + 2
    // (P[(c+2) % 3]) from registers loaded one iteration earlier, then loads its layer of chunk c + 3.
    if (p.g_n > 0 && !p.g_defer && tid == 0 && bx == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        if (p.g_lazy && p.g_step > 1) {
            // lazy flags: the previous launch's peer stores are performed (its grid has retired): publish its step
            __threadfence_system();
            for (int rk = 0; rk < p.g_n; ++rk) st_release_sys(p.g_flag[rk] + p.g_rank, p.g_step - 1);
        }
        // buffer-rotation guard (pb_peer_gather.wait_step), ONCE per launch and under the layer sweep: this thread
        // polls the system-scope flags, every CTA's epilogue then needs a single gpu-scope load of the go word
        // (polling 8 system-scope flags in each of the 435 CTA tails cost ~15 us per step at 8 GPUs)
        if (p.g_wait) {
            const unsigned long long *mine = p.g_flag[p.g_rank];
            const long long t0 = clock64();
            for (int rk = 0; rk < p.g_n; ++rk)
                while (ld_acquire_sys(mine + rk) < p.g_wait)
                    if (clock64() - t0 > kSpinLimit) { atomicExch(p.g_done + 1, 1u); break; }
        }
        __threadfence();
        atomicExch(p.g_done + 4, (unsigned)p.g_step);
    }
    prefetch(0);
    produce(0);
    prefetch(1);
    __syncthreads();
    if (is_chain) chain(0);
    produce(1);
    prefetch(2);
    __syncthreads();
    // warp-specialised main loops (same barrier sequence in both): the chain warp's state and the consumers'
    // state never live in the same registers
    if (is_chain) {
        for (int c = 0; c < nchunks; ++c) {
            if (c + 1 < nchunks) chain(c + 1);
            if (c + 2 < nchunks) {
                produce(c + 2);
                prefetch(c + 3);
            }
            __syncthreads();
        }
    } else {
        for (int c = 0; c < nchunks; ++c) {
            if (c + 2 < nchunks) {
                produce(c + 2);
                prefetch(c + 3);
            }
            if (is_cons) {
                if (*(volatile int *)bad != 0) consume(c, std::true_type());
                else consume(c, std::false_type());
            }
            __syncthreads();
        }
    }
    double result = 0.0;
    if (is_cons) {
        // I_L = flux_zero/pi (fluxes.py:1266-1270) enters with weight T_L; fold X[2L-2], then the
        // surface row 2L-1 (fluxes.py:178-181) closes the chain
        const int kl = (L - 1) - (nchunks - 1) * CH;
        const double *q = ptile + ((nchunks - 1) % 3) * psz + kl * NP5 * 32 + cw;
        const double gam = q[P5_GAM * 32], EP = q[P5_EP * 32], EM = q[P5_EM * 32];
        const double e1 = EP + gam * EM, e2 = EP - gam * EM, e3 = gam * EP + EM, e4 = gam * EP - EM;
        const double cpd = ap_p * T0, cmd = am_p * T0;
        const double ipi = 1.0 / PB_PI;
        double Pq = P + T1 * (e1 * ipi);
        const double q1 = f1p + T1 * (e2 * ipi);
        double Rq = R + T1 * (cpd * ipi);
        Rq = Rq + Pq * DS;
        Pq = q1 - Pq * CSp;
        const double b_surface = 0.0 + r * g.u0 * g.f0 * T0;
        const double A_ = e1 - r * e3, B_ = e2 - r * e4;
        const double D_ = b_surface - cpd + r * cmd;
        const double x = pbm::krcp(B_ - A_ * CSp);
        result = Rq + Pq * ((D_ - A_ * DS) * x);
    }
    const bool active = is_cons && (w < p.W) && (a < p.G);
    if (active && p.xint) p.xint[((int64_t)b * p.G + a) * p.W + w] = result;
    if (p.fuse_albedo) {
        if (p.g_n > 0 && !p.g_defer && tid == 0) {
            // hold the peer stores until CTA 0 has seen every rank publish wait_step (go word = this launch's step)
            const long long t0 = clock64();
            while (*(volatile unsigned *)(p.g_done + 4) != (unsigned)p.g_step)
                if (clock64() - t0 > kSpinLimit) { atomicExch(p.g_done + 1, 1u); break; }
            __threadfence();
        }
        // compress_disco (disco.py:138-149): sequential sum over (ig, it) in index order
        double *red = ctile;  // the chain tiles are dead (the closing step above still reads ptile)
        if (is_cons) red[ca * kWavesPerCta + cw] = result;
        __syncthreads();
        if (is_cons && ca == 0 && w < p.W) {
            double acc = 0.0;
            for (int aa = 0; aa < p.G; ++aa) {
                const int ig = aa / p.nt, it = aa - ig * p.nt;
                acc = acc + red[aa * kWavesPerCta + cw] * p.gweight[ig] * p.tweight[it];
            }
            const double sym = (p.nt == 1) ? 2.0 * PB_PI : 1.0;
            const double alb = sym * 0.5 * acc / g.f0 * (p.cos_theta + 1.0);
            p.albedo[(int64_t)b * p.W + w] = alb;
            // fused all-gather: this rank's slab goes to row g_rank of every rank's buffer (NVLink P2P stores)
            if (p.g_defer) p.g_alb[p.g_rank][(int64_t)p.g_rank * p.W + w] = alb;   // local row only; pushed by the next launch
            else for (int rk = 0; rk < p.g_n; ++rk) p.g_alb[rk][(int64_t)p.g_rank * p.W + w] = alb;
        }
        if (p.g_n > 0 && !p.g_lazy && !p.g_defer) {
            // last CTA to finish publishes the step on every rank (ordering argument: toon_reflected_toa4.cuh)
            __syncthreads();
            if (tid == 0) {
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

#ifndef PB_REFL5_UNROLL
#define PB_REFL5_UNROLL 2
#endif
#ifndef PB_REFL5_REGS
#define PB_REFL5_REGS 128
#endif
template <int MP, bool SAME>
__global__ void __maxnreg__(PB_REFL5_REGS) refl_toa_kernel5(ReflParams p)
{
    refl5_body<MP, SAME>(p);
}
