// transit.cu - transmission-spectrum chord integration for sm_100a.
//
// Replaces picaso/fluxes.py:2582-2663 (get_transit_1d, Brown 2001 eq. 11).
//
// Two kernels.  A tiny one builds the TRANSPOSED chord matrix MT[k][i] = 2 * delta_length[i, i-k-1]
// (path through layer k along the chord tangent at level i, doubled for the two halves,
// fluxes.py:2624-2644, :2656; zero for k >= i) plus z*dz.  The main kernel assigns one
// wavelength per thread and evaluates the lower-triangular contraction
//     tau_i = sum_{k<i} sigma_k MT[k][i],   sigma_k = DTAU_k / colden_k * mmw_k * amu
// 32 tangent levels at a time in registers: per k one coalesced sigma load and shared-memory broadcasts of
// MT[k][i0..i0+31] feed 32 DFMAs; the sigma loads of the next four layers are in flight while the current four
// are folded in (software pipeline: ncu showed the v1 loop waiting on its own loads, long_scoreboard 4.0 of
// 6.2 stall cycles per issue at 12 warps per SM).  The reference is compiled with fastmath, so the summation
// order is free.
#include "pb_common.cuh"

namespace {

constexpr int kThreads = 128;
// A/B on B200, 80 x 50 000 (profiles/r2_ab_transit.log): (levels per pass, loads per trip) without the software pipeline
// (16,4) 80.2 us, (32,4) 80.1, (48,4) 71.8, (96,4) 114.8 (255 registers: two residency waves); with it (PB_TRANSIT_PF)
// (16,4) 61.8, (32,4) 57.8, (32,8) 57.6, (48,4) 59.8
#ifndef PB_TRANSIT_BLK
#define PB_TRANSIT_BLK 32
#endif
#ifndef PB_TRANSIT_UNROLL
#define PB_TRANSIT_UNROLL 4
#endif
#ifndef PB_TRANSIT_PF
#define PB_TRANSIT_PF 1
#endif
constexpr int kBlk = PB_TRANSIT_BLK;  // tangent levels per pass (accumulators per thread)
constexpr int kUnroll = PB_TRANSIT_UNROLL;

__global__ void transit_path_kernel(int V, int Vp, const double *z, const double *dz, const double *player,
                                    const double *tlayer, double k_b, double *MT, double *zdz)
{
    const int b = blockIdx.y;
    z += (int64_t)b * V; dz += (int64_t)b * V; player += (int64_t)b * V; tlayer += (int64_t)b * V;
    MT += (int64_t)b * V * Vp; zdz += (int64_t)b * Vp;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < Vp) zdz[idx] = idx < V ? z[idx] * dz[idx] : 0.0;
    if (idx >= V * Vp) return;
    const int k = idx / Vp, i = idx - k * Vp;
    double m = 0.0;
    if (i < V && k < i) {
        const int j = i - k - 1;
        const double ref = z[i], inner = z[i - j], outer = z[i - j - 1];
        double seg = 0.0;
        // fluxes.py:2636-2639 (the j == 0 case drops the vanishing inner root)
        if (inner != ref && outer != ref)
            seg = sqrt(outer * outer - ref * ref) - sqrt(inner * inner - ref * ref);
        else if (inner == ref)
            seg = sqrt(outer * outer - ref * ref);
        m = 2.0 * (seg * player[k] / tlayer[k] / k_b);
    }
    MT[idx] = m;
}

__global__ void __launch_bounds__(kThreads) transit_kernel(int V, int Vp, int W, int64_t ld, int64_t bs_layer,
                                                         const double *DTAU, const double *scale /*[B][L]*/,
                                                         const double *MT, const double *zdz,
                                                         const double *zmin, double rstar, double *F)
{
    extern __shared__ double s_mt[];  // [L][Vp], zdz[Vp], scale[L]
    const int L = V - 1;
    const int b = blockIdx.y;
    const double *mt = MT + (int64_t)b * V * Vp;
    for (int i = threadIdx.x; i < L * Vp; i += kThreads) s_mt[i] = mt[i];
    double *s_zdz = s_mt + L * Vp;
    for (int i = threadIdx.x; i < Vp; i += kThreads) s_zdz[i] = zdz[(int64_t)b * Vp + i];
    double *s_sc = s_zdz + Vp;
    for (int i = threadIdx.x; i < L; i += kThreads) s_sc[i] = scale[(int64_t)b * L + i];
    __syncthreads();
    const int w = blockIdx.x * kThreads + threadIdx.x;
    if (w >= W) return;
    const double *col = DTAU + (int64_t)b * bs_layer + w;
    double acc = 0.0;
    for (int i0 = 0; i0 < V; i0 += kBlk) {
        double t[kBlk];
#pragma unroll
        for (int u = 0; u < kBlk; ++u) t[u] = 0.0;
        const int kend = (i0 + kBlk - 1 < L) ? i0 + kBlk - 1 : L;  // rows k < i <= i0 + kBlk - 1
#if PB_TRANSIT_PF
        // software pipeline: the raw sigma values of the NEXT kUnroll layers are requested before the current kUnroll
        // layers are folded in, and nothing touches them until the next trip (the in-order issue of a warp stalls at
        // the first instruction that needs a load, so the scale multiply waits until the value is consumed)
        double nxt[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) nxt[u] = u < kend ? __ldg(col + (int64_t)u * ld) : 0.0;
        for (int k = 0; k < kend; k += kUnroll) {
            double cur[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) cur[u] = nxt[u];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int kn = k + kUnroll + u;
                nxt[u] = kn < kend ? __ldg(col + (int64_t)kn * ld) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                if (k + u < kend) {
                    const double s = cur[u] * s_sc[k + u];
                    const double *m = s_mt + (k + u) * Vp + i0;
#pragma unroll
                    for (int v = 0; v < kBlk; ++v) t[v] = fma(s, m[v], t[v]);
                }
            }
        }
#else
        // four sigma loads in flight per trip: the loop is otherwise one load -> kBlk dependent-free DFMAs
#pragma unroll kUnroll
        for (int k = 0; k < kend; ++k) {
            const double s = __ldg(col + (int64_t)k * ld) * s_sc[k];
            const double *m = s_mt + k * Vp + i0;
#pragma unroll
            for (int u = 0; u < kBlk; ++u) t[u] = fma(s, m[u], t[u]);
        }
#endif
#pragma unroll
        for (int u = 0; u < kBlk; ++u) acc += (1.0 - exp(-t[u])) * s_zdz[i0 + u];  // padded levels carry zdz = 0
    }
    const double q = zmin[b] / rstar;
    F[(int64_t)b * W + w] = q * q + 2.0 / (rstar * rstar) * acc;
}

} // namespace

extern "C" int pb_transit_1d(pb_ctx *ctx, const pb_transit_args *a, int memspace)
{
    if (!ctx || !a) return PB_ERR_ARG;
    const int V = a->nlevel, L = V - 1, W = a->nwno;
    const int B = a->nbatch > 0 ? a->nbatch : 1;
    if (V < 2 || W < 0) return pb_fail(ctx, PB_ERR_ARG, "transit: bad sizes nlevel=%d nwno=%d", V, W);
    if (W == 0) return PB_OK;
    if (a->ld < W) return pb_fail(ctx, PB_ERR_ARG, "transit: ld < nwno");
    if (!a->DTAU || !a->z || !a->dz || !a->player || !a->tlayer || !a->mmw || !a->colden || !a->F)
        return pb_fail(ctx, PB_ERR_ARG, "transit: NULL argument");
    const int Vp = (V + kBlk - 1) / kBlk * kBlk;
    const size_t smem = ((size_t)L * Vp + Vp + L) * sizeof(double);
    if (smem > 220 * 1024) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "transit: nlevel=%d exceeds the shared-memory chord matrix (max ~165 levels)", V);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const size_t nW = (size_t)W * sizeof(double);
    size_t need = 16 * 256 + 6 * pb_align((size_t)B * V * 8) + pb_align((size_t)B * V * Vp * 8) +
                  pb_align((size_t)B * Vp * 8) + pb_align((size_t)B * 8);
    if (host) need += pb_align((size_t)B * L * nW) + pb_align(B * nW);
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, need));
    PB_TRY(pb_pinned_reserve(ctx, (6 * (size_t)B * V + (size_t)B + 64) * sizeof(double)));

    const double *d_dtau;
    int64_t ldo;
    PB_TRY(pb_stage_in(ctx, a->DTAU, memspace, (int64_t)B * L, W, a->ld, &d_dtau, &ldo));
    const int64_t ld = host ? W : a->ld;
    // per-layer scale  mmw*amu/colden  (fluxes.py:2622, :2650) and min(z) on the host: O(nlevel)
    std::vector<double> scale((size_t)B * L), zmin(B);
    for (int b = 0; b < B; ++b) {
        for (int k = 0; k < L; ++k)
            scale[(size_t)b * L + k] = 1.0 / a->colden[(size_t)b * L + k] * (a->mmw[(size_t)b * L + k] * a->amu);
        double m = a->z[(size_t)b * V];
        for (int i = 1; i < V; ++i) m = a->z[(size_t)b * V + i] < m ? a->z[(size_t)b * V + i] : m;
        zmin[b] = m;
    }
    const double *d_z, *d_dz, *d_p, *d_t, *d_scale, *d_zmin;
    PB_TRY(pb_upload_small(ctx, a->z, (size_t)B * V, &d_z));
    PB_TRY(pb_upload_small(ctx, a->dz, (size_t)B * V, &d_dz));
    PB_TRY(pb_upload_small(ctx, a->player, (size_t)B * V, &d_p));
    PB_TRY(pb_upload_small(ctx, a->tlayer, (size_t)B * V, &d_t));
    PB_TRY(pb_upload_small(ctx, scale.data(), (size_t)B * L, &d_scale));
    PB_TRY(pb_upload_small(ctx, zmin.data(), (size_t)B, &d_zmin));
    double *d_MT, *d_zdz, *d_F = a->F;
    PB_TRY(pb_arena_alloc(ctx, (size_t)B * V * Vp * 8, (void **)&d_MT));
    PB_TRY(pb_arena_alloc(ctx, (size_t)B * Vp * 8, (void **)&d_zdz));
    if (host) PB_TRY(pb_arena_alloc(ctx, B * nW, (void **)&d_F));

    PB_TRY(pb_upload_flush(ctx));
    dim3 gp((V * Vp + 127) / 128, B);
    transit_path_kernel<<<gp, 128, 0, ctx->stream>>>(V, Vp, d_z, d_dz, d_p, d_t, a->k_b, d_MT, d_zdz);
    PB_CHECK_LAUNCH(ctx);
    if (smem > 48 * 1024)
        PB_CUDA(ctx, pb_ensure_smem(ctx, transit_kernel, smem));
    dim3 grid((W + kThreads - 1) / kThreads, B);
    transit_kernel<<<grid, kThreads, smem, ctx->stream>>>(V, Vp, W, ld, (int64_t)L * ld, d_dtau, d_scale, d_MT, d_zdz,
                                                          d_zmin, a->rstar, d_F);
    PB_CHECK_LAUNCH(ctx);
    if (host) {
        PB_CUDA(ctx, cudaMemcpyAsync(a->F, d_F, B * nW, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
