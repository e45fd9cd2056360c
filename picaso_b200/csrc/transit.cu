// transit.cu - transmission-spectrum chord integration for sm_100a.
//
// Replaces picaso/fluxes.py:2582-2663 (get_transit_1d, Brown 2001 eq. 11).
//
// Two kernels: a tiny one builds the chord matrix M[i][k] = 2 * delta_length[i, i-k-1]
// (the path through layer k seen from the tangent level i, already doubled for the two
// halves of the chord, fluxes.py:2624-2644, :2656) plus z*dz; the main kernel assigns
// one wavelength per thread, stages its sigma_k = DTAU_k / colden_k * mmw_k * amu column
// slice in shared memory (coalesced, read from HBM once) and evaluates the
// lower-triangular contraction tau_i = sum_{k<i} sigma_k M[i][k] four tangent levels at
// a time for ILP, M being broadcast through L1.
#include "pb_common.cuh"

namespace {

constexpr int kThreads = 64;

__global__ void transit_path_kernel(int V, const double *z, const double *dz, const double *player,
                                    const double *tlayer, double k_b, double *M, double *zdz)
{
    const int b = blockIdx.y;
    z += (int64_t)b * V; dz += (int64_t)b * V; player += (int64_t)b * V; tlayer += (int64_t)b * V;
    M += (int64_t)b * V * V; zdz += (int64_t)b * V;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < V) zdz[idx] = z[idx] * dz[idx];
    if (idx >= V * V) return;
    const int i = idx / V, k = idx - i * V;
    double m = 0.0;
    if (k < i) {
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
    M[idx] = m;
}

__global__ void __launch_bounds__(kThreads) transit_kernel(int V, int W, int64_t ld, int64_t bs_layer,
                                                         const double *DTAU, const double *scale /*[B][L]*/,
                                                         const double *M, const double *zdz,
                                                         const double *zmin, double rstar, double *F)
{
    extern __shared__ double s_sig[];  // [L][kThreads]
    const int L = V - 1;
    const int b = blockIdx.y;
    const int w = blockIdx.x * kThreads + threadIdx.x;
    const bool active = w < W;
    const double *Mb = M + (int64_t)b * V * V;
    const double *zb = zdz + (int64_t)b * V;
    const double *sc = scale + (int64_t)b * L;
    if (active) {
        const double *col = DTAU + (int64_t)b * bs_layer + w;
        for (int k = 0; k < L; ++k) s_sig[k * kThreads + threadIdx.x] = col[(int64_t)k * ld] * sc[k];
    }
    // each thread only reads back its own column: no barrier needed
    if (!active) return;
    double acc = 0.0;
    int i = 1;
    for (; i + 3 < V; i += 4) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        const double *m0 = Mb + (int64_t)i * V, *m1 = m0 + V, *m2 = m1 + V, *m3 = m2 + V;
        for (int k = 0; k < i; ++k) {
            const double s = s_sig[k * kThreads + threadIdx.x];
            t0 = fma(s, __ldg(m0 + k), t0);
            t1 = fma(s, __ldg(m1 + k), t1);
            t2 = fma(s, __ldg(m2 + k), t2);
            t3 = fma(s, __ldg(m3 + k), t3);
        }
        // remaining triangle entries of the 4-row block
        {
            const double s0 = s_sig[i * kThreads + threadIdx.x];
            t1 = fma(s0, __ldg(m1 + i), t1);
            t2 = fma(s0, __ldg(m2 + i), t2);
            t3 = fma(s0, __ldg(m3 + i), t3);
            const double s1 = s_sig[(i + 1) * kThreads + threadIdx.x];
            t2 = fma(s1, __ldg(m2 + i + 1), t2);
            t3 = fma(s1, __ldg(m3 + i + 1), t3);
            const double s2 = s_sig[(i + 2) * kThreads + threadIdx.x];
            t3 = fma(s2, __ldg(m3 + i + 2), t3);
        }
        acc += (1.0 - exp(-t0)) * zb[i] + (1.0 - exp(-t1)) * zb[i + 1] +
               (1.0 - exp(-t2)) * zb[i + 2] + (1.0 - exp(-t3)) * zb[i + 3];
    }
    for (; i < V; ++i) {
        double t0 = 0.0;
        const double *m0 = Mb + (int64_t)i * V;
        for (int k = 0; k < i; ++k) t0 = fma(s_sig[k * kThreads + threadIdx.x], __ldg(m0 + k), t0);
        acc += (1.0 - exp(-t0)) * zb[i];
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
    const size_t smem = (size_t)L * kThreads * sizeof(double);
    if (smem > 200 * 1024) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "transit: nlevel=%d exceeds the shared-memory tile (max 400 layers)", V);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const size_t nW = (size_t)W * sizeof(double);
    size_t need = 16 * 256 + 6 * pb_align((size_t)B * V * 8) + pb_align((size_t)B * V * V * 8) + pb_align((size_t)B * 8);
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
        for (int k = 0; k < L; ++k) scale[(size_t)b * L + k] = 1.0 / a->colden[(size_t)b * L + k] * (a->mmw[(size_t)b * L + k] * a->amu);
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
    double *d_M, *d_zdz, *d_F = a->F;
    PB_TRY(pb_arena_alloc(ctx, (size_t)B * V * V * 8, (void **)&d_M));
    PB_TRY(pb_arena_alloc(ctx, (size_t)B * V * 8, (void **)&d_zdz));
    if (host) PB_TRY(pb_arena_alloc(ctx, B * nW, (void **)&d_F));

    dim3 gp((V * V + 127) / 128, B);
    transit_path_kernel<<<gp, 128, 0, ctx->stream>>>(V, d_z, d_dz, d_p, d_t, a->k_b, d_M, d_zdz);
    PB_CHECK_LAUNCH(ctx);
    if (smem > 48 * 1024)
        PB_CUDA(ctx, cudaFuncSetAttribute(transit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + kThreads - 1) / kThreads, B);
    transit_kernel<<<grid, kThreads, smem, ctx->stream>>>(V, W, ld, (int64_t)L * ld, d_dtau, d_scale, d_M, d_zdz, d_zmin, a->rstar, d_F);
    PB_CHECK_LAUNCH(ctx);
    if (host) {
        PB_CUDA(ctx, cudaMemcpyAsync(a->F, d_F, B * nW, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
