// ck_mix.cu - resort-rebin mixing of per-gas correlated-k tables for sm_100a.
//
// Replaces picaso/deq_chem.py:334-386 (mix_all_gases_gasesfly), :388-432
// (do_mixing_mono_gasesfly), :538-597 (mix_2_gases) and the bilinear (1/T, log10 P)
// interpolation of optics.RetrieveCKs.mix_my_opacities_gasesfly (optics.py:1164-1197).
//
// One warp owns one (layer, wavelength) bin and walks the four (P, T) table neighbours.
// For each neighbour the gases are folded in pairwise: the Nk*Nk (<= 64) random-overlap
// products k = (m1 k1[i] + m2 k2[j]) / (m1 + m2) live two per lane, are sorted by (k, flat
// index) - the index tie-break reproduces numpy's stable mergesort - with a 64-element
// bitonic network in the warp's shared-memory slice, the weights w_i w_j are prefix-summed
// with shuffles, and lanes 0..Nk-1 each resample log10 k at one Gauss point (np.interp
// semantics).  The four neighbours' ln k are combined in the reference's summation order and
// exp() * N_A is written as molecular_opa[layer][wave][gauss].
#include "pb_common.cuh"
#include <math_constants.h>

namespace {

constexpr int kWarps = 4;
constexpr int kMaxGas = 32;
constexpr int kElems = 64;

struct MixParams {
    int L, W, K, ngas, np, nt;
    const double *kappa[kMaxGas];  // ln kappa [np][nt][W][K] per gas
    const double *mixes;           // [ngas][L]
    const int *indices;            // [4][L] p_low, p_hi, t_low, t_hi
    const double *t_interp, *p_interp;
    const double *gauss_pts, *gauss_wts;
    double *molecular_opa;  // [L][W][K]
    double *ln_mixed;       // optional [L][W][K][4]
};

__device__ __forceinline__ bool key_less(double ka, int ia, double kb, int ib)
{
    return ka < kb || (ka == kb && ia < ib);
}

__global__ void __launch_bounds__(kWarps * 32) ck_mix_kernel(MixParams p)
{
    __shared__ double s_key[kWarps][kElems];
    __shared__ double s_x[kWarps][kElems];
    __shared__ unsigned char s_idx[kWarps][kElems];
    __shared__ double s_gp[8], s_gw[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < p.K) {
        s_gp[threadIdx.x] = p.gauss_pts[threadIdx.x];
        s_gw[threadIdx.x] = p.gauss_wts[threadIdx.x];
    }
    __syncthreads();
    const int64_t bin = (int64_t)blockIdx.x * kWarps + wid;
    if (bin >= (int64_t)p.L * p.W) return;
    const int il = (int)(bin / p.W), iw = (int)(bin - (int64_t)il * p.W);
    const int K = p.K, n = K * K;
    double *key = s_key[wid], *xs = s_x[wid];
    unsigned char *idx = s_idx[wid];
    const bool glane = lane < K;
    const double gp = glane ? s_gp[lane] : 0.0;
    const int e0 = 2 * lane, e1 = e0 + 1;
    const int i0 = e0 / K, j0 = e0 - i0 * K, i1 = e1 / K, j1 = e1 - i1 * K;

    double lnk[4];
#pragma unroll 1
    for (int ct = 0; ct < 4; ++ct) {
        const int p_ind = p.indices[(ct >> 1) * p.L + il];
        const int t_ind = p.indices[(2 + (ct & 1)) * p.L + il];
        const int64_t row = (((int64_t)p_ind * p.nt + t_ind) * p.W + iw) * K;
        double k1 = glane ? exp(p.kappa[0][row + lane]) : 0.0;
        double mix_t = p.mixes[il];
#pragma unroll 1
        for (int g = 1; g < p.ngas; ++g) {
            const double k2 = glane ? exp(p.kappa[g][row + lane]) : 0.0;
            const double m2 = p.mixes[(int64_t)g * p.L + il];
            const double mt = mix_t + m2;
            // deq_chem.py:575-578
            const double a0 = __shfl_sync(0xffffffffu, k1, i0 < K ? i0 : 0), b0 = __shfl_sync(0xffffffffu, k2, j0);
            const double a1 = __shfl_sync(0xffffffffu, k1, i1 < K ? i1 : 0), b1 = __shfl_sync(0xffffffffu, k2, j1);
            // two elements per lane, in registers: sorted position 2*lane + s after the network
            double q0 = e0 < n ? (mix_t * a0 + m2 * b0) / mt : CUDART_INF;
            double q1 = e1 < n ? (mix_t * a1 + m2 * b1) / mt : CUDART_INF;
            int x0 = e0, x1 = e1;
            // deq_chem.py:582 (stable argsort) as a 64-element bitonic network on (key, flat index): partner
            // distance 1 is the lane's own pair, larger distances are lane-xor shuffles - no shared memory,
            // no warp barriers (the index makes the order strict, so min/max selection is unambiguous)
#pragma unroll
            for (int k = 2; k <= kElems; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const bool up = ((2 * lane) & k) == 0;
                    if (j == 1) {
                        const bool swap = key_less(q1, x1, q0, x0) == up;
                        const double tq = swap ? q1 : q0;
                        const int tx = swap ? x1 : x0;
                        q1 = swap ? q0 : q1; x1 = swap ? x0 : x1;
                        q0 = tq; x0 = tx;
                    } else {
                        const int m = j >> 1;
                        const double oq0 = __shfl_xor_sync(0xffffffffu, q0, m), oq1 = __shfl_xor_sync(0xffffffffu, q1, m);
                        const int ox0 = __shfl_xor_sync(0xffffffffu, x0, m), ox1 = __shfl_xor_sync(0xffffffffu, x1, m);
                        const bool keep_min = ((lane & m) == 0) == up;
                        const bool l0 = key_less(oq0, ox0, q0, x0), l1 = key_less(oq1, ox1, q1, x1);  // other < mine
                        const bool t0 = (l0 == keep_min), t1 = (l1 == keep_min);                      // take the other
                        q0 = t0 ? oq0 : q0; x0 = t0 ? ox0 : x0;
                        q1 = t1 ? oq1 : q1; x1 = t1 ? ox1 : x1;
                    }
                }
            }
            key[e0] = q0; key[e1] = q1;
            idx[e0] = (unsigned char)x0; idx[e1] = (unsigned char)x1;
            // deq_chem.py:585-590: cumulative weights -> x in (0, 1]
            const int s0 = idx[e0], s1 = idx[e1];
            const double w0 = e0 < n ? s_gw[s0 / K] * s_gw[s0 % K] : 0.0;
            const double w1 = e1 < n ? s_gw[s1 / K] * s_gw[s1 % K] : 0.0;
            double incl = w0 + w1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const double total = __shfl_sync(0xffffffffu, incl, 31);
            const double c0 = incl - w1, c1 = incl;
            const double y0 = log10(key[e0]), y1 = log10(key[e1]);
            __syncwarp();
            xs[e0] = c0 / total;
            xs[e1] = c1 / total;
            key[e0] = y0;
            key[e1] = y1;
            __syncwarp();
            // deq_chem.py:595: np.interp(gauss_pts, x, log10 k) then 10**
            if (glane) {
                double v;
                if (gp < xs[0]) v = key[0];
                else if (gp > xs[n - 1]) v = key[n - 1];
                else {
                    int lo = 0, hi = n;  // largest j with xs[j] <= gp
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (xs[mid] <= gp) lo = mid; else hi = mid;
                    }
                    if (lo == n - 1 || xs[lo] == gp) v = key[lo];
                    else {
                        const double slope = (key[lo + 1] - key[lo]) / (xs[lo + 1] - xs[lo]);
                        v = slope * (gp - xs[lo]) + key[lo];
                    }
                }
                k1 = exp10(v);
            }
            mix_t = mt;
            __syncwarp();
        }
        lnk[ct] = log(k1);
        if (p.ln_mixed && glane) p.ln_mixed[((int64_t)bin * K + lane) * 4 + ct] = lnk[ct];
    }
    if (glane) {
        // optics.py:1189-1197
        const double t = p.t_interp[il], q = p.p_interp[il];
        const double ln_kappa = (((1 - t) * (1 - q) * lnk[0]) + ((t) * (1 - q) * lnk[1]) + ((t) * (q) * lnk[3]) +
                                 ((1 - t) * (q) * lnk[2]));
        p.molecular_opa[(int64_t)bin * K + lane] = exp(ln_kappa) * 6.02214086e+23;
    }
}

}  // namespace

extern "C" int pb_ck_mix(pb_ctx *ctx, const pb_ck_mix_args *a, int memspace)
{
    if (!ctx || !a) return PB_ERR_ARG;
    const int L = a->nlayer, W = a->nwno, K = a->ngauss, G = a->ngas;
    if (L < 0 || W < 0 || G < 1 || G > kMaxGas) return pb_fail(ctx, PB_ERR_ARG, "ck_mix: bad sizes nlayer=%d nwno=%d ngas=%d (max %d gases)", L, W, G, kMaxGas);
    if (K < 1 || K > 8) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "ck_mix: ngauss=%d, supported 1..8 (Nk^2 <= 64 products per warp)", K);
    if (L == 0 || W == 0) return PB_OK;
    if (!a->kappas || !a->mixes || !a->indices || !a->t_interp || !a->p_interp || !a->gauss_pts || !a->gauss_wts || !a->molecular_opa)
        return pb_fail(ctx, PB_ERR_ARG, "ck_mix: NULL argument");
    for (int l = 0; l < L; ++l) {
        for (int k = 0; k < 4; ++k) {
            const int v = a->indices[k * L + l], lim = k < 2 ? a->np : a->nt;
            if (v < 0 || v >= lim) return pb_fail(ctx, PB_ERR_ARG, "ck_mix: indices[%d][%d]=%d out of range", k, l, v);
        }
    }
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const size_t nout = (size_t)L * W * K * sizeof(double);
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, 16 * 256 + (host ? pb_align(nout) + (a->ln_mixed ? pb_align(4 * nout) : 0) : 0) +
                                     pb_align(((size_t)G * L + 4 * L + 2 * L + 16 + 8) * 8)));
    PB_TRY(pb_pinned_reserve(ctx, ((size_t)G * L + 4 * L + 2 * L + 2 * K + 64) * sizeof(double)));
    MixParams p{};
    p.L = L; p.W = W; p.K = K; p.ngas = G; p.np = a->np; p.nt = a->nt;
    for (int g = 0; g < G; ++g) {
        if (!a->kappas[g]) return pb_fail(ctx, PB_ERR_ARG, "ck_mix: kappas[%d] is NULL", g);
        p.kappa[g] = a->kappas[g];
    }
    const double *tmp;
    PB_TRY(pb_upload_small(ctx, a->mixes, (size_t)G * L, &p.mixes));
    PB_TRY(pb_upload_small(ctx, (const double *)a->indices, ((size_t)4 * L * sizeof(int) + 7) / 8, &tmp));
    p.indices = (const int *)tmp;
    PB_TRY(pb_upload_small(ctx, a->t_interp, (size_t)L, &p.t_interp));
    PB_TRY(pb_upload_small(ctx, a->p_interp, (size_t)L, &p.p_interp));
    PB_TRY(pb_upload_small(ctx, a->gauss_pts, (size_t)K, &p.gauss_pts));
    PB_TRY(pb_upload_small(ctx, a->gauss_wts, (size_t)K, &p.gauss_wts));
    p.molecular_opa = a->molecular_opa;
    p.ln_mixed = a->ln_mixed;
    if (host) {
        PB_TRY(pb_arena_alloc(ctx, nout, (void **)&p.molecular_opa));
        if (a->ln_mixed) PB_TRY(pb_arena_alloc(ctx, 4 * nout, (void **)&p.ln_mixed));
    }
    PB_TRY(pb_upload_flush(ctx));
    const int64_t bins = (int64_t)L * W;
    ck_mix_kernel<<<(unsigned)((bins + kWarps - 1) / kWarps), kWarps * 32, 0, ctx->stream>>>(p);
    PB_CHECK_LAUNCH(ctx);
    if (host) {
        PB_CUDA(ctx, cudaMemcpyAsync(a->molecular_opa, p.molecular_opa, nout, cudaMemcpyDeviceToHost, ctx->stream));
        if (a->ln_mixed) PB_CUDA(ctx, cudaMemcpyAsync(a->ln_mixed, p.ln_mixed, 4 * nout, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
