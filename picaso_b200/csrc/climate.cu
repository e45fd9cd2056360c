// climate.cu - the radiative-transfer call of the climate solver, on the device in one call.
//
// Replaces picaso/climate.py:1686-1952 (get_fluxes): for every correlated-k gauss point the
// reference slices X[:, :, ig] out of the [nlayer, nwno, ngauss] opacity arrays, calls
// get_reflected_1d(get_lvl_flux=1) with a single mu = 0.5 stream and get_thermal_1d(calc_type=1),
// weights the four level arrays by gauss_wts[ig] and reduces them over wavelength
// (np.sum(axis=3) for the visible net fluxes, a sequential dwni-weighted loop for the IR ones,
// compress_thermal over the disk angles in between).
//
// Here: one de-interleave pass turns [rows][nwno][ngauss] into ngauss dense [rows][nwno] blocks
// (wavelength fastest, what the flux kernels coalesce on), the level-flux kernels run ONCE with
// nbatch = ngauss, and two reduction kernels fold gauss weights, disk weights, dwni and the
// wavelength sums.  Level arrays never leave HBM; what returns to the host is 4 [nlevel] vectors
// and 4 [nlevel][nwno] arrays.
#include <cstdlib>

#include "pb_common.cuh"

namespace {

// [rows][W][K] -> [K][rows][W]
__global__ void deinterleave_kernel(int64_t n /* rows*W */, int K, const double *__restrict__ in,
                                    double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *src = in + i * K;
    for (int k = 0; k < K; ++k) out[(int64_t)k * n + i] = src[k];
}

// out[k][w] = in[w]
__global__ void replicate_kernel(int W, int K, const double *__restrict__ in, double fill, double *__restrict__ out)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const double v = in ? in[w] : fill;
    for (int k = 0; k < K; ++k) out[(int64_t)k * W + w] = v;
}

constexpr int kRedThreads = 256;

// deterministic CTA-wide sum (fixed shuffle tree + fixed order over warps)
__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kRedThreads / 32; ++i) t += sh[i];
    return t;  // valid on thread 0
}

// Visible part (climate.py:1838-1840, :1872-1873).  Level arrays [K][V][W] (one mu = 0.5 stream).
// One CTA per level.
__global__ void __launch_bounds__(kRedThreads) climate_visible_reduce(
    int V, int W, int K, const double *__restrict__ fm, const double *__restrict__ fp,
    const double *__restrict__ fmm, const double *__restrict__ fpm, const double *__restrict__ gw,
    double *__restrict__ plus_v, double *__restrict__ minus_v, double *__restrict__ net_layer,
    double *__restrict__ net)
{
    __shared__ double sh[kRedThreads / 32];
    const int v = blockIdx.x;
    double nl = 0.0, nn = 0.0;
    for (int k = 0; k < K; ++k) {
        const int64_t base = ((int64_t)k * V + v) * W;
        double s_fm = 0.0, s_fp = 0.0, s_fmm = 0.0, s_fpm = 0.0;
        for (int w = threadIdx.x; w < W; w += kRedThreads) {
            s_fm += fm[base + w];
            s_fp += fp[base + w];
            s_fmm += fmm[base + w];
            s_fpm += fpm[base + w];
        }
        const double t_fpm = block_sum(s_fpm, sh), t_fmm = block_sum(s_fmm, sh);
        const double t_fp = block_sum(s_fp, sh), t_fm = block_sum(s_fm, sh);
        if (threadIdx.x == 0) {
            nl += (t_fpm - t_fmm) * gw[k];
            nn += (t_fp - t_fm) * gw[k];
        }
    }
    if (threadIdx.x == 0) {
        net_layer[v] = nl;
        net[v] = nn;
    }
    for (int w = threadIdx.x; w < W; w += kRedThreads) {
        double ap = 0.0, am = 0.0;
        for (int k = 0; k < K; ++k) {
            const int64_t i = ((int64_t)k * V + v) * W + w;
            ap += fp[i] * gw[k];
            am += fm[i] * gw[k];
        }
        plus_v[(int64_t)v * W + w] = ap;
        minus_v[(int64_t)v * W + w] = am;
    }
}

// IR part (climate.py:1916-1942): gauss-weighted accumulation, compress_thermal (disco.py:152-180)
// over the G = ng*nt disk angles, dwni-weighted wavelength sums.  Level arrays [K][G][V][W].
__global__ void __launch_bounds__(kRedThreads) climate_ir_reduce(
    int V, int W, int K, int G, int nt, const double *fm, const double *fp,
    const double *fmm, const double *fpm, const double *__restrict__ gw,
    const double *__restrict__ gweight, const double *__restrict__ tweight, const double *__restrict__ dwni,
    double *plus_ir, double *minus_ir, double *net_layer, double *net)
{
    __shared__ double sh[kRedThreads / 32];
    const int v = blockIdx.x;
    // blockIdx.y: temperature profile of a Jacobian batch (level arrays [P][K][G][V][W], net fluxes [P][V]); the
    // [V][W] arrays are only written for a single profile
    const int64_t pofs = (int64_t)blockIdx.y * K * G * V * W;
    fm += pofs; fp += pofs; fmm += pofs; fpm += pofs;
    net_layer += (int64_t)blockIdx.y * V; net += (int64_t)blockIdx.y * V;
    const double sym = (nt == 1) ? 1.0 : 1.0 / (2.0 * PB_PI);
    double s_net = 0.0, s_lay = 0.0;
    for (int w = threadIdx.x; w < W; w += kRedThreads) {
        double c_fm = 0.0, c_fp = 0.0, c_fmm = 0.0, c_fpm = 0.0;
        for (int a = 0; a < G; ++a) {
            double a_fm = 0.0, a_fp = 0.0, a_fmm = 0.0, a_fpm = 0.0;
            for (int k = 0; k < K; ++k) {
                const int64_t i = (((int64_t)k * G + a) * V + v) * W + w;
                a_fm += fm[i] * gw[k];
                a_fp += fp[i] * gw[k];
                a_fmm += fmm[i] * gw[k];
                a_fpm += fpm[i] * gw[k];
            }
            const int ig = a / nt, it = a - ig * nt;
            const double wgt_g = gweight[ig], wgt_t = tweight[it];
            c_fm = c_fm + a_fm * wgt_g * wgt_t;
            c_fp = c_fp + a_fp * wgt_g * wgt_t;
            c_fmm = c_fmm + a_fmm * wgt_g * wgt_t;
            c_fpm = c_fpm + a_fpm * wgt_g * wgt_t;
        }
        c_fm *= sym; c_fp *= sym; c_fmm *= sym; c_fpm *= sym;
        const double dw = dwni[w];
        s_lay += (c_fpm - c_fmm) * dw;
        s_net += (c_fp - c_fm) * dw;
        if (plus_ir) plus_ir[(int64_t)v * W + w] = c_fp * dw;
        if (minus_ir) minus_ir[(int64_t)v * W + w] = c_fm * dw;
    }
    const double t_lay = block_sum(s_lay, sh), t_net = block_sum(s_net, sh);
    if (threadIdx.x == 0) {
        net_layer[v] = t_lay;
        net[v] = t_net;
    }
}

int aux_reserve(pb_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->aux_cap) return PB_OK;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->aux) PB_CUDA(ctx, cudaFree(ctx->aux));
    ctx->aux = nullptr;
    ctx->aux_cap = 0;
    const size_t cap = pb_align(bytes + bytes / 8, 1 << 20);
    cudaError_t e = cudaMalloc((void **)&ctx->aux, cap);
    if (e != cudaSuccess) return pb_fail(ctx, PB_ERR_NOMEM, "climate scratch cudaMalloc(%zu) -> %s", cap, cudaGetErrorString(e));
    ctx->aux_cap = cap;
    return PB_OK;
}

} // namespace

// Jacobian batching (pb_climate_args.nprofiles > 0): the thermal half of get_fluxes for P temperature profiles over one
// set of opacities.  The three opacity arrays are staged (de-interleaved) once; pb_thermal_toon_1d runs with
// nbatch = P x K and opacity_period = K, in chunks of profiles whose level arrays fit the scratch budget; one
// climate_ir_reduce launch per chunk folds gauss weights, disk weights, dwni and the wavelength sums per profile.
static int climate_jacobian(pb_ctx *ctx, const pb_climate_args *a, int memspace)
{
    const int L = a->nlayer, W = a->nwno, K = a->ngauss, G = a->numg * a->numt, V = L + 1, P = a->nprofiles;
    if (!a->DTAU_OG || !a->W0_no_raman || !a->COSB_OG || !a->tlevels || !a->plevel || !a->wno || !a->dwno ||
        !a->ubar1 || !a->gweight || !a->tweight || !a->jac_out)
        return pb_fail(ctx, PB_ERR_ARG, "climate jacobian: needs DTAU_OG, W0_no_raman, COSB_OG, tlevels, plevel, wno, dwno, ubar1, gweight, tweight, jac_out");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const size_t nLW = (size_t)L * W * sizeof(double), nVW = (size_t)V * W * sizeof(double), nW = (size_t)W * sizeof(double);
    const bool restage = host || K > 1;
    // profiles per chunk: four level arrays of Pc x K x G x V x W doubles within ~3 GB
    const size_t per_profile = 4 * (size_t)K * G * nVW;
    int Pc = (int)(((size_t)3 << 30) / per_profile);
    if (Pc < 1) Pc = 1;
    if (Pc > P) Pc = P;
    if ((int64_t)Pc * K > 65535) Pc = 65535 / K;
    size_t need = 64 * 256 + (restage ? 3 * pb_align(K * nLW) : 0) + (host ? pb_align(K * nLW) : 0) + pb_align(K * nW) +
                  4 * pb_align((size_t)Pc * K * G * nVW) + pb_align(2 * (size_t)P * V * 8) + pb_align((size_t)(K + G + 16) * 8 * 4);
    PB_TRY(aux_reserve(ctx, need));
    size_t off = 0;
    auto take = [&](size_t bytes) -> double * {
        double *ptr = (double *)(ctx->aux + off);
        off += pb_align(bytes);
        return ptr;
    };
    double *raw = host ? take(K * nLW) : nullptr;
    auto stage = [&](const double *src, const double **dst) -> int {
        if (!restage) { *dst = src; return PB_OK; }
        const int64_t n = (int64_t)L * W;
        double *d = take((size_t)K * n * sizeof(double));
        const double *in = src;
        if (host) {
            PB_CUDA(ctx, cudaMemcpyAsync(raw, src, (size_t)K * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            in = raw;
        }
        if (K == 1) PB_CUDA(ctx, cudaMemcpyAsync(d, in, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        else {
            deinterleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, K, in, d);
            PB_CHECK_LAUNCH(ctx);
        }
        *dst = d;
        return PB_OK;
    };
    pb_arena_reset(ctx);
    {
        const size_t Gc = (size_t)G;
        const size_t outer = (size_t)(K + 4 * (size_t)W + a->numg + a->numt + 64) * sizeof(double) + 16 * 64;
        const size_t inner = (2 * (size_t)Pc * K * V + 3 * Gc + (size_t)Pc * K + 64) * sizeof(double);
        PB_TRY(pb_pinned_reserve(ctx, outer > inner ? outer : inner));
    }
    const double *d_dtau = nullptr, *d_w0 = nullptr, *d_cosb = nullptr;
    PB_TRY(stage(a->DTAU_OG, &d_dtau));
    PB_TRY(stage(a->W0_no_raman, &d_w0));
    PB_TRY(stage(a->COSB_OG, &d_cosb));
    const double *d_gw = nullptr, *d_wno = nullptr, *d_dwno = nullptr, *d_gweight = nullptr, *d_tweight = nullptr, *d_s1 = nullptr;
    PB_TRY(pb_upload_small(ctx, a->gauss_wts, K, &d_gw));
    PB_TRY(pb_upload_small(ctx, a->wno, W, &d_wno));
    PB_TRY(pb_upload_small(ctx, a->dwno, W, &d_dwno));
    PB_TRY(pb_upload_small(ctx, a->gweight, a->numg, &d_gweight));
    PB_TRY(pb_upload_small(ctx, a->tweight, a->numt, &d_tweight));
    if (a->surf_reflect) PB_TRY(pb_upload_small(ctx, a->surf_reflect, W, &d_s1));
    double *d_surf = take(K * nW);
    PB_TRY(pb_upload_flush(ctx));
    replicate_kernel<<<(W + 255) / 256, 256, 0, ctx->stream>>>(W, K, d_s1, 0.0, d_surf);
    PB_CHECK_LAUNCH(ctx);
    double *lv[4];
    for (int i = 0; i < 4; ++i) lv[i] = take((size_t)Pc * K * G * nVW);
    double *o_net = take(2 * (size_t)P * V * 8);   // [2][P][V]: layer, level
    if (off > ctx->aux_cap) return pb_fail(ctx, PB_ERR_NOMEM, "climate jacobian: scratch layout overflow (%zu > %zu)", off, ctx->aux_cap);
    std::vector<double> tl, pl;
    for (int p0 = 0; p0 < P; p0 += Pc) {
        const int pc = P - p0 < Pc ? P - p0 : Pc;
        tl.assign((size_t)pc * K * V, 0.0);
        pl.assign((size_t)pc * K * V, 0.0);
        for (int pr = 0; pr < pc; ++pr)
            for (int k = 0; k < K; ++k)
                for (int v = 0; v < V; ++v) {
                    tl[((size_t)pr * K + k) * V + v] = a->tlevels[(size_t)(p0 + pr) * V + v];
                    pl[((size_t)pr * K + k) * V + v] = a->plevel[v];
                }
        pb_thermal_args t;
        memset(&t, 0, sizeof(t));
        t.nlayer = L; t.nwno = W; t.numg = a->numg; t.numt = a->numt; t.nbatch = pc * K; t.ld = W;
        t.dtau = d_dtau; t.w0 = d_w0; t.cosb = d_cosb;
        t.wno = d_wno; t.dwno = d_dwno; t.surf_reflect = d_surf;
        t.tlevel = tl.data(); t.plevel = pl.data();
        t.ubar1 = a->ubar1; t.gweight = a->gweight; t.tweight = a->tweight;
        t.hard_surface = 0; t.calc_type = 1;
        t.opacity_period = K;
        t.flux_minus = lv[0]; t.flux_plus = lv[1]; t.flux_minus_mdpt = lv[2]; t.flux_plus_mdpt = lv[3];
        PB_TRY(pb_thermal_toon_1d(ctx, &t, PB_DEVICE));
        dim3 grid(V, pc);
        climate_ir_reduce<<<grid, kRedThreads, 0, ctx->stream>>>(V, W, K, G, a->numt, lv[0], lv[1], lv[2], lv[3], d_gw,
                                                                 d_gweight, d_tweight, d_dwno, nullptr, nullptr,
                                                                 o_net + (size_t)p0 * V, o_net + (size_t)(P + p0) * V);
        PB_CHECK_LAUNCH(ctx);
    }
    // jac_out [P][2][V] <- o_net [2][P][V]: two strided copies
    PB_CUDA(ctx, cudaMemcpy2DAsync(a->jac_out, 2 * (size_t)V * 8, o_net, (size_t)V * 8, (size_t)V * 8, P,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaMemcpy2DAsync(a->jac_out + V, 2 * (size_t)V * 8, o_net + (size_t)P * V, (size_t)V * 8, (size_t)V * 8, P,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

extern "C" int pb_climate_get_fluxes(pb_ctx *ctx, const pb_climate_args *a, int memspace)
{
    if (!ctx || !a) return PB_ERR_ARG;
    const int L = a->nlayer, W = a->nwno, K = a->ngauss, G = a->numg * a->numt;
    const int V = L + 1;
    if (L < 1 || W < 1 || K < 1 || G < 1) return pb_fail(ctx, PB_ERR_ARG, "climate: bad sizes L=%d W=%d K=%d G=%d", L, W, K, G);
    if (!a->gauss_wts) return pb_fail(ctx, PB_ERR_ARG, "climate: gauss_wts missing");
    if (a->nprofiles > 0) return climate_jacobian(ctx, a, memspace);
    if (a->reflected) {
        if (!a->DTAU || !a->TAU || !a->W0 || !a->COSB || !a->GCOS2 || !a->ftau_cld || !a->ftau_ray ||
            !a->DTAU_OG || !a->TAU_OG || !a->W0_OG || !a->COSB_OG)
            return pb_fail(ctx, PB_ERR_ARG, "climate: reflected needs the 11 opacity arrays");
        if (!a->packed && (!a->flux_net_v_layer || !a->flux_net_v))
            return pb_fail(ctx, PB_ERR_ARG, "climate: reflected outputs missing");
    }
    if (a->thermal) {
        if (!a->DTAU_OG || !a->W0_no_raman || !a->COSB_OG || !a->tlevel || !a->plevel || !a->wno || !a->dwno ||
            !a->ubar1 || !a->gweight || !a->tweight)
            return pb_fail(ctx, PB_ERR_ARG, "climate: thermal needs DTAU_OG, W0_no_raman, COSB_OG, tlevel, plevel, wno, dwno, ubar1, gweight, tweight");
        if (!a->packed && (!a->flux_net_ir_layer || !a->flux_net_ir))
            return pb_fail(ctx, PB_ERR_ARG, "climate: thermal outputs missing");
    }
    if (!a->reflected && !a->thermal) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const size_t nLW = (size_t)L * W * sizeof(double), nVW = (size_t)V * W * sizeof(double), nW = (size_t)W * sizeof(double);

    // ---- scratch layout (context-owned, survives the arena resets of the flux entry points) ----
    const bool restage = host || K > 1;
    const int n_lay = restage ? ((a->reflected ? 9 : 0) + (a->thermal ? (a->reflected ? 1 : 3) : 0)) : 0;
    const int n_lev = (restage && a->reflected) ? 2 : 0;
    const int Gmax = a->thermal ? (G > 1 ? G : 1) : 1;
    size_t need = 64 * 256;
    need += (size_t)n_lay * pb_align(K * nLW) + (size_t)n_lev * pb_align(K * nVW);
    if (host) need += pb_align(K * nVW);                        // raw [rows][W][K] staging
    need += 2 * pb_align(K * nW) + 2 * pb_align(nW);             // surf, F0PI replicated; wno, dwno
    need += ((a->reflected && a->thermal) ? 8 : 4) * pb_align((size_t)K * Gmax * nVW);  // level arrays
    need += pb_align(4 * nVW + 4 * (size_t)V * 8) + 256;           // reduced outputs (one block)
    need += pb_align((size_t)(K + G + 16) * 8 * 4);
    PB_TRY(aux_reserve(ctx, need));
    size_t off = 0;
    auto take = [&](size_t bytes) -> double * {
        double *ptr = (double *)(ctx->aux + off);
        off += pb_align(bytes);
        return ptr;
    };
    double *raw = host ? take(K * nVW) : nullptr;
    auto stage = [&](const double *src, int rows, const double **dst) -> int {
        if (!src) { *dst = nullptr; return PB_OK; }
        if (!restage) { *dst = src; return PB_OK; }
        const int64_t n = (int64_t)rows * W;
        double *d = take((size_t)K * n * sizeof(double));
        const double *in = src;
        if (host) {
            PB_CUDA(ctx, cudaMemcpyAsync(raw, src, (size_t)K * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            in = raw;
        }
        if (K == 1) {
            PB_CUDA(ctx, cudaMemcpyAsync(d, in, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            deinterleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, K, in, d);
            PB_CHECK_LAUNCH(ctx);
        }
        *dst = d;
        return PB_OK;
    };
    // O(K), O(G), O(W) vectors are host pointers in both memory spaces: they travel in ONE packed asynchronous
    // copy through the context's pinned ring (pageable cudaMemcpyAsync calls would each block the host)
    pb_arena_reset(ctx);  // next pinned slot
    {
        // The flux entry points called below reserve pinned space themselves, and a reserve that has to grow the ring
        // frees EVERY slot - including the one holding this call's vectors.  Reserve the largest of the three needs
        // here (pb_pinned_reserve grows all slots to twice the request), so the inner reserves never reallocate.
        const size_t Gc = (size_t)a->numg * a->numt;
        const size_t outer = (size_t)(K + 4 * (size_t)W + a->numg + a->numt + 64) * sizeof(double) + 16 * 64;
        const size_t inner_refl = 4 * (Gc + (size_t)K + 16) * sizeof(double);
        const size_t inner_therm = (2 * (size_t)K * V + 3 * Gc + (size_t)K + 64) * sizeof(double);
        size_t need_pin = outer;
        if (inner_refl > need_pin) need_pin = inner_refl;
        if (inner_therm > need_pin) need_pin = inner_therm;
        PB_TRY(pb_pinned_reserve(ctx, need_pin));
    }
    auto small_to_dev = [&](const double *src, size_t n, const double **dst) -> int {
        return pb_upload_small(ctx, src, n, dst);
    };
    const double *d_dtau = nullptr, *d_tau = nullptr, *d_w0 = nullptr, *d_cosb = nullptr, *d_gcos2 = nullptr,
                 *d_fcld = nullptr, *d_fray = nullptr, *d_dtau_og = nullptr, *d_tau_og = nullptr, *d_w0_og = nullptr,
                 *d_cosb_og = nullptr, *d_w0nr = nullptr;
    if (a->reflected) {
        PB_TRY(stage(a->DTAU, L, &d_dtau));
        PB_TRY(stage(a->TAU, V, &d_tau));
        PB_TRY(stage(a->W0, L, &d_w0));
        PB_TRY(stage(a->COSB, L, &d_cosb));
        PB_TRY(stage(a->GCOS2, L, &d_gcos2));
        PB_TRY(stage(a->ftau_cld, L, &d_fcld));
        PB_TRY(stage(a->ftau_ray, L, &d_fray));
        PB_TRY(stage(a->TAU_OG, V, &d_tau_og));
        PB_TRY(stage(a->W0_OG, L, &d_w0_og));
    }
    PB_TRY(stage(a->DTAU_OG, L, &d_dtau_og));
    PB_TRY(stage(a->COSB_OG, L, &d_cosb_og));
    if (a->thermal) PB_TRY(stage(a->W0_no_raman, L, &d_w0nr));

    const double *d_gw = nullptr, *d_wno = nullptr, *d_dwno = nullptr, *d_gweight = nullptr, *d_tweight = nullptr;
    PB_TRY(small_to_dev(a->gauss_wts, K, &d_gw));
    if (a->thermal) {
        PB_TRY(small_to_dev(a->wno, W, &d_wno));
        PB_TRY(small_to_dev(a->dwno, W, &d_dwno));
        PB_TRY(small_to_dev(a->gweight, a->numg, &d_gweight));
        PB_TRY(small_to_dev(a->tweight, a->numt, &d_tweight));
    }
    double *d_surf = take(K * nW), *d_f0 = take(K * nW);
    {
        // per-wave vectors are host pointers (they come from the solver's tuples), shared by all gauss points
        const double *d_s1 = nullptr, *d_f1 = nullptr;
        if (a->surf_reflect) PB_TRY(small_to_dev(a->surf_reflect, W, &d_s1));
        if (a->F0PI) PB_TRY(small_to_dev(a->F0PI, W, &d_f1));
        PB_TRY(pb_upload_flush(ctx));
        replicate_kernel<<<(W + 255) / 256, 256, 0, ctx->stream>>>(W, K, d_s1, 0.0, d_surf);
        PB_CHECK_LAUNCH(ctx);
        replicate_kernel<<<(W + 255) / 256, 256, 0, ctx->stream>>>(W, K, d_f1, 1.0, d_f0);
        PB_CHECK_LAUNCH(ctx);
    }
    // level arrays: the visible half needs [K][1][V][W], the IR half [K][G][V][W]; separate blocks when both run
    // (they run concurrently)
    double *lv[4], *lvi[4];
    for (int i = 0; i < 4; ++i) lv[i] = take((size_t)K * Gmax * nVW);
    for (int i = 0; i < 4; ++i) lvi[i] = (a->reflected && a->thermal) ? take((size_t)K * Gmax * nVW) : lv[i];
    // reduced outputs in ONE contiguous block - [4][V] net fluxes (visible layer, visible level, IR layer, IR
    // level), then [4][V][W] (plus_v, minus_v, plus_ir, minus_ir) - so that `packed` callers get them with a
    // single device-to-host copy instead of eight host-blocking ones
    double *blk = take(4 * (size_t)V * 8 + 4 * nVW);
    double *o_lay = blk, *o_net = blk + V, *o_lay_ir = blk + 2 * V, *o_net_ir = blk + 3 * V;
    double *o_plus = blk + 4 * V, *o_minus = o_plus + (size_t)V * W, *o_plus_ir = o_minus + (size_t)V * W,
           *o_minus_ir = o_plus_ir + (size_t)V * W;
    if (off > ctx->aux_cap) return pb_fail(ctx, PB_ERR_NOMEM, "climate: scratch layout overflow (%zu > %zu)", off, ctx->aux_cap);
    if (a->packed && !(a->reflected && a->thermal))  // the half that does not run reads as zeros (climate.py:1757-1786)
        PB_CUDA(ctx, cudaMemsetAsync(blk, 0, 4 * (size_t)V * 8 + 4 * nVW, ctx->stream));

    if (a->reflected && a->thermal) PB_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));  // staging done
    if (a->reflected) {
        // climate.py:1803-1816: one stream at mu0 = mu1 = 0.5, fluxes only
        pb_reflected_args r;
        memset(&r, 0, sizeof(r));
        r.nlayer = L; r.nwno = W; r.numg = 1; r.numt = 1; r.nbatch = K; r.ld = W;
        r.dtau = d_dtau; r.w0 = d_w0; r.cosb = d_cosb; r.gcos2 = d_gcos2; r.ftau_cld = d_fcld; r.ftau_ray = d_fray;
        r.dtau_og = d_dtau_og; r.w0_og = d_w0_og; r.cosb_og = d_cosb_og; r.tau = d_tau; r.tau_og = d_tau_og;
        r.surf_reflect = d_surf; r.F0PI = d_f0; r.b_top = nullptr;
        const double half = 0.5, one = 1.0;
        r.ubar0 = &half; r.ubar1 = &half; r.gweight = &one; r.tweight = &one;
        r.cos_theta = a->cos_theta;
        r.single_phase = a->single_phase; r.multi_phase = a->multi_phase; r.toon_coefficients = 0;
        r.frac_a = a->frac_a; r.frac_b = a->frac_b; r.frac_c = a->frac_c;
        r.constant_back = a->constant_back; r.constant_forward = a->constant_forward;
        r.get_toa_intensity = 0; r.get_lvl_flux = 1;
        r.flux_minus = lv[0]; r.flux_plus = lv[1]; r.flux_minus_mdpt = lv[2]; r.flux_plus_mdpt = lv[3];
        PB_TRY(pb_reflected_toon_1d(ctx, &r, PB_DEVICE));
        climate_visible_reduce<<<V, kRedThreads, 0, ctx->stream>>>(V, W, K, lv[0], lv[1], lv[2], lv[3], d_gw, o_plus,
                                                                   o_minus, o_lay, o_net);
        PB_CHECK_LAUNCH(ctx);
        // the copies back to (pageable) host arrays block the host until the stream gets there: they are issued
        // further down, after the IR half has been enqueued on the side stream
    }
    // The visible and the IR halves are independent and each is a handful of small, latency-bound launches
    // (the level kernels walk nlayer layers three times with ~20 CTAs per gauss point): run the IR half on the
    // context's side stream, concurrently with the visible half.
    cudaStream_t main_stream = ctx->stream;
    const bool overlap = a->reflected && a->thermal;
    if (overlap) {
        // everything staged so far (inputs, weights) was enqueued on the main stream before the visible half;
        // the side stream must see it, so fork from an event recorded BEFORE the visible launches.  The event
        // was recorded below, right after staging.
        PB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
        ctx->stream = ctx->copy_stream;
    }
    auto run_thermal = [&]() -> int {
        // climate.py:1887-1892: OG optical depths, W0_no_raman, hard_surface = 0, calc_type = 1
        std::vector<double> tl((size_t)K * V), pl((size_t)K * V);
        for (int k = 0; k < K; ++k)
            for (int v = 0; v < V; ++v) {
                tl[(size_t)k * V + v] = a->tlevel[v];
                pl[(size_t)k * V + v] = a->plevel[v];
            }
        pb_thermal_args t;
        memset(&t, 0, sizeof(t));
        t.nlayer = L; t.nwno = W; t.numg = a->numg; t.numt = a->numt; t.nbatch = K; t.ld = W;
        t.dtau = d_dtau_og; t.w0 = d_w0nr; t.cosb = d_cosb_og;
        t.wno = d_wno; t.dwno = d_dwno; t.surf_reflect = d_surf;
        t.tlevel = tl.data(); t.plevel = pl.data();
        t.ubar1 = a->ubar1; t.gweight = a->gweight; t.tweight = a->tweight;
        t.hard_surface = 0; t.calc_type = 1;
        t.flux_minus = lvi[0]; t.flux_plus = lvi[1]; t.flux_minus_mdpt = lvi[2]; t.flux_plus_mdpt = lvi[3];
        PB_TRY(pb_thermal_toon_1d(ctx, &t, PB_DEVICE));
        climate_ir_reduce<<<V, kRedThreads, 0, ctx->stream>>>(V, W, K, G, a->numt, lvi[0], lvi[1], lvi[2], lvi[3], d_gw,
                                                              d_gweight, d_tweight, d_dwno, o_plus_ir, o_minus_ir,
                                                              o_lay_ir, o_net_ir);
        PB_CHECK_LAUNCH(ctx);
        if (!a->packed) {
            if (a->flux_plus_ir) PB_CUDA(ctx, cudaMemcpyAsync(a->flux_plus_ir, o_plus_ir, nVW, cudaMemcpyDeviceToHost, ctx->stream));
            if (a->flux_minus_ir) PB_CUDA(ctx, cudaMemcpyAsync(a->flux_minus_ir, o_minus_ir, nVW, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(ctx, cudaMemcpyAsync(a->flux_net_ir_layer, o_lay_ir, (size_t)V * 8, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(ctx, cudaMemcpyAsync(a->flux_net_ir, o_net_ir, (size_t)V * 8, cudaMemcpyDeviceToHost, ctx->stream));
        } else if (overlap) {
            PB_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[15], ctx->stream));  // IR half done (side stream)
        }
        // tl / pl are consumed by pb_thermal_toon_1d's packed pinned upload before it returns
        return PB_OK;
    };
    int rc_thermal = PB_OK;
    if (a->thermal) rc_thermal = run_thermal();
    if (overlap) ctx->stream = main_stream;  // restored on every path before any return below
    if (a->packed) {
        if (overlap && rc_thermal == PB_OK) PB_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_chunk[15], 0));
        const size_t nb = 4 * (size_t)V * 8 + (a->packed_full ? 4 * nVW : 0);
        if (rc_thermal == PB_OK) PB_CUDA(ctx, cudaMemcpyAsync(a->packed, blk, nb, cudaMemcpyDeviceToHost, main_stream));
    } else if (a->reflected) {
        if (a->flux_plus_v) PB_CUDA(ctx, cudaMemcpyAsync(a->flux_plus_v, o_plus, nVW, cudaMemcpyDeviceToHost, main_stream));
        if (a->flux_minus_v) PB_CUDA(ctx, cudaMemcpyAsync(a->flux_minus_v, o_minus, nVW, cudaMemcpyDeviceToHost, main_stream));
        PB_CUDA(ctx, cudaMemcpyAsync(a->flux_net_v_layer, o_lay, (size_t)V * 8, cudaMemcpyDeviceToHost, main_stream));
        PB_CUDA(ctx, cudaMemcpyAsync(a->flux_net_v, o_net, (size_t)V * 8, cudaMemcpyDeviceToHost, main_stream));
    }
    if (overlap) {
        if (rc_thermal == PB_OK) PB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        else cudaStreamSynchronize(ctx->copy_stream);
    }
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return rc_thermal;
}

// ---- one call per reflected spectrum: opacity -> flux -> disk integration (include/picaso_b200.h) ----------
extern "C" int pb_spectrum_reflected(pb_ctx *ctx, pb_optab *tab, const pb_spectrum_args *a)
{
    if (!ctx || !tab || !a) return PB_ERR_ARG;
    const int L = a->opacity.nlayer, W = a->nwno, G = a->numg * a->numt, V = L + 1;
    if (L < 1 || W < 1 || G < 1) return pb_fail(ctx, PB_ERR_ARG, "spectrum: bad sizes L=%d W=%d G=%d", L, W, G);
    if (a->opacity.ngauss > 1) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "spectrum: correlated-k tables (ngauss > 1) go through pb_compute_opacity + one flux call per gauss point");
    if (!a->albedo || !a->ubar0 || !a->ubar1 || !a->gweight || !a->tweight)
        return pb_fail(ctx, PB_ERR_ARG, "spectrum: albedo, ubar0, ubar1, gweight, tweight are required");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nLW = pb_align((size_t)L * W * sizeof(double)), nVW = pb_align((size_t)V * W * sizeof(double));
    const size_t nW = (size_t)W * sizeof(double);
    PB_TRY(aux_reserve(ctx, 9 * nLW + 2 * nVW + pb_align(nW) + pb_align((size_t)G * nW) + 256));
    size_t off = 0;
    auto take = [&](size_t bytes) -> double * {
        double *ptr = (double *)(ctx->aux + off);
        off += pb_align(bytes);
        return ptr;
    };
    pb_opacity_args o = a->opacity;
    o.DTAU = take(nLW); o.TAU = take(nVW); o.W0 = take(nLW); o.COSB = take(nLW); o.ftau_cld = take(nLW);
    o.ftau_ray = take(nLW); o.GCOS2 = take(nLW); o.DTAU_OG = take(nLW); o.TAU_OG = take(nVW); o.W0_OG = take(nLW);
    o.COSB_OG = take(nLW);
    o.W0_no_raman = nullptr; o.f_deltaM = nullptr; o.TAUGAS = o.TAURAY = o.TAUCLD = nullptr;
    double *d_alb = take(nW), *d_xint = a->xint_at_top ? take((size_t)G * nW) : nullptr;
    PB_TRY(pb_compute_opacity(ctx, tab, &o, PB_DEVICE));
    pb_reflected_args r;
    memset(&r, 0, sizeof(r));
    r.nlayer = L; r.nwno = W; r.numg = a->numg; r.numt = a->numt; r.nbatch = 1; r.ld = W;
    r.dtau = o.DTAU; r.tau = o.TAU; r.w0 = o.W0; r.cosb = o.COSB; r.gcos2 = o.GCOS2; r.ftau_cld = o.ftau_cld;
    r.ftau_ray = o.ftau_ray; r.dtau_og = o.DTAU_OG; r.tau_og = o.TAU_OG; r.w0_og = o.W0_OG; r.cosb_og = o.COSB_OG;
    r.surf_reflect = a->surf_reflect; r.F0PI = a->F0PI; r.b_top = a->b_top;
    r.ubar0 = a->ubar0; r.ubar1 = a->ubar1; r.gweight = a->gweight; r.tweight = a->tweight;
    r.cos_theta = a->cos_theta;
    r.single_phase = a->single_phase; r.multi_phase = a->multi_phase; r.toon_coefficients = a->toon_coefficients;
    r.frac_a = a->frac_a; r.frac_b = a->frac_b; r.frac_c = a->frac_c;
    r.constant_back = a->constant_back; r.constant_forward = a->constant_forward;
    r.get_toa_intensity = 1; r.get_lvl_flux = 0;
    r.albedo = d_alb; r.xint_at_top = d_xint;
    PB_TRY(pb_reflected_toon_1d(ctx, &r, PB_DEVICE));
    PB_CUDA(ctx, cudaMemcpyAsync(a->albedo, d_alb, nW, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_xint) PB_CUDA(ctx, cudaMemcpyAsync(a->xint_at_top, d_xint, (size_t)G * nW, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

extern "C" int pb_spectrum_thermal(pb_ctx *ctx, pb_optab *tab, const pb_spectrum_thermal_args *a)
{
    if (!ctx || !tab || !a) return PB_ERR_ARG;
    const int L = a->opacity.nlayer, W = a->nwno, G = a->numg * a->numt;
    if (L < 1 || W < 1 || G < 1) return pb_fail(ctx, PB_ERR_ARG, "spectrum_thermal: bad sizes L=%d W=%d G=%d", L, W, G);
    if (a->opacity.ngauss > 1) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "spectrum_thermal: correlated-k tables (ngauss > 1) go through pb_compute_opacity + one flux call per gauss point");
    if (!a->thermal || !a->tlevel || !a->plevel || !a->ubar1 || !a->gweight || !a->tweight || !a->wno)
        return pb_fail(ctx, PB_ERR_ARG, "spectrum_thermal: thermal, tlevel, plevel, ubar1, gweight, tweight, wno are required");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nLW = pb_align((size_t)L * W * sizeof(double)), nW = (size_t)W * sizeof(double);
    PB_TRY(aux_reserve(ctx, 3 * nLW + pb_align(nW) + pb_align((size_t)G * nW) + 256));
    size_t off = 0;
    auto take = [&](size_t bytes) -> double * {
        double *ptr = (double *)(ctx->aux + off);
        off += pb_align(bytes);
        return ptr;
    };
    pb_opacity_args o = a->opacity;
    o.DTAU = o.TAU = o.W0 = o.COSB = o.ftau_cld = o.ftau_ray = o.GCOS2 = o.TAU_OG = o.W0_OG = o.f_deltaM = nullptr;
    o.TAUGAS = o.TAURAY = o.TAUCLD = nullptr;
    o.DTAU_OG = take(nLW); o.W0_no_raman = take(nLW); o.COSB_OG = take(nLW);
    double *d_th = take(nW), *d_ft = a->flux_at_top ? take((size_t)G * nW) : nullptr;
    PB_TRY(pb_compute_opacity(ctx, tab, &o, PB_DEVICE));
    pb_thermal_args t;
    memset(&t, 0, sizeof(t));
    t.nlayer = L; t.nwno = W; t.numg = a->numg; t.numt = a->numt; t.nbatch = 1; t.ld = W;
    t.dtau = o.DTAU_OG; t.w0 = o.W0_no_raman; t.cosb = o.COSB_OG;   // justdoit.py:337-342
    t.wno = a->wno; t.dwno = nullptr; t.surf_reflect = a->surf_reflect;
    t.tlevel = a->tlevel; t.plevel = a->plevel;
    t.ubar1 = a->ubar1; t.gweight = a->gweight; t.tweight = a->tweight;
    t.hard_surface = a->hard_surface; t.calc_type = 0;
    t.thermal = d_th; t.flux_at_top = d_ft;
    PB_TRY(pb_thermal_toon_1d(ctx, &t, PB_DEVICE));
    PB_CUDA(ctx, cudaMemcpyAsync(a->thermal, d_th, nW, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_ft) PB_CUDA(ctx, cudaMemcpyAsync(a->flux_at_top, d_ft, (size_t)G * nW, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

extern "C" int pb_spectrum_transit(pb_ctx *ctx, pb_optab *tab, const pb_spectrum_transit_args *a)
{
    if (!ctx || !tab || !a) return PB_ERR_ARG;
    const int L = a->opacity.nlayer, W = a->nwno;
    if (L < 1 || W < 1) return pb_fail(ctx, PB_ERR_ARG, "spectrum_transit: bad sizes L=%d W=%d", L, W);
    if (a->opacity.ngauss > 1) return pb_fail(ctx, PB_ERR_UNSUPPORTED, "spectrum_transit: correlated-k tables (ngauss > 1) go through pb_compute_opacity + one call per gauss point");
    if (!a->F || !a->z || !a->dz || !a->player || !a->tlayer || !a->mmw || !a->colden)
        return pb_fail(ctx, PB_ERR_ARG, "spectrum_transit: F, z, dz, player, tlayer, mmw, colden are required");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nLW = pb_align((size_t)L * W * sizeof(double)), nW = (size_t)W * sizeof(double);
    PB_TRY(aux_reserve(ctx, nLW + pb_align(nW) + 256));
    pb_opacity_args o = a->opacity;
    o.DTAU = o.TAU = o.W0 = o.COSB = o.ftau_cld = o.ftau_ray = o.GCOS2 = o.TAU_OG = o.W0_OG = o.COSB_OG = nullptr;
    o.W0_no_raman = o.f_deltaM = o.TAUGAS = o.TAURAY = o.TAUCLD = nullptr;
    o.DTAU_OG = (double *)ctx->aux;
    double *d_F = (double *)(ctx->aux + nLW);
    PB_TRY(pb_compute_opacity(ctx, tab, &o, PB_DEVICE));
    pb_transit_args t;
    memset(&t, 0, sizeof(t));
    t.nlevel = L + 1; t.nwno = W; t.nbatch = 1; t.ld = W;
    t.DTAU = o.DTAU_OG;                                              // justdoit.py:392-396
    t.z = a->z; t.dz = a->dz; t.player = a->player; t.tlayer = a->tlayer; t.mmw = a->mmw; t.colden = a->colden;
    t.rstar = a->rstar; t.k_b = a->k_b; t.amu = a->amu;
    t.F = d_F;
    PB_TRY(pb_transit_1d(ctx, &t, PB_DEVICE));
    PB_CUDA(ctx, cudaMemcpyAsync(a->F, d_F, nW, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// ---- bound calls (numba nopython -> ctypes) -------------------------------------------------------------
namespace {
struct BoundCall { pb_ctx *ctx; pb_climate_args args; int memspace; bool live; };
std::vector<BoundCall> &bound_calls()
{
    static std::vector<BoundCall> v;
    return v;
}
} // namespace

extern "C" int pb_climate_bind(pb_ctx *ctx, const pb_climate_args *args, int memspace, int *handle_out)
{
    if (!ctx || !args || !handle_out) return pb_fail(ctx, PB_ERR_ARG, "climate_bind: bad arguments");
    auto &v = bound_calls();
    for (size_t i = 0; i < v.size(); ++i)
        if (!v[i].live) {
            v[i] = BoundCall{ctx, *args, memspace, true};
            *handle_out = (int)i;
            return PB_OK;
        }
    v.push_back(BoundCall{ctx, *args, memspace, true});
    *handle_out = (int)v.size() - 1;
    return PB_OK;
}

extern "C" int pb_climate_run_bound(unsigned long long ctx_address, int handle)
{
    auto &v = bound_calls();
    if (handle < 0 || (size_t)handle >= v.size() || !v[handle].live) return PB_ERR_ARG;
    BoundCall &b = v[handle];
    if ((unsigned long long)(uintptr_t)b.ctx != ctx_address) return PB_ERR_ARG;
    return pb_climate_get_fluxes(b.ctx, &b.args, b.memspace);
}

extern "C" int pb_climate_unbind(pb_ctx *ctx, int handle)
{
    auto &v = bound_calls();
    if (handle < 0 || (size_t)handle >= v.size() || !v[handle].live || v[handle].ctx != ctx)
        return pb_fail(ctx, PB_ERR_ARG, "climate_unbind: no such binding");
    v[handle].live = false;
    return PB_OK;
}
