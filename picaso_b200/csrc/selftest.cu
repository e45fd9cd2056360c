// selftest.cu - exposes the in-kernel math primitives (pb_math.cuh) so that the GPU test
// suite can check them against libm on arbitrary inputs.
#include "pb_common.cuh"
#include "pb_math.cuh"

namespace {
__global__ void math_kernel(int n, const double *x, double *e, double *r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    e[i] = pbm::exp(x[i]);
    r[i] = pbm::rcp(x[i]);
}
__global__ void exp_tab_kernel(int n, const double *x, double *e)
{
    __shared__ double tab[pbm::kExpTabDoubles];
    pbm::exp_tab_fill(tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) e[i] = pbm::exp_tab(x[i], tab + (threadIdx.x & 15));
}
} // namespace

extern "C" int pb_selftest_exp_tab(pb_ctx *ctx, const double *x, int n, double *exp_out)
{
    if (!ctx || !x || !exp_out || n < 0) return pb_fail(ctx, PB_ERR_ARG, "selftest_exp_tab: bad arguments");
    if (n == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)n * sizeof(double);
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, 2 * pb_align(nb) + 1024));
    const double *dx;
    int64_t ldo;
    PB_TRY(pb_stage_in(ctx, x, PB_HOST, 1, n, n, &dx, &ldo));
    double *de;
    PB_TRY(pb_arena_alloc(ctx, nb, (void **)&de));
    exp_tab_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, dx, de);
    PB_CHECK_LAUNCH(ctx);
    PB_CUDA(ctx, cudaMemcpyAsync(exp_out, de, nb, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

extern "C" int pb_selftest_math(pb_ctx *ctx, const double *x, int n, double *exp_out, double *rcp_out)
{
    if (!ctx || !x || !exp_out || !rcp_out || n < 0) return pb_fail(ctx, PB_ERR_ARG, "selftest_math: bad arguments");
    if (n == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)n * sizeof(double);
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, 3 * pb_align(nb) + 1024));
    const double *dx;
    int64_t ldo;
    PB_TRY(pb_stage_in(ctx, x, PB_HOST, 1, n, n, &dx, &ldo));
    double *de, *dr;
    PB_TRY(pb_arena_alloc(ctx, nb, (void **)&de));
    PB_TRY(pb_arena_alloc(ctx, nb, (void **)&dr));
    math_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, dx, de, dr);
    PB_CHECK_LAUNCH(ctx);
    PB_CUDA(ctx, cudaMemcpyAsync(exp_out, de, nb, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaMemcpyAsync(rcp_out, dr, nb, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// ---------------------------------------------------------------------------------------
// pb_microbench: the machine numbers the fp64 flux kernels are designed against (second roof of
// bench.py's roofline block: the HBM copy peak is in MEASURED_PEAKS.json, there is no fp64 figure).
//   which = 0  DFMA throughput: 8 independent chains per thread, every SM full
//           1  DFMA dependent-chain latency: one warp per CTA, one chain
//           2  DFMA throughput with only lanes 0..15 of every warp active
//           3  LDS.64 throughput (conflict-free), every SM full
//           4  MUFU.RCP64H + 2 Newton steps (pbm::rcp), dependent chain latency
//           5  DFMA throughput at 3 warps per SM sub-partition (the occupancy of the headline launch), ILP 2
// out[0] = elapsed ms (CUDA events), out[1] = instructions (or loads) per thread,
// out[2] = threads launched, out[3] = SM cycles per iteration seen by thread 0 (clock64).
// ---------------------------------------------------------------------------------------
namespace {
template <int ILP, bool HALF>
__global__ void mb_dfma_kernel(int iters, double seed, double *sink, long long *cyc)
{
    double a[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = seed + k + threadIdx.x * 1e-3;
    const double m = 1.0 - 1e-9, c = 1e-9;
    const bool on = !HALF || (threadIdx.x & 31) < 16;
    const long long t0 = clock64();
    if (on) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], m, c);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += a[k];
    if (s == 123.456) sink[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void mb_lds_kernel(int iters, double *sink, long long *cyc)
{
    __shared__ double buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i;
    __syncthreads();
    double s = 0.0;
    int idx = threadIdx.x & 1023;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s += buf[(idx + k * 256) & 4095];
        idx = (idx + 32) & 1023;
    }
    const long long t1 = clock64();
    if (s == 123.456) sink[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void mb_rcp_kernel(int iters, double seed, double *sink, long long *cyc)
{
    double a = seed + threadIdx.x * 1e-3;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = pbm::rcp(a) + 0.5;
    const long long t1 = clock64();
    if (a == 123.456) sink[0] = a;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
} // namespace

extern "C" int pb_microbench(pb_ctx *ctx, int which, int iters, double *out)
{
    if (!ctx || !out || iters < 1 || which < 0 || which > 5) return pb_fail(ctx, PB_ERR_ARG, "microbench: bad arguments");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_arena_reset(ctx);
    PB_TRY(pb_arena_reserve(ctx, 1024));
    double *sink;
    long long *cyc;
    PB_TRY(pb_arena_alloc(ctx, 256, (void **)&sink));
    PB_TRY(pb_arena_alloc(ctx, 256, (void **)&cyc));
    const int nsm = ctx->sm_count > 0 ? ctx->sm_count : 148;
    double per_thread = 0.0, threads = 0.0;
    float ms = 0.f;
    for (int rep = 0; rep < 2; ++rep) {  // first pass warms the clocks
        PB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
        switch (which) {
        case 0: mb_dfma_kernel<8, false><<<nsm * 2, 1024, 0, ctx->stream>>>(iters, 1.0, sink, cyc);
                per_thread = 8.0 * iters; threads = nsm * 2048.0; break;
        case 1: mb_dfma_kernel<1, false><<<nsm, 32, 0, ctx->stream>>>(iters, 1.0, sink, cyc);
                per_thread = iters; threads = nsm * 32.0; break;
        case 2: mb_dfma_kernel<8, true><<<nsm * 2, 1024, 0, ctx->stream>>>(iters, 1.0, sink, cyc);
                per_thread = 8.0 * iters; threads = nsm * 2048.0; break;
        case 3: mb_lds_kernel<<<nsm * 2, 1024, 0, ctx->stream>>>(iters, sink, cyc);
                per_thread = 8.0 * iters; threads = nsm * 2048.0; break;
        case 4: mb_rcp_kernel<<<nsm, 32, 0, ctx->stream>>>(iters, 1.5, sink, cyc);
                per_thread = iters; threads = nsm * 32.0; break;
        default: mb_dfma_kernel<2, false><<<nsm, 384, 0, ctx->stream>>>(iters, 1.0, sink, cyc);
                per_thread = 2.0 * iters; threads = nsm * 384.0; break;
        }
        PB_CHECK_LAUNCH(ctx);
        PB_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
        PB_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
        PB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
    }
    long long h = 0;
    PB_CUDA(ctx, cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = ms; out[1] = per_thread; out[2] = threads; out[3] = (double)h / iters;
    return PB_OK;
}
