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
} // namespace

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
