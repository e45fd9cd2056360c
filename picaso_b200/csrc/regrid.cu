// regrid.cu - bin-mean rebinning of spectra onto a data grid, sm_100a.
//
// Replaces picaso/justplotit.py:31-63 (mean_regrid), i.e. scipy.stats.binned_statistic(x, y, 'mean',
// bins=edges): mean of y over the samples whose x falls in [edge_i, edge_i+1) (last bin closed).
// The retrieval driver (picaso/driver.py:176-245) applies it to every model spectrum, always with the
// same model grid x and data grid - so the bin membership is planned once on the host
// (picaso_b200/regrid.py: RegridPlan, np.digitize + scipy's right-edge rule) and lives in HBM as one
// (start, count) pair per bin: for a monotonic x every bin is a contiguous index range.
// One thread per (bin, spectrum) adds its range in index order - the order np.bincount uses - so the
// result is bit-identical to scipy's for finite inputs; empty bins give NaN like scipy.
#include "pb_common.cuh"

namespace {

__global__ void mean_regrid_kernel(int nbins, int64_t ld, int64_t ld_out, const int *__restrict__ start,
                                   const int *__restrict__ count, const double *__restrict__ y, double scale,
                                   double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= nbins) return;
    const int s = start[i], n = count[i];
    const double *row = y + (int64_t)b * ld;
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(scale, row[s + k]));  // no FMA: scipy multiplies, rounds, then adds
    out[(int64_t)b * ld_out + i] = n > 0 ? acc / (double)n : __longlong_as_double(0x7ff8000000000000LL);
}

} // namespace

extern "C" int pb_regrid_plan_create(pb_ctx *ctx, int nbins, const int *start, const int *count, pb_regrid_plan **out)
{
    if (!ctx || !out || nbins < 0 || (nbins > 0 && (!start || !count))) return pb_fail(ctx, PB_ERR_ARG, "regrid plan: bad arguments");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_regrid_plan *p = new pb_regrid_plan();
    p->nbins = nbins;
    p->max_index = 0;
    for (int i = 0; i < nbins; ++i) {
        if (start[i] < 0 || count[i] < 0) {
            delete p;
            return pb_fail(ctx, PB_ERR_ARG, "regrid plan: negative range in bin %d", i);
        }
        if (count[i] > 0 && start[i] + count[i] > p->max_index) p->max_index = start[i] + count[i];
    }
    const size_t nb = (size_t)(nbins > 0 ? nbins : 1) * sizeof(int);
    cudaError_t e = cudaMalloc((void **)&p->start, 2 * nb);
    if (e != cudaSuccess) {
        delete p;
        return pb_fail(ctx, PB_ERR_NOMEM, "regrid plan cudaMalloc -> %s", cudaGetErrorString(e));
    }
    p->count = p->start + (nbins > 0 ? nbins : 1);
    if (nbins > 0) {
        PB_CUDA(ctx, cudaMemcpyAsync(p->start, start, (size_t)nbins * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaMemcpyAsync(p->count, count, (size_t)nbins * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out = p;
    return PB_OK;
}

extern "C" int pb_regrid_plan_destroy(pb_ctx *ctx, pb_regrid_plan *plan)
{
    if (!ctx || !plan) return PB_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (plan->start) cudaFree(plan->start);
    delete plan;
    return PB_OK;
}

extern "C" int pb_mean_regrid(pb_ctx *ctx, const pb_regrid_plan *plan, int nbatch, int nwno, int64_t ld,
                              const double *y, double scale, double *out, int memspace)
{
    if (!ctx || !plan || !y || !out || nbatch < 0 || nwno < 0 || ld < nwno)
        return pb_fail(ctx, PB_ERR_ARG, "mean_regrid: bad arguments");
    if (plan->max_index > nwno) return pb_fail(ctx, PB_ERR_ARG, "mean_regrid: plan reaches index %d, spectrum has %d points", plan->max_index, nwno);
    if (nbatch > 65535) return pb_fail(ctx, PB_ERR_ARG, "mean_regrid: nbatch = %d exceeds 65535 (spectra map to gridDim.y); split the batch", nbatch);
    if (nbatch == 0 || plan->nbins == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = memspace == PB_HOST;
    const double *d_y = y;
    double *d_out = out;
    int64_t ldd = ld;
    if (host) {
        pb_arena_reset(ctx);
        PB_TRY(pb_arena_reserve(ctx, 4 * 256 + pb_align((size_t)nbatch * nwno * 8) + pb_align((size_t)nbatch * plan->nbins * 8)));
        PB_TRY(pb_stage_in(ctx, y, memspace, nbatch, nwno, ld, &d_y, &ldd));
        PB_TRY(pb_arena_alloc(ctx, (size_t)nbatch * plan->nbins * 8, (void **)&d_out));
    }
    dim3 grid((plan->nbins + 127) / 128, nbatch);
    mean_regrid_kernel<<<grid, 128, 0, ctx->stream>>>(plan->nbins, ldd, plan->nbins, plan->start, plan->count, d_y, scale, d_out);
    PB_CHECK_LAUNCH(ctx);
    if (host) {
        PB_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)nbatch * plan->nbins * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}
