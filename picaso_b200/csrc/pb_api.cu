// pb_api.cu - context, memory, timing and staging entry points of libpicaso_b200.so.
#include "pb_common.cuh"

static char g_create_err[512] = "";

int pb_fail(pb_ctx *ctx, int code, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" {

int pb_version(void) { return 100; }

int pb_device_count(int *count)
{
    if (!count) return PB_ERR_ARG;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return pb_fail(nullptr, PB_ERR_CUDA, "cudaGetDeviceCount -> %s", cudaGetErrorString(e));
    }
    return PB_OK;
}

int pb_create(int device, pb_ctx **out)
{
    if (!out) return PB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return pb_fail(nullptr, PB_ERR_CUDA,
                       "no CUDA device available (%s); picaso_b200 has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n)
        return pb_fail(nullptr, PB_ERR_ARG, "device %d out of range [0,%d)", device, n);
    pb_ctx *ctx = new pb_ctx();
    ctx->device = device;
#define CREATE_CUDA(call)                                                                    \
    do {                                                                                     \
        cudaError_t e2 = (call);                                                             \
        if (e2 != cudaSuccess) {                                                             \
            pb_fail(nullptr, PB_ERR_CUDA, "pb_create: %s -> %s", #call, cudaGetErrorString(e2)); \
            delete ctx;                                                                      \
            return PB_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)
    CREATE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaEventCreate(&ctx->ev_start));
    CREATE_CUDA(cudaEventCreate(&ctx->ev_stop));
    CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    for (auto &e : ctx->ev_chunk) CREATE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
#undef CREATE_CUDA
    *out = ctx;
    return PB_OK;
}

void pb_destroy(pb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->aux) cudaFree(ctx->aux);
    if (ctx->rec) cudaFree(ctx->rec);
    if (ctx->rec2) cudaFree(ctx->rec2);
    for (auto &sl : ctx->pin) {
        if (sl.host) cudaFreeHost(sl.host);
        if (sl.dev) cudaFree(sl.dev);
        if (sl.ev) cudaEventDestroy(sl.ev);
        free(sl.shadow);
    }
    cudaEventDestroy(ctx->ev_start);
    cudaEventDestroy(ctx->ev_stop);
    cudaEventDestroy(ctx->ev_copy);
    for (auto &e : ctx->ev_chunk)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->own_stream);
    cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

const char *pb_last_error(const pb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int pb_device_name(pb_ctx *ctx, char *buf, size_t buflen)
{
    if (!ctx || !buf || !buflen) return PB_ERR_ARG;
    cudaDeviceProp prop;
    PB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    snprintf(buf, buflen, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor,
             prop.multiProcessorCount);
    return PB_OK;
}

int pb_sm_count(pb_ctx *ctx, int *count)
{
    if (!ctx || !count) return PB_ERR_ARG;
    *count = ctx->sm_count;
    return PB_OK;
}

int pb_dev_alloc(pb_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess)
        return pb_fail(ctx, PB_ERR_NOMEM, "cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    return PB_OK;
}

int pb_dev_free(pb_ctx *ctx, void *ptr)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaFree(ptr));
    return PB_OK;
}

int pb_host_alloc(pb_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess)
        return pb_fail(ctx, PB_ERR_NOMEM, "cudaHostAlloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    return PB_OK;
}

int pb_host_free(pb_ctx *ctx, void *ptr)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_CUDA(ctx, cudaFreeHost(ptr));
    return PB_OK;
}

int pb_memcpy_h2d(pb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PB_OK;
}

int pb_memcpy_d2h(pb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

int pb_memset(pb_ctx *ctx, void *dst, int value, size_t bytes)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return PB_OK;
}

int pb_sync(pb_ctx *ctx)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_set_stream(pb_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PB_OK;
}

int pb_timer_start(pb_ctx *ctx)
{
    if (!ctx) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return PB_OK;
}

int pb_timer_stop(pb_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    PB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
    return PB_OK;
}

uint64_t pb_launch_count(const pb_ctx *ctx) { return ctx ? ctx->launches : 0; }

} // extern "C"

// ---- arena / staging -----------------------------------------------------------------

void pb_arena_reset(pb_ctx *ctx)
{
    ctx->arena_off = 0;
    // next pinned slot; wait (normally a no-op) until the copy that last used it has finished
    ctx->pin_cur = (ctx->pin_cur + 1) % pb_ctx::kPinSlots;
    pb_ctx::PinSlot &sl = ctx->pin[ctx->pin_cur];
    if (sl.pending) {
        cudaEventSynchronize(sl.ev);
        sl.pending = false;
    }
    ctx->pin_off = 0;
    ctx->pin_flushed = 0;
}

int pb_arena_reserve(pb_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->arena_cap) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) PB_CUDA(ctx, cudaFree(ctx->arena));
    ctx->arena = nullptr;
    ctx->arena_cap = 0;
    size_t cap = pb_align(bytes + bytes / 8, 1 << 20);
    cudaError_t e = cudaMalloc((void **)&ctx->arena, cap);
    if (e != cudaSuccess)
        return pb_fail(ctx, PB_ERR_NOMEM, "staging arena cudaMalloc(%zu) -> %s", cap,
                       cudaGetErrorString(e));
    ctx->arena_cap = cap;
    return PB_OK;
}

int pb_arena_alloc(pb_ctx *ctx, size_t bytes, void **out)
{
    size_t off = pb_align(ctx->arena_off);
    if (off + bytes > ctx->arena_cap)
        return pb_fail(ctx, PB_ERR_NOMEM, "staging arena overflow: need %zu, cap %zu (reserve bug)",
                       off + bytes, ctx->arena_cap);
    *out = ctx->arena + off;
    ctx->arena_off = off + bytes;
    return PB_OK;
}

int pb_pinned_reserve(pb_ctx *ctx, size_t bytes)
{
    pb_ctx::PinSlot &cur = ctx->pin[ctx->pin_cur];
    if (!cur.ev) PB_CUDA(ctx, cudaEventCreateWithFlags(&cur.ev, cudaEventDisableTiming));
    if (bytes <= cur.cap) return PB_OK;
    // Grow EVERY slot of the ring now: a caller that needs this much will need it on its next calls too, and
    // paying cudaHostAlloc once per slot over the first kPinSlots calls made those calls 3-4x slower
    // (128-atmosphere batches: 3.8 ms instead of 0.65 ms per launch until the ring had turned once).
    PB_CUDA(ctx, cudaDeviceSynchronize());  // kernels / copies on any stream may still use the old blocks
    const size_t cap = pb_align(bytes * 2, 1 << 16);
    for (auto &sl : ctx->pin) {
        if (sl.cap >= cap) continue;
        if (!sl.ev) PB_CUDA(ctx, cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
        sl.pending = false;
        if (sl.host) PB_CUDA(ctx, cudaFreeHost(sl.host));
        if (sl.dev) PB_CUDA(ctx, cudaFree(sl.dev));
        free(sl.shadow);
        sl.host = sl.dev = sl.shadow = nullptr;
        sl.cap = 0;
        sl.shadow_valid = 0;
        cudaError_t e = cudaHostAlloc((void **)&sl.host, cap, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc((void **)&sl.dev, cap);
        if (e != cudaSuccess)
            return pb_fail(ctx, PB_ERR_NOMEM, "pinned slot allocation (%zu bytes) -> %s", cap, cudaGetErrorString(e));
        sl.shadow = (char *)malloc(cap < pb_ctx::kShadowMax ? cap : pb_ctx::kShadowMax);
        sl.cap = cap;
    }
    return PB_OK;
}

int pb_upload_small(pb_ctx *ctx, const double *host, size_t n, const double **dev_out)
{
    pb_ctx::PinSlot &sl = ctx->pin[ctx->pin_cur];
    const size_t bytes = n * sizeof(double);
    const size_t off = pb_align(ctx->pin_off, 64);
    if (off + bytes > sl.cap) return pb_fail(ctx, PB_ERR_NOMEM, "pinned slot overflow (reserve bug)");
    memcpy(sl.host + off, host, bytes);
    ctx->pin_off = off + bytes;
    *dev_out = (const double *)(sl.dev + off);
    return PB_OK;
}

int pb_upload_flush(pb_ctx *ctx)
{
    pb_ctx::PinSlot &sl = ctx->pin[ctx->pin_cur];
    if (ctx->pin_off > ctx->pin_flushed) {
        const size_t a = ctx->pin_flushed, b = ctx->pin_off;
        static const bool no_shadow = getenv("PB_NO_UPLOAD_SHADOW") != nullptr;  // A/B switch
        const bool mirrored = sl.shadow && b <= pb_ctx::kShadowMax && !no_shadow;
        // the device block already holds exactly these bytes (copied by an earlier call on this stream): nothing to do
        if (mirrored && b <= sl.shadow_valid && sl.shadow_stream == ctx->stream && memcmp(sl.shadow + a, sl.host + a, b - a) == 0) {
            ctx->pin_flushed = b;
            return PB_OK;
        }
        PB_CUDA(ctx, cudaMemcpyAsync(sl.dev + a, sl.host + a, b - a, cudaMemcpyHostToDevice, ctx->stream));
        PB_CUDA(ctx, cudaEventRecord(sl.ev, ctx->stream));
        sl.pending = true;
        if (mirrored && (a <= sl.shadow_valid) && (sl.shadow_valid == 0 || sl.shadow_stream == ctx->stream)) {
            memcpy(sl.shadow + a, sl.host + a, b - a);
            if (b > sl.shadow_valid) sl.shadow_valid = b;
            sl.shadow_stream = ctx->stream;
        } else if (a < sl.shadow_valid) {
            sl.shadow_valid = a;   // the mirror no longer describes the bytes from `a` on
        }
        ctx->pin_flushed = b;
    }
    return PB_OK;
}

int pb_stage_in(pb_ctx *ctx, const double *src, int memspace, int64_t rows, int64_t width,
                int64_t ld, const double **dev_out, int64_t *ld_out)
{
    if (!src) {
        *dev_out = nullptr;
        if (ld_out) *ld_out = width;
        return PB_OK;
    }
    if (memspace == PB_DEVICE) {
        *dev_out = src;
        if (ld_out) *ld_out = ld;
        return PB_OK;
    }
    void *d = nullptr;
    PB_TRY(pb_arena_alloc(ctx, (size_t)rows * width * sizeof(double), &d));
    if (ld == width || rows == 1) {
        PB_CUDA(ctx, cudaMemcpyAsync(d, src, (size_t)rows * width * sizeof(double),
                                     cudaMemcpyHostToDevice, ctx->stream));
    } else {
        PB_CUDA(ctx, cudaMemcpy2DAsync(d, width * sizeof(double), src, ld * sizeof(double),
                                       width * sizeof(double), rows, cudaMemcpyHostToDevice,
                                       ctx->stream));
    }
    *dev_out = (const double *)d;
    if (ld_out) *ld_out = width;
    return PB_OK;
}

// ---- peer memory for the fused all-gather (include/picaso_b200.h: pb_peer_gather) ----
namespace {
__global__ void gather_wait_kernel(const unsigned long long *flags, int n, unsigned long long step, int *timed_out)
{
    const int r = threadIdx.x;
    if (r >= n) return;
    const long long t0 = clock64();
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
        if (v >= step) break;
        if (clock64() - t0 > 4000000000LL) {
            if (timed_out) *timed_out = 1;
            break;
        }
    } while (true);
}
// rank `rank` stores `value` into word offset + rank of every rank's flag array (release, system scope); the
// preceding system-scope fence makes everything this GPU wrote before (earlier kernels of the stream) visible first
struct PeerFlagTab { unsigned long long *p[32]; };
__global__ void peer_signal_kernel(PeerFlagTab tab, int n, int rank, int offset, unsigned long long value)
{
    if (threadIdx.x >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(tab.p[threadIdx.x] + offset + rank), "l"(value) : "memory");
}
} // namespace

int pb_ipc_export(pb_ctx *ctx, void *dev_ptr, void *handle64)
{
    if (!ctx || !dev_ptr || !handle64) return PB_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    PB_CUDA(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, sizeof(h));
    return PB_OK;
}

int pb_ipc_open(pb_ctx *ctx, const void *handle64, void **dev_ptr)
{
    if (!ctx || !dev_ptr || !handle64) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    PB_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PB_OK;
}

int pb_ipc_close(pb_ctx *ctx, void *dev_ptr)
{
    if (!ctx || !dev_ptr) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return PB_OK;
}

int pb_gather_wait(pb_ctx *ctx, const unsigned long long *flags, int nranks, unsigned long long step, int *timed_out_dev)
{
    if (!ctx || !flags || nranks < 1 || nranks > 32) return PB_ERR_ARG;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    gather_wait_kernel<<<1, 32, 0, ctx->stream>>>(flags, nranks, step, timed_out_dev);
    PB_CHECK_LAUNCH(ctx);
    return PB_OK;
}

extern "C" int pb_peer_signal(pb_ctx *ctx, unsigned long long *const *flags, int nranks, int rank, int offset,
                              unsigned long long value)
{
    if (!ctx || !flags || nranks < 1 || nranks > 32 || rank < 0 || rank >= nranks || offset < 0)
        return pb_fail(ctx, PB_ERR_ARG, "peer_signal: bad arguments");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PeerFlagTab tab;  // the pointer table is a host array: it travels in the launch parameters
    for (int r = 0; r < 32; ++r) tab.p[r] = r < nranks ? flags[r] : nullptr;
    peer_signal_kernel<<<1, 32, 0, ctx->stream>>>(tab, nranks, rank, offset, value);
    PB_CHECK_LAUNCH(ctx);
    return PB_OK;
}
