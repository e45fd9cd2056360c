// pb_common.cuh - context object, error plumbing and staging helpers shared by the
// kernels of libpicaso_b200.so.  Host-side runtime only; no kernel lives here.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/picaso_b200.h"

#define PB_PI 3.14159265358979323846

struct pb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;   // every kernel of this context is launched here
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_copy = nullptr;
    static constexpr int kChunkEvents = 16;
    cudaEvent_t ev_chunk[kChunkEvents] = {};  // copy-stream -> compute-stream hand-off per wavelength chunk
    bool push_pending[8] = {};               // pb_peer_gather push mode: ev_chunk[slot] has been recorded
    uint64_t launches = 0;
    char err[512] = {0};
    // kernels whose dynamic shared-memory limit this context has already raised (pb_ensure_smem)
    struct SmemAttr { const void *fn; int bytes; } smem_attr[24] = {};
    // grow-only device arena used to stage PB_HOST calls and small geometry vectors
    char *arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    // second grow-only block for entry points that call other entry points (climate.cu): survives arena resets
    char *aux = nullptr;
    size_t aux_cap = 0;
    // per-layer records of the level-flux kernels (toon_thermal.cu: therm_layer_records_kernel)
    char *rec = nullptr;
    size_t rec_cap = 0;
    char *rec2 = nullptr;   // toon_reflected.cu: refl_layer_records_kernel
    size_t rec2_cap = 0;
    // grow-only pinned bounce buffer for small host vectors (geometry) so that their
    // H2D copies are truly asynchronous
    // H2D copies are truly asynchronous and ONE copy per API call carries all of them.  A ring
    // of slots (each with an event recorded after its copy) keeps a later asynchronous call
    // from overwriting host bytes a pending copy still reads.
    struct PinSlot {
        char *host = nullptr, *dev = nullptr;
        size_t cap = 0;
        cudaEvent_t ev = nullptr;
        bool pending = false;
        // Host mirror of the first `shadow_valid` bytes of `dev` (what earlier flushes of this slot copied there, on
        // `shadow_stream`).  A flush whose bytes equal the mirror issues NO copy: a solver called in a loop with the same
        // geometry (ubar0/ubar1/gweight/tweight) stops paying an in-stream H2D copy + event per call once the ring
        // has turned (pb_upload_flush).
        char *shadow = nullptr;
        size_t shadow_valid = 0;
        cudaStream_t shadow_stream = nullptr;
    };
    static constexpr size_t kShadowMax = 16 * 1024;  // larger blobs (batched level vectors) are not mirrored
    static constexpr int kPinSlots = 8;
    PinSlot pin[kPinSlots];
    int pin_cur = 0;
    size_t pin_off = 0, pin_flushed = 0;
};

struct pb_regrid_plan {
    int nbins = 0, max_index = 0;
    int *start = nullptr, *count = nullptr;  // device, one allocation
};

int pb_fail(pb_ctx *ctx, int code, const char *fmt, ...);

#define PB_CUDA(ctx, call)                                                                  \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return pb_fail((ctx), PB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, \
                           cudaGetErrorString(e__));                                        \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (context, kernel, size): it is a driver call
// that does not belong on the per-launch path of a solver called in a loop
template <typename K>
inline cudaError_t pb_ensure_smem(pb_ctx *ctx, K kern, size_t bytes)
{
    const void *fn = (const void *)kern;
    for (auto &e : ctx->smem_attr) {
        if (e.fn == fn) {
            if (e.bytes >= (int)bytes) return cudaSuccess;
            cudaError_t rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (rc == cudaSuccess) e.bytes = (int)bytes;
            return rc;
        }
        if (!e.fn) {
            cudaError_t rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (rc == cudaSuccess) { e.fn = fn; e.bytes = (int)bytes; }
            return rc;
        }
    }
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);  // table full
}

#define PB_CHECK_LAUNCH(ctx)                       \
    do {                                           \
        (ctx)->launches++;                         \
        PB_CUDA((ctx), cudaGetLastError());        \
    } while (0)

#define PB_TRY(expr)                   \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != PB_OK) return rc__; \
    } while (0)

// Arena: reset at the start of every API call, bump-allocated, 256-B aligned.  Growing
// synchronises the stream first (earlier kernels may still use the old block).
int pb_arena_reserve(pb_ctx *ctx, size_t bytes);
void pb_arena_reset(pb_ctx *ctx);
int pb_arena_alloc(pb_ctx *ctx, size_t bytes, void **out);
// stage a small host vector in the current pinned slot; the returned device pointer becomes
// valid after pb_upload_flush (which every entry point calls before its first launch)
int pb_upload_small(pb_ctx *ctx, const double *host, size_t n, const double **dev_out);
int pb_upload_flush(pb_ctx *ctx);
// reserve pinned bounce space (call before the first pb_upload_small of an API call)
int pb_pinned_reserve(pb_ctx *ctx, size_t bytes);

// 2-D staged copies of [rows][ld] -> dense [rows][width] device blocks
int pb_stage_in(pb_ctx *ctx, const double *src, int memspace, int64_t rows, int64_t width,
                int64_t ld, const double **dev_out, int64_t *ld_out);

static inline size_t pb_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
