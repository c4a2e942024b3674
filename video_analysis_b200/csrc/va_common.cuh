// va_common.cuh -- shared declarations of libva_b200 (sm_100a).
#pragma once

#ifdef VA_EMU
#include "cuda_emu.h"   // tests/emu: development-time CPU shim, never shipped
#else
#include <cuda_runtime.h>
#endif

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "va_b200.h"

// ---------------------------------------------------------------------------------
// ctx: scratch only.  All image buffers belong to the caller.
// ---------------------------------------------------------------------------------
struct va_ctx {
    int device;
    int sm_count;
    int max_w, max_h, max_batch;
    long long launches;
    char err[512];

    // labelling scratch
    int32_t *lab_parent;     // [max_batch * max_h * lab_pitch]  union-find forest, sparse (run starts only)
    size_t lab_pitch;        // elements per row: power of two >= max_w
    int32_t *lab_rowcnt;     // [2][max_batch * max_h]  roots per row -> exclusive prefix; rows' foreground flags
    int32_t *lab_parent1, *lab_rowcnt1;   // second scratch set (slot 1 of va_label_forest / va_label_write), allocated on first use
    // last use of each scratch set: every entry point that touches a set first makes its stream wait for
    // this event and re-records it when it has enqueued its kernels, so callers on different streams
    // (two filter chains, a chain next to region_stats, ...) never run forest kernels on one set at once
    void *lab_event[5];      // [2]: the chain intermediates of va_chain_run, [3]: the row offsets of va_label_export_chunks
    int *exp_rowoff;         // [max_batch * max_h] chunk counts per row -> exclusive prefix (allocated on first use)
    void *rs_tab;            // resize coefficient tables of the launch in flight (event [4] orders its users)
    size_t rs_tab_bytes;
    // morphology scratch (intermediate of open / close is kept in shared memory; none needed)
    // chain intermediates (allocated on first use by va_chain_run)
    uint8_t *ch_mono, *ch_blur;
    uint32_t *ch_mask, *ch_morph;
    size_t ch_pitch, ch_pitch_w;
};

#ifdef VA_EMU
static inline int va_scratch_acquire(va_ctx *, va_stream, int) { return 0; }
static inline int va_scratch_release(va_ctx *, va_stream, int) { return 0; }
#else
static inline int va_scratch_acquire(va_ctx *ctx, va_stream stream, int slot) {
    if (!ctx->lab_event[slot]) return 0;
    return cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ctx->lab_event[slot], 0) == cudaSuccess ? 0 : -1;
}
static inline int va_scratch_release(va_ctx *ctx, va_stream stream, int slot) {
    if (!ctx->lab_event[slot]) {
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return -1;
        ctx->lab_event[slot] = (void *)ev;
    }
    return cudaEventRecord((cudaEvent_t)ctx->lab_event[slot], (cudaStream_t)stream) == cudaSuccess ? 0 : -1;
}
#endif

#define VA_SET_ERR(ctx, ...)                                              \
    do {                                                                  \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);   \
    } while (0)

#define VA_FAIL(ctx, code, ...)        \
    do {                               \
        VA_SET_ERR(ctx, __VA_ARGS__);  \
        return (code);                 \
    } while (0)

#define VA_CUDA(ctx, call)                                                                \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            VA_FAIL(ctx, VA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));   \
    } while (0)

#ifdef VA_EMU
#define VA_LAUNCH(ctx, kfn, grid, block, smem, stream, ...)              \
    do {                                                                 \
        emu::launch(kfn, dim3(grid), dim3(block), smem, __VA_ARGS__);    \
        (ctx)->launches++;                                               \
    } while (0)
#define VA_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(emu::dyn_smem_ptr)
#else
#define VA_LAUNCH(ctx, kfn, grid, block, smem, stream, ...)                           \
    do {                                                                              \
        kfn<<<dim3(grid), dim3(block), smem, (cudaStream_t)(stream)>>>(__VA_ARGS__);  \
        (ctx)->launches++;                                                            \
        VA_CUDA(ctx, cudaGetLastError());                                             \
    } while (0)
#define VA_DYN_SMEM(T, name)                                        \
    extern __shared__ __align__(16) unsigned char name##_raw__[];   \
    T *name = reinterpret_cast<T *>(name##_raw__)
#endif

#define VA_CHECK_CTX(ctx)                  \
    do {                                   \
        if (!(ctx)) return VA_ERR_INVALID; \
        (ctx)->err[0] = 0;                 \
    } while (0)

#define VA_REQUIRE(ctx, cond, ...)                              \
    do {                                                        \
        if (!(cond)) VA_FAIL(ctx, VA_ERR_INVALID, __VA_ARGS__); \
    } while (0)

static inline int va_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline bool va_aligned(const void *p, size_t a) { return ((uintptr_t)p % a) == 0; }

// persistent-style grid: a multiple of the SM count, capped by the work available
static inline int va_grid(const va_ctx *ctx, long long work_items, int ctas_per_sm) {
    long long g = (long long)ctx->sm_count * ctas_per_sm;
    if (g > work_items) g = work_items;
    if (g < 1) g = 1;
    return (int)g;
}

// internal entry points shared between translation units
int va_gauss_build_taps(double sigma, int *taps, int capacity);
