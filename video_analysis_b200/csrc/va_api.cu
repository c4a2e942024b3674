// va_api.cu -- ctx management, error strings and the fused-sequence entry point.
#include <cstdlib>

#include "va_common.cuh"

extern "C" int va_version(void) { return 100; }   // 0.1.0

extern "C" const char *va_status_string(int status) {
    switch (status) {
        case VA_OK: return "ok";
        case VA_ERR_INVALID: return "invalid argument";
        case VA_ERR_CUDA: return "CUDA runtime error";
        case VA_ERR_NOMEM: return "out of device memory";
        case VA_ERR_CAPACITY: return "exceeds ctx capacity";
        case VA_ERR_UNSUPPORTED: return "not supported on the device path";
        default: return "unknown status";
    }
}

extern "C" const char *va_last_error(const va_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }
extern "C" long long va_launch_count(const va_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int va_create(va_ctx **out, int device, int max_w, int max_h, int max_batch) {
    if (!out) return VA_ERR_INVALID;
    *out = nullptr;
    if (max_w <= 0 || max_h <= 0 || max_batch <= 0) return VA_ERR_INVALID;
    va_ctx *ctx = (va_ctx *)calloc(1, sizeof(va_ctx));
    if (!ctx) return VA_ERR_NOMEM;
    ctx->device = device;
    ctx->max_w = max_w; ctx->max_h = max_h; ctx->max_batch = max_batch;
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return VA_ERR_CUDA; }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) {
        free(ctx);
        return VA_ERR_CUDA;
    }
    ctx->sm_count = sms;
#ifndef VA_EMU
    {   // stream-ordered temporaries (cudaMallocAsync in the blur / resize entry points) stay in the pool across
        // synchronisations instead of going back to the driver every time the stream runs dry
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
#endif
    ctx->lab_pitch = 32;                       // power of two: forest index = (y << log2 pitch) + x
    while (ctx->lab_pitch < (size_t)max_w) ctx->lab_pitch <<= 1;
    if (ctx->lab_pitch * (size_t)max_h >= ((size_t)1 << 31)) { free(ctx); return VA_ERR_CAPACITY; }
    const size_t n_parent = ctx->lab_pitch * (size_t)max_h * (size_t)max_batch;
    if (cudaMalloc((void **)&ctx->lab_parent, n_parent * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc((void **)&ctx->lab_rowcnt, (size_t)2 * max_h * max_batch * sizeof(int32_t)) != cudaSuccess) {
        cudaGetLastError();
        va_destroy(ctx);
        return VA_ERR_NOMEM;
    }
    *out = ctx;
    return VA_OK;
}

extern "C" int va_reserve(va_ctx *ctx, int max_w, int max_h, int max_batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, max_w > 0 && max_h > 0 && max_batch > 0, "va_reserve: bad size");
    const int w = max_w > ctx->max_w ? max_w : ctx->max_w, h = max_h > ctx->max_h ? max_h : ctx->max_h;
    const int b = max_batch > ctx->max_batch ? max_batch : ctx->max_batch;
    if (w == ctx->max_w && h == ctx->max_h && b == ctx->max_batch) return VA_OK;
    size_t lab_pitch = 32;
    while (lab_pitch < (size_t)w) lab_pitch <<= 1;
    if (lab_pitch * (size_t)h >= ((size_t)1 << 31)) VA_FAIL(ctx, VA_ERR_CAPACITY, "va_reserve: %dx%d frames exceed the forest's index range", w, h);
    VA_CUDA(ctx, cudaSetDevice(ctx->device));
    VA_CUDA(ctx, cudaDeviceSynchronize());              // nothing in flight uses the old scratch any more
    int32_t *parent = nullptr, *rowcnt = nullptr;
    if (cudaMalloc((void **)&parent, lab_pitch * (size_t)h * (size_t)b * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc((void **)&rowcnt, (size_t)2 * h * b * sizeof(int32_t)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(parent);
        VA_FAIL(ctx, VA_ERR_NOMEM, "va_reserve: cannot allocate the labelling scratch for %dx%dx%d", w, h, b);
    }
    cudaFree(ctx->lab_parent);
    cudaFree(ctx->lab_rowcnt);
    ctx->lab_parent = parent;
    ctx->lab_rowcnt = rowcnt;
    ctx->lab_pitch = lab_pitch;
    // everything sized by the capacity and allocated on first use starts over
    cudaFree(ctx->lab_parent1); ctx->lab_parent1 = nullptr;
    cudaFree(ctx->lab_rowcnt1); ctx->lab_rowcnt1 = nullptr;
    cudaFree(ctx->ch_mono); ctx->ch_mono = nullptr;
    cudaFree(ctx->ch_blur); ctx->ch_blur = nullptr;
    cudaFree(ctx->ch_mask); ctx->ch_mask = nullptr;
    cudaFree(ctx->ch_morph); ctx->ch_morph = nullptr;
    cudaFree(ctx->exp_rowoff); ctx->exp_rowoff = nullptr;
    ctx->max_w = w; ctx->max_h = h; ctx->max_batch = b;
    return VA_OK;
}

extern "C" int va_destroy(va_ctx *ctx) {
    if (!ctx) return VA_OK;
    cudaFree(ctx->lab_parent);
    cudaFree(ctx->lab_rowcnt);
    cudaFree(ctx->lab_parent1);
    cudaFree(ctx->lab_rowcnt1);
    cudaFree(ctx->ch_mono);
    cudaFree(ctx->ch_blur);
    cudaFree(ctx->ch_mask);
    cudaFree(ctx->ch_morph);
    cudaFree(ctx->exp_rowoff);
    cudaFree(ctx->rs_tab);
#ifndef VA_EMU
    for (int i = 0; i < 5; i++)
        if (ctx->lab_event[i]) cudaEventDestroy((cudaEvent_t)ctx->lab_event[i]);
#endif
    free(ctx);
    return VA_OK;
}

// ---------------------------------------------------------------------------------
// the chain of BASELINE.json configs 1-3 on one device-resident batch
// ---------------------------------------------------------------------------------
static int chain_scratch(va_ctx *ctx) {
    if (ctx->ch_blur) return VA_OK;
    ctx->ch_pitch = ((size_t)ctx->max_w + 127) / 128 * 128;
    ctx->ch_pitch_w = (((size_t)ctx->max_w + 31) / 32 + 3) / 4 * 4;
    const size_t n8 = ctx->ch_pitch * ctx->max_h * ctx->max_batch;
    const size_t nw = ctx->ch_pitch_w * ctx->max_h * ctx->max_batch * 4;
    if (cudaMalloc((void **)&ctx->ch_mono, n8) != cudaSuccess || cudaMalloc((void **)&ctx->ch_blur, n8) != cudaSuccess ||
        cudaMalloc((void **)&ctx->ch_mask, nw) != cudaSuccess || cudaMalloc((void **)&ctx->ch_morph, nw) != cudaSuccess) {
        cudaGetLastError();
        VA_FAIL(ctx, VA_ERR_NOMEM, "va_chain_run: cannot allocate chain scratch");
    }
    return VA_OK;
}

extern "C" int va_chain_run(va_ctx *ctx, va_stream stream, const va_chain_desc *d, const va_chain_io *io) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, d && io && io->rgb && io->bg, "va_chain_run: null pointer");
    const int w = d->w, h = d->h, batch = d->batch;
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "va_chain_run: bad size");
    if (w > ctx->max_w || h > ctx->max_h || batch > ctx->max_batch)
        VA_FAIL(ctx, VA_ERR_CAPACITY, "va_chain_run: %dx%dx%d exceeds the ctx capacity %dx%dx%d", w, h, batch,
                ctx->max_w, ctx->max_h, ctx->max_batch);
    int rc = chain_scratch(ctx);
    if (rc != VA_OK) return rc;
    VA_REQUIRE(ctx, va_scratch_acquire(ctx, stream, 2) == 0, "va_chain_run: cannot order the chain scratch");
    const size_t sp = ctx->ch_pitch, sf = ctx->ch_pitch * (size_t)ctx->max_h;
    const size_t wp = ctx->ch_pitch_w, wf = ctx->ch_pitch_w * (size_t)ctx->max_h;

    // mono (+ blur)
    uint8_t *mono = io->mono ? io->mono : ctx->ch_mono;
    const size_t mono_p = io->mono ? io->mono_pitch : sp, mono_f = io->mono ? io->mono_fstride : sf;
    uint8_t *blur = io->blur ? io->blur : ctx->ch_blur;
    const size_t blur_p = io->blur ? io->blur_pitch : sp, blur_f = io->blur ? io->blur_fstride : sf;
    const uint8_t *gray = nullptr;
    size_t gray_p = 0, gray_f = 0;
    if (d->sigma > 0) {
        bool fused = false;
        if (d->fuse_luma_blur && !io->mono) {
            rc = va_luma_gauss_u8(ctx, stream, io->rgb, io->rgb_pitch, io->rgb_fstride, blur, blur_p, blur_f,
                                  w, h, batch, d->mono_mode, d->sigma);
            if (rc == VA_OK) fused = true;
            else if (rc != VA_ERR_UNSUPPORTED) return rc;
        }
        if (!fused) {
            rc = va_luma_u8(ctx, stream, io->rgb, io->rgb_pitch, io->rgb_fstride, mono, mono_p, mono_f, w, h, batch, d->mono_mode);
            if (rc != VA_OK) return rc;
            rc = va_gauss_u8(ctx, stream, mono, mono_p, mono_f, blur, blur_p, blur_f, w, h, 1, batch, d->sigma);
            if (rc != VA_OK) return rc;
        }
        gray = blur; gray_p = blur_p; gray_f = blur_f;
    } else {
        rc = va_luma_u8(ctx, stream, io->rgb, io->rgb_pitch, io->rgb_fstride, mono, mono_p, mono_f, w, h, batch, d->mono_mode);
        if (rc != VA_OK) return rc;
        gray = mono; gray_p = mono_p; gray_f = mono_f;
    }

    // background model / difference / threshold
    uint32_t *mask = io->mask ? io->mask : ctx->ch_mask;
    const size_t mask_p = io->mask ? io->mask_pitch_w : wp, mask_f = io->mask ? io->mask_fstride_w : wf;
    rc = va_ema_diff_thresh(ctx, stream, gray, gray_p, gray_f, io->bg, io->bg_pitch_e, mask, mask_p, mask_f,
                            w, h, batch, d->alpha, d->thr, d->first_frame_inits);
    if (rc != VA_OK) return rc;

    // morphology
    const uint32_t *seg = mask;
    size_t seg_p = mask_p, seg_f = mask_f;
    if (d->morph_op >= 0) {
        uint32_t *mo = io->morph ? io->morph : ctx->ch_morph;
        const size_t mo_p = io->morph ? io->morph_pitch_w : wp, mo_f = io->morph ? io->morph_fstride_w : wf;
        rc = va_morph_bits(ctx, stream, mask, mask_p, mask_f, mo, mo_p, mo_f, w, h, batch,
                           d->morph_op, d->morph_shape, d->morph_kx, d->morph_ky);
        if (rc != VA_OK) return rc;
        seg = mo; seg_p = mo_p; seg_f = mo_f;
    }

    // labelling
    if (d->connectivity) {
        VA_REQUIRE(ctx, io->labels, "va_chain_run: labelling requested without a labels buffer");
        rc = va_label_bits(ctx, stream, seg, seg_p, seg_f, io->labels, io->labels_pitch_e, io->labels_fstride_e,
                           io->counts, w, h, batch, d->connectivity);
        if (rc != VA_OK) return rc;
    }
    VA_REQUIRE(ctx, va_scratch_release(ctx, stream, 2) == 0, "va_chain_run: cannot order the chain scratch");
    return VA_OK;
}
