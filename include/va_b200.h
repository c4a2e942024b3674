/*
 * va_b200.h -- C ABI of libva_b200.so: the B200 (sm_100a) filter -> segment path
 * of david-zwicker/video-analysis.
 *
 * The reference is pure Python and has no FFI; each entry point below replaces
 * the body of one reference `_process_frame` / analysis helper (cited per
 * function, paths relative to the reference root) for a BATCH of frames.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - every function returns VA_OK (0) or a negative va_status; nothing throws.
 *    va_last_error(ctx) holds a human readable message for the last failure.
 *  - all image pointers are DEVICE pointers owned by the caller.  Kernels are
 *    enqueued on `stream` (a cudaStream_t passed as void*) and never
 *    synchronise; the call returns as soon as the work is queued.
 *  - u8 images: `pitch` = bytes between rows, `fstride` = bytes between frames
 *    of the batch.  Colour frames are interleaved (H, W, 3) as in the reference
 *    (video/io/backend_ffmpeg.py:318-319).  Nothing needs to be aligned; 16-byte
 *    aligned pointers / pitches take the vector paths.
 *  - bit masks: 32-bit words, bit i of word j of a row = pixel x = 32*j + i
 *    (LSB = lowest x).  `pitch_w` / `fstride_w` are counted in words.  Bits at
 *    x >= w in the last word of a row are written as 0 and ignored on input.
 *  - labels: int32, `pitch_e` / `fstride_e` counted in elements.
 *  - one ctx per GPU per host thread; a ctx owns only scratch memory.
 */
#ifndef VA_B200_H
#define VA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct va_ctx va_ctx;
typedef void *va_stream; /* cudaStream_t */

typedef enum va_status {
    VA_OK = 0,
    VA_ERR_INVALID = -1,     /* bad argument (ValueError on the Python side) */
    VA_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    VA_ERR_NOMEM = -3,       /* scratch allocation failed */
    VA_ERR_CAPACITY = -4,    /* w, h or batch exceed what the ctx was created for */
    VA_ERR_UNSUPPORTED = -5  /* parameter combination not implemented on the device */
} va_status;

/* morphology: operations and structuring-element shapes (cv2.MORPH_* order) */
enum { VA_MORPH_ERODE = 0, VA_MORPH_DILATE = 1, VA_MORPH_OPEN = 2, VA_MORPH_CLOSE = 3 };
enum { VA_SE_RECT = 0, VA_SE_CROSS = 1, VA_SE_ELLIPSE = 2 };
/* luma modes: FilterMonochrome(mode=...) video/filters.py:351-368 */
enum { VA_MONO_MEAN = -1, VA_MONO_CH0 = 0, VA_MONO_CH1 = 1, VA_MONO_CH2 = 2 };

int va_version(void);
const char *va_status_string(int status);

/* scratch is sized for frames up to max_w x max_h and batches up to max_batch */
int va_create(va_ctx **out, int device, int max_w, int max_h, int max_batch);
/* grows the capacity of a live ctx (never shrinks it) without invalidating the handle: the device is synchronised, the
 * scratch reallocated.  Other users of the same ctx keep working; only a va_label_forest / va_label_write pair must
 * not straddle the call. */
int va_reserve(va_ctx *ctx, int max_w, int max_h, int max_batch);
int va_destroy(va_ctx *ctx);
const char *va_last_error(const va_ctx *ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
long long va_launch_count(const va_ctx *ctx);

/* K1 FilterMonochrome._process_frame, video/filters.py:359-374:
 *    mode VA_MONO_MEAN: out = (c0 + c1 + c2) / 3  (== np.mean(axis=2).astype(u8), :366)
 *    mode 0/1/2:        out = in[:, :, mode]      (:368; also FilterCrop's color_channel, :245)
 * A crop (FilterCrop._process_frame, video/filters.py:238-248) is expressed by
 * the caller offsetting `in` to the first pixel of the rectangle. */
int va_luma_u8(va_ctx *ctx, va_stream stream,
               const uint8_t *in, size_t in_pitch, size_t in_fstride,
               uint8_t *out, size_t out_pitch, size_t out_fstride,
               int w, int h, int batch, int mode);

/* strided 2-D copy of `row_bytes` x h x batch (FilterCrop on frames that stay
 * colour / monochrome, video/filters.py:241) */
int va_copy2d_u8(va_ctx *ctx, va_stream stream,
                 const uint8_t *in, size_t in_pitch, size_t in_fstride,
                 uint8_t *out, size_t out_pitch, size_t out_fstride,
                 int row_bytes, int h, int batch);

/* K2 FilterBlur._process_frame, video/filters.py:388-392:
 *    cv2.GaussianBlur(u8, (0, 0), sigma), bit-exact: 8-bit fixed-point kernel,
 *    out = (sum_ky sum_kx K[ky] K[kx] src + 32768) >> 16, BORDER_REFLECT_101.
 * channels = 1 or 3 (interleaved, filtered independently). in != out. */
int va_gauss_u8(va_ctx *ctx, va_stream stream,
                const uint8_t *in, size_t in_pitch, size_t in_fstride,
                uint8_t *out, size_t out_pitch, size_t out_fstride,
                int w, int h, int channels, int batch, double sigma);
/* the integer taps va_gauss_u8 uses for `sigma` (host side; taps[] needs room
 * for 6*sigma+3 entries); returns ksize or a negative status */
int va_gauss_taps(double sigma, int *taps, int capacity);

/* K1+K2 fused: FilterBlur(FilterMonochrome(video)) without materialising the
 * monochrome frame (same results as va_luma_u8 followed by va_gauss_u8) */
int va_luma_gauss_u8(va_ctx *ctx, va_stream stream,
                     const uint8_t *in, size_t in_pitch, size_t in_fstride,
                     uint8_t *out, size_t out_pitch, size_t out_fstride,
                     int w, int h, int batch, int mode, double sigma);

/* K2b FilterResize._process_frame, video/filters.py:308-315, for the exact
 * integer case cv2.resize(INTER_AREA) by 1/2: out = (a + b + c + d + 2) >> 2.
 * (w, h) is the INPUT size; output is (w/2, h/2); w and h must be even. */
int va_resize_half_u8(va_ctx *ctx, va_stream stream,
                      const uint8_t *in, size_t in_pitch, size_t in_fstride,
                      uint8_t *out, size_t out_pitch, size_t out_fstride,
                      int w, int h, int channels, int batch);

/* FilterResize._process_frame, video/filters.py:308-315, cv2.resize(INTER_AREA) when both scale factors
 * are integers (OpenCV's ResizeAreaFast): out = round_half_even(float(sum over kx x ky) * (1.f / (kx ky))),
 * and (sum + 2) >> 2 for 2 x 2.  (w, h) is the INPUT size, a multiple of (kx, ky); output (w / kx, h / ky). */
int va_resize_area_u8(va_ctx *ctx, va_stream stream,
                      const uint8_t *in, size_t in_pitch, size_t in_fstride,
                      uint8_t *out, size_t out_pitch, size_t out_fstride,
                      int w, int h, int channels, int batch, int kx, int ky);

/* the same call site with cv2.INTER_NEAREST, any output size (dw, dh):
 * out(x, y) = in(min(floor(x * (1 / (dw / w))), w - 1), min(floor(y * (1 / (dh / h))), h - 1)), in doubles */
int va_resize_nearest_u8(va_ctx *ctx, va_stream stream,
                         const uint8_t *in, size_t in_pitch, size_t in_fstride,
                         uint8_t *out, size_t out_pitch, size_t out_fstride,
                         int w, int h, int dw, int dh, int channels, int batch);

/* the same call site with cv2.INTER_AREA shrinking by any (non-integer) factors to (dw, dh): OpenCV's
 * computeResizeAreaTab / resizeArea_<uchar, float> -- cell tables in doubles, weights and sums in float32
 * with every operation rounded on its own, round half to even.  Integer factors dispatch to the calls above.
 * Where either direction enlarges, OpenCV's INTER_AREA is its fixed-point linear interpolation with the area
 * coefficient rule (s = floor(d scale), f = (d + 1) - (s + 1) / scale, f = f <= 0 ? 0 : f - floor(f)): same here. */
int va_resize_area_any_u8(va_ctx *ctx, va_stream stream,
                          const uint8_t *in, size_t in_pitch, size_t in_fstride,
                          uint8_t *out, size_t out_pitch, size_t out_fstride,
                          int w, int h, int dw, int dh, int channels, int batch);

/* the same call site with cv2.INTER_LINEAR, any output size (OpenCV's 8-bit fixed-point path: 11-bit
 * coefficients, out = (((b0 (H0 >> 4)) >> 16) + ((b1 (H1 >> 4)) >> 16) + 2) >> 2 with H = S[sx] a0 + S[sx+1] a1;
 * shrinking by exactly 2 x 2 is INTER_AREA, as in cv2) */
int va_resize_linear_u8(va_ctx *ctx, va_stream stream,
                        const uint8_t *in, size_t in_pitch, size_t in_fstride,
                        uint8_t *out, size_t out_pitch, size_t out_fstride,
                        int w, int h, int dw, int dh, int channels, int batch);

/* the same call site with cv2.INTER_CUBIC ('auto' when enlarging, video/filters.py:282-284), any output size:
 * OpenCV's own 8-bit path (a = -0.75, 11-bit coefficients, float32 column pass on the 8-lane vector body of a
 * row, (sum + 2^21) >> 22 on its tail).  Bit-exact against cv2 with cv2.ipp.setUseIPP(False); the cv2 wheel's
 * default routes this interpolation through Intel IPP, whose result differs by at most 1 LSB. */
int va_resize_cubic_u8(va_ctx *ctx, va_stream stream,
                       const uint8_t *in, size_t in_pitch, size_t in_fstride,
                       uint8_t *out, size_t out_pitch, size_t out_fstride,
                       int w, int h, int dw, int dh, int channels, int batch);

/* the same call site with cv2.INTER_LANCZOS4, any output size: OpenCV's 8-bit fixed-point path (8 taps per
 * direction, 11-bit coefficients from interpolateLanczos4, out = saturate((sum + 2^21) >> 22)).  The coefficient
 * tables are computed on the host by this call and copied in a stream-ordered allocation.  Bit-exact against cv2. */
int va_resize_lanczos4_u8(va_ctx *ctx, va_stream stream,
                          const uint8_t *in, size_t in_pitch, size_t in_fstride,
                          uint8_t *out, size_t out_pitch, size_t out_fstride,
                          int w, int h, int dw, int dh, int channels, int batch);

/* K3 running-average background + |difference| > thr -> packed mask bits.
 * Not in the reference; fold shape follows
 * video/analysis/video.py:14-35, signed difference video/filters.py:564-568:
 *    d = float(x_t) - bg;  mask_t = |d| > thr;  bg = bg + alpha * d   (float32,
 *    multiply and add rounded separately, frames of the batch in order)
 * `bg` (float32, pitch_e elements per row) is the state before frame 0 of the
 * batch and is updated in place.  If first_frame_inits != 0, frame 0 of the
 * batch initialises bg = float(x_0) and gets an all-zero mask. */
int va_ema_diff_thresh(va_ctx *ctx, va_stream stream,
                       const uint8_t *in, size_t in_pitch, size_t in_fstride,
                       float *bg, size_t bg_pitch_e,
                       uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                       int w, int h, int batch,
                       float alpha, float thr, int first_frame_inits);

/* mask = in > thr (strict, == cv2.threshold(..., THRESH_BINARY)) -> packed bits */
int va_threshold_bits(va_ctx *ctx, va_stream stream,
                      const uint8_t *in, size_t in_pitch, size_t in_fstride,
                      uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                      int w, int h, int batch, int thr);

/* packed bits <-> u8 masks ({0,255} out; nonzero in) */
int va_unpack_bits_u8(va_ctx *ctx, va_stream stream,
                      const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                      uint8_t *out, size_t out_pitch, size_t out_fstride,
                      int w, int h, int batch);
int va_pack_bits_u8(va_ctx *ctx, va_stream stream,
                    const uint8_t *in, size_t in_pitch, size_t in_fstride,
                    uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                    int w, int h, int batch);

/* K4 binary morphology on packed bits; replaces cv2.erode / cv2.dilate
 * (video/analysis/image.py:248-256) and cv2.morphologyEx(OPEN / CLOSE) with
 * cv2.getStructuringElement(shape, (kx, ky)) and OpenCV's default border
 * (pixels outside the image never win the min / max).  in != out. */
int va_morph_bits(va_ctx *ctx, va_stream stream,
                  const uint32_t *in, size_t in_pitch_w, size_t in_fstride_w,
                  uint32_t *out, size_t out_pitch_w, size_t out_fstride_w,
                  int w, int h, int batch, int op, int shape, int kx, int ky);

/* va_label_bits in two halves that share scratch set `slot` (0 or 1) of the ctx: the union-find forest of a batch
 * (counts[] complete after it) and the write of its label image.  Each is enqueued on the stream it is given; the
 * caller orders write(k) after forest(k) and, per slot, the next forest after the write that used the slot.
 * Lets the store-bound write of one batch overlap the latency-bound forest kernels of the next. */
int va_label_forest(va_ctx *ctx, va_stream stream,
                    const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                    int32_t *counts, int w, int h, int batch, int connectivity, int slot);
int va_label_write(va_ctx *ctx, va_stream stream,
                   const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                   int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                   int w, int h, int batch, int slot);

/* BASELINE.json configs[3] -- concurrent streams batched on the leading axis of one launch: FilterCrop
 * (video/filters.py:158-248) with one rectangle SIZE (w, h) but a position per stream, then FilterMonochrome
 * (video/filters.py:359-374).  xy: DEVICE table {left_0, top_0, left_1, top_1, ...}; the caller guarantees
 * left_s + w <= in_w and top_s + h <= in_h (the host layer applies the reference's _check_coordinate rules). */
int va_luma_crop_multi_u8(va_ctx *ctx, va_stream stream,
                          const uint8_t *in, size_t in_pitch, size_t in_fstride, int in_w, int in_h,
                          uint8_t *out, size_t out_pitch, size_t out_fstride,
                          int w, int h, int batch, int mode, const int32_t *xy);

/* the front of BASELINE.json configs[3] fused into one pass: crop position per stream + monochrome + static mask
 * (u8, nonzero = keep; smask = NULL: none; smask_fstride = 0: one mask for all streams) + threshold (value > thr)
 * -> packed mask bits.  Needs w % 32 == 0, word-aligned frames and 16-byte aligned mask rows, otherwise
 * VA_ERR_UNSUPPORTED (the three separate calls cover every case). */
int va_streams_threshold_bits(va_ctx *ctx, va_stream stream,
                              const uint8_t *in, size_t in_pitch, size_t in_fstride, int in_w, int in_h,
                              const uint8_t *smask, size_t smask_pitch, size_t smask_fstride,
                              uint32_t *bits, size_t bits_pitch_w, size_t bits_fstride_w,
                              int w, int h, int batch, int mode, int thr, const int32_t *xy);

/* va_label_write with int16 labels -- ndimage.label(mask, output=np.int16) at the same call site: half the bytes to
 * write and to copy to the host.  The caller checks counts[] <= 32767 first (scipy raises "insufficient bit-depth in
 * requested output type"; larger labels would wrap). */
int va_label_write_i16(va_ctx *ctx, va_stream stream,
                       const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                       int16_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                       int w, int h, int batch, int slot);

/* VideoComposer.highlight_mask, video/io/composer.py:131-154: where the mask is set,
 * frame[mask, channel] = uint8(strength + (255 - strength) / 255 * frame[mask, channel]).  lut256 (HOST pointer,
 * copied into the launch) is that expression on 0..255; channel -1 = all channels, 0..2 = one channel of an
 * interleaved colour frame.  The mask is the packed-bit image K3 / K4 produce (LSB = lowest x). */
int va_highlight_mask_u8(va_ctx *ctx, va_stream stream,
                         const uint8_t *in, size_t in_pitch, size_t in_fstride,
                         const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                         uint8_t *out, size_t out_pitch, size_t out_fstride,
                         int w, int h, int channels, int batch, int channel, const uint8_t *lut256);

/* K5 connected-component labelling; replaces
 * ndimage.measurements.label(mask) at video/analysis/regions.py:162.
 * connectivity 4 (scipy default) or 8.  labels: 0 background, 1..n numbered in
 * raster order of each component's first pixel.  counts[b] = n of frame b. */
int va_label_bits(va_ctx *ctx, va_stream stream,
                  const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                  int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                  int32_t *counts, int w, int h, int batch, int connectivity);

/* per-label pixel counts (video/analysis/regions.py:165-166) and the largest
 * region (regions.py:169): areas[b * max_labels + (l-1)], largest[b] = label */
int va_region_areas(va_ctx *ctx, va_stream stream,
                    const int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                    int32_t *areas, int max_labels, int32_t *largest,
                    int w, int h, int batch);

/* per-region statistics straight from a packed mask -- the label image is never written (nor
 * copied to the host).  Regions are numbered like va_label_bits numbers them.  For region l of
 * frame b, stats[(b * max_regions + l - 1) * VA_REGION_FIELDS + k] holds, as exact integers,
 *   k = 0..5  raw moments m00 m10 m01 m20 m11 m02 of cv2.moments(region.astype(uint8)), the input
 *             of regionprops (video/analysis/image.py:347-352; m00 = area of regions.py:165-166)
 *   k = 6..9  xmin ymin xmax ymax (find_bounding_box, video/analysis/regions.py:113-149)
 * counts[b] = number of regions of frame b (rows of regions beyond max_regions are not written);
 * largest[b] = label of the largest region (regions.py:169), may be NULL. */
#define VA_REGION_FIELDS 10
int va_region_stats(va_ctx *ctx, va_stream stream,
                    const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                    int64_t *stats, int max_regions, int32_t *counts, int32_t *largest,
                    int w, int h, int batch, int connectivity);

/* K6 apply-mask: out = mask != 0 ? in : 0 (not in the reference; idioms
 * video/io/composer.py:154,186,208).  mask is u8 (H, W); mask_fstride = 0
 * broadcasts one static mask over the batch. */
int va_apply_mask_u8(va_ctx *ctx, va_stream stream,
                     const uint8_t *in, size_t in_pitch, size_t in_fstride,
                     const uint8_t *mask, size_t mask_pitch, size_t mask_fstride,
                     uint8_t *out, size_t out_pitch, size_t out_fstride,
                     int w, int h, int channels, int batch);

/* ---- remaining filter bodies and temporal statistics ("next" rows of SURVEY.md 8f) ---- */

/* FilterNormalize._process_frame, video/filters.py:101-135, for uint8 -> uint8 frames: the
 * clip / scale / cast is applied as a 256-entry table (HOST pointer) computed by the caller with
 * the reference's expression. */
int va_lut_u8(va_ctx *ctx, va_stream stream,
              const uint8_t *in, size_t in_pitch, size_t in_fstride,
              uint8_t *out, size_t out_pitch, size_t out_fstride,
              int row_bytes, int h, int batch, const uint8_t *lut256);

/* FilterTimeDifference._compare_frames, video/filters.py:564-568:
 *    out[t] = int16(in[t + 1]) - in[t], t in [0, batch); `in` holds batch + 1 frames */
int va_time_diff_i16(va_ctx *ctx, va_stream stream,
                     const uint8_t *in, size_t in_pitch, size_t in_fstride,
                     int16_t *out, size_t out_pitch_e, size_t out_fstride_e,
                     int row_elems, int h, int batch);

/* FilterRotate._process_frame, video/filters.py:338-344: np.rot90(frame, k), k in 0..3
 * (counter-clockwise).  (w, h) is the input size; the output is (h, w) for odd k. */
int va_rot90_u8(va_ctx *ctx, va_stream stream,
                const uint8_t *in, size_t in_pitch, size_t in_fstride,
                uint8_t *out, size_t out_pitch, size_t out_fstride,
                int w, int h, int channels, int batch, int k);

/* measure_mean / measure_mean_std, video/analysis/video.py:26-55 (float64 state, reference
 * operation order).  m2 == NULL: cumulative mean only.  n0 = frames already folded in. */
int va_mean_update_f64(va_ctx *ctx, va_stream stream,
                       const uint8_t *in, size_t in_pitch, size_t in_fstride,
                       double *mean, double *m2, size_t pitch_e,
                       int row_elems, int h, int batch, long long n0);

/* frame-sharded EMA.  The recurrence is affine,
 *    bg_t = a * bg_{t-1} + alpha * x_t,   a = 1 - alpha,
 * so a rank can fold its frames from a zero state into S and the true state is
 * recovered from the ranks before it.
 * va_ema_partial:  S = a^batch * S + sum_i alpha * a^(batch-1-i) * x_i
 *                  (S is overwritten when accumulate == 0)
 * va_ema_fold:     carry = scale * carry + S          (element-wise, float32) */
int va_ema_partial(va_ctx *ctx, va_stream stream,
                   const uint8_t *in, size_t in_pitch, size_t in_fstride,
                   float *S, size_t s_pitch_e, int w, int h, int batch,
                   float alpha, int accumulate);
int va_ema_fold(va_ctx *ctx, va_stream stream, float *carry, const float *S,
                size_t pitch_e, int w, int h, float scale);

/* seeded synthetic RGB video generated on the device; replaces
 * the unseeded VideoGaussianNoise (video/io/computed.py:15-41).  `blobs` is a
 * HOST array of n_blobs x 5 int32 (x0, y0, vx16, vy16, r). Frames t0..t0+batch-1. */
int va_synth_rgb(va_ctx *ctx, va_stream stream,
                 uint8_t *out, size_t out_pitch, size_t out_fstride,
                 int w, int h, int t0, int batch, uint32_t seed,
                 const int32_t *blobs, int n_blobs);

/* the whole chain of BASELINE.json configs 1-3 on one device-resident batch:
 * mono -> blur -> EMA/diff/threshold -> morphology -> label.  Intermediates live
 * in ctx scratch unless an output pointer is given (NULL = not wanted). */
typedef struct va_chain_desc {
    int w, h, batch;
    int mono_mode;           /* VA_MONO_* */
    double sigma;            /* <= 0: no blur */
    float alpha, thr;
    int first_frame_inits;   /* this batch starts the video */
    int morph_op;            /* VA_MORPH_* or -1 for none */
    int morph_shape, morph_kx, morph_ky;
    int connectivity;        /* 4 or 8; 0 = no labelling */
    int fuse_luma_blur;      /* 1: use va_luma_gauss_u8 when `mono` is not wanted */
} va_chain_desc;

typedef struct va_chain_io {
    const uint8_t *rgb; size_t rgb_pitch, rgb_fstride;         /* in  (H, W, 3) u8 */
    float *bg; size_t bg_pitch_e;                              /* in/out state */
    uint8_t *mono; size_t mono_pitch, mono_fstride;            /* out, optional */
    uint8_t *blur; size_t blur_pitch, blur_fstride;            /* out, optional */
    uint32_t *mask; size_t mask_pitch_w, mask_fstride_w;       /* out, optional (after threshold) */
    uint32_t *morph; size_t morph_pitch_w, morph_fstride_w;    /* out, optional (after morphology) */
    int32_t *labels; size_t labels_pitch_e, labels_fstride_e;  /* out, optional */
    int32_t *counts;                                           /* out, optional */
} va_chain_io;

int va_chain_run(va_ctx *ctx, va_stream stream, const va_chain_desc *desc, const va_chain_io *io);

/* Sparse egress of a label image (the label image of analysis/regions.py:162 is mostly background): the non-empty
 * chunks -- VA_CHUNK_E = 64 consecutive labels of a row, numbered y * ceil(w / 64) + c -- of every frame in raster
 * order.  ids [batch][cap], data [batch][cap][64] and n_chunks [batch] are written by the device; they may be
 * PAGE-LOCKED HOST memory (the kernels then store over PCIe and no copy has to be sized or issued), n_chunks_dev
 * (optional) is a device copy of the counts.  A frame with more than cap non-empty chunks reports its true count and
 * exports the first cap.  mask is the packed image the labels were made from (va_label_bits).  With ids = data = NULL
 * only the counts are produced (how sparse is this batch?).
 * runs [batch][cap][4] / n_runs [batch] (optional, both or neither; 16-byte aligned): chunks whose foreground pixels
 * form a single horizontal run -- they carry one label -- are then exported as (id, label, low word, high word of the
 * 64-bit pixel mask) records of 16 bytes instead of 260, and ids / data / n_chunks hold the remaining chunks only.
 * The count-only call accepts n_runs alone and then counts the two kinds separately. */
int va_label_export_chunks(va_ctx *ctx, va_stream stream,
                           const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                           const int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                           int w, int h, int batch,
                           int32_t *ids, int32_t *data, int32_t *n_chunks, int32_t *n_chunks_dev,
                           int32_t *runs, int32_t *n_runs, int cap);

/* HOST function (no CUDA call, no ctx): rebuilds dense (batch, h, pitch_e) int32 label images from exported chunks
 * (and run chunks, when runs / n_runs are given).  dirty_ids [batch][cap] / n_dirty [batch] name the chunks of `dense`
 * that are non-zero from its previous use: they are cleared first and replaced by the chunk list written now, so a
 * ring of result buffers never has to be zeroed as a whole (n_chunks + n_runs <= cap).  `threads` host threads share
 * the frames. */
int va_host_densify_chunks(int32_t *dense, size_t pitch_e, size_t fstride_e, int w, int h, int batch,
                           const int32_t *ids, const int32_t *data, const int32_t *n_chunks,
                           const int32_t *runs, const int32_t *n_runs, int cap,
                           int32_t *dirty_ids, int32_t *n_dirty, int threads);

#ifdef __cplusplus
}
#endif
#endif /* VA_B200_H */
