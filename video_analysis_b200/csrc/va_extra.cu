// va_extra.cu -- the remaining per-frame filter bodies of video/filters.py and the temporal
// statistics of video/analysis/video.py ("next" rows of SURVEY.md 8f).  Simple HBM-streaming
// kernels; none of them is on the benchmarked chain.
#include "va_device.cuh"

#include <cfloat>
#include <cmath>
#include <vector>

// =================================================================================
// FilterNormalize._process_frame (video/filters.py:101-135) for uint8 -> uint8: the clip /
// scale / cast is a 256-entry table that the host computes with the reference's expression
// =================================================================================
struct Lut256 { unsigned char v[256]; };

__global__ void __launch_bounds__(256)
lut_u8_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
              uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
              int row_bytes, int h, int batch, int vec, const __grid_constant__ Lut256 lut) {
    __shared__ unsigned char s[256];
    s[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const unsigned rows = (unsigned)(h * batch);
    if (vec) {
        const unsigned chunks = (unsigned)row_bytes >> 4;
        const unsigned total = rows * chunks;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const unsigned row = i / chunks, c = i - row * chunks;
            const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
            uint4 q = va_ld_stream16(in + (size_t)b * in_fstride + (size_t)y * in_pitch + 16 * c);
            unsigned *w = reinterpret_cast<unsigned *>(&q);
#pragma unroll
            for (int k = 0; k < 4; k++)
                w[k] = s[w[k] & 0xff] | (s[(w[k] >> 8) & 0xff] << 8) | (s[(w[k] >> 16) & 0xff] << 16) | (s[w[k] >> 24] << 24);
            va_st_stream16(out + (size_t)b * out_fstride + (size_t)y * out_pitch + 16 * c, q);
        }
    } else {
        const unsigned total = rows * (unsigned)row_bytes;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const unsigned row = i / (unsigned)row_bytes, c = i - row * (unsigned)row_bytes;
            const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
            out[(size_t)b * out_fstride + (size_t)y * out_pitch + c] = s[in[(size_t)b * in_fstride + (size_t)y * in_pitch + c]];
        }
    }
}

extern "C" int va_lut_u8(va_ctx *ctx, va_stream stream,
                         const uint8_t *in, size_t in_pitch, size_t in_fstride,
                         uint8_t *out, size_t out_pitch, size_t out_fstride,
                         int row_bytes, int h, int batch, const uint8_t *lut256) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && lut256, "va_lut_u8: null pointer");
    VA_REQUIRE(ctx, row_bytes > 0 && h > 0 && batch > 0, "va_lut_u8: bad size");
    Lut256 lut;
    memcpy(lut.v, lut256, 256);
    const int vec = row_bytes % 16 == 0 && va_aligned(in, 16) && va_aligned(out, 16) && in_pitch % 16 == 0 &&
                    out_pitch % 16 == 0 && in_fstride % 16 == 0 && out_fstride % 16 == 0;
    const long long items = (long long)h * batch * (vec ? row_bytes / 16 : row_bytes);
    VA_REQUIRE(ctx, items < (1ll << 31), "va_lut_u8: batch too large for one launch");
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = lut_u8_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, row_bytes, h, batch, vec, lut);
    return VA_OK;
}

// =================================================================================
// FilterTimeDifference._compare_frames (video/filters.py:564-568):
//   out[t] = int16(frame[t + 1]) - frame[t]      for t in [0, batch)   (in holds batch + 1 frames)
// =================================================================================
__global__ void __launch_bounds__(256)
time_diff_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                 int16_t *__restrict__ out, size_t out_pitch_e, size_t out_fstride_e,
                 int row_elems, int h, int batch) {
    const unsigned total = (unsigned)row_elems * (unsigned)h * (unsigned)batch;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned row = i / (unsigned)row_elems, x = i - row * (unsigned)row_elems;
        const int t = (int)(row / (unsigned)h), y = (int)(row - (unsigned)t * (unsigned)h);
        const uint8_t *p = in + (size_t)t * in_fstride + (size_t)y * in_pitch + x;
        out[(size_t)t * out_fstride_e + (size_t)y * out_pitch_e + x] = (int16_t)((int)p[in_fstride] - (int)p[0]);
    }
}

extern "C" int va_time_diff_i16(va_ctx *ctx, va_stream stream,
                                const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                int16_t *out, size_t out_pitch_e, size_t out_fstride_e,
                                int row_elems, int h, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out, "va_time_diff_i16: null pointer");
    VA_REQUIRE(ctx, row_elems > 0 && h > 0 && batch > 0, "va_time_diff_i16: bad size");
    const long long items = (long long)row_elems * h * batch;
    VA_REQUIRE(ctx, items < (1ll << 31), "va_time_diff_i16: batch too large for one launch");
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = time_diff_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch_e, out_fstride_e, row_elems, h, batch);
    return VA_OK;
}

// =================================================================================
// FilterRotate._process_frame (video/filters.py:338-344): np.rot90(frame, k), counter-clockwise
//   (w, h) is the INPUT size; output is (h, w) for odd k
// =================================================================================
__global__ void __launch_bounds__(256)
rot90_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
             uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
             int w, int h, int channels, int batch, int k) {
    const int ow = (k & 1) ? h : w, oh = (k & 1) ? w : h;
    const unsigned total = (unsigned)ow * (unsigned)oh * (unsigned)batch;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned row = i / (unsigned)ow;
        const int j = (int)(i - row * (unsigned)ow);                       // output column
        const int b = (int)(row / (unsigned)oh), r = (int)(row - (unsigned)b * (unsigned)oh);   // output row
        int sy, sx;                                                        // source pixel
        if (k == 0) { sy = r; sx = j; }
        else if (k == 1) { sy = j; sx = w - 1 - r; }
        else if (k == 2) { sy = h - 1 - r; sx = w - 1 - j; }
        else { sy = h - 1 - j; sx = r; }
        const uint8_t *p = in + (size_t)b * in_fstride + (size_t)sy * in_pitch + (size_t)sx * channels;
        uint8_t *o = out + (size_t)b * out_fstride + (size_t)r * out_pitch + (size_t)j * channels;
        for (int c = 0; c < channels; c++) o[c] = p[c];
    }
}

extern "C" int va_rot90_u8(va_ctx *ctx, va_stream stream,
                           const uint8_t *in, size_t in_pitch, size_t in_fstride,
                           uint8_t *out, size_t out_pitch, size_t out_fstride,
                           int w, int h, int channels, int batch, int k) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_rot90_u8: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && (channels == 1 || channels == 3), "va_rot90_u8: bad size");
    VA_REQUIRE(ctx, k >= 0 && k <= 3, "va_rot90_u8: k must be 0..3");
    const long long items = (long long)w * h * batch;
    VA_REQUIRE(ctx, items < (1ll << 31), "va_rot90_u8: batch too large for one launch");
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = rot90_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, channels, batch, k);
    return VA_OK;
}

// =================================================================================
// measure_mean / measure_mean_std (video/analysis/video.py:26-55), float64 state, the
// reference's operation order (no FMA contraction):
//   mean only:  mean = mean * n / (n + 1) + frame / (n + 1)
//   with M2:    delta = frame - mean;  mean = mean + delta / (n + 1);  M2 = M2 + delta * (frame - mean)
// n0 = number of frames already folded into the state.
// =================================================================================
__global__ void __launch_bounds__(256)
mean_update_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                   double *__restrict__ mean, double *__restrict__ m2, size_t pitch_e,
                   int row_elems, int h, int batch, long long n0) {
    const unsigned total = (unsigned)row_elems * (unsigned)h;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned y = i / (unsigned)row_elems, x = i - y * (unsigned)row_elems;
        const uint8_t *p = in + (size_t)y * in_pitch + x;
        const size_t o = (size_t)y * pitch_e + x;
        double mu = mean[o];
        if (m2) {
            double q = m2[o];
            for (int t = 0; t < batch; t++) {
                const double f = (double)p[(size_t)t * in_fstride];
                const double np1 = (double)(n0 + t + 1);
                const double delta = __dsub_rn(f, mu);
                mu = __dadd_rn(mu, __ddiv_rn(delta, np1));
                q = __dadd_rn(q, __dmul_rn(delta, __dsub_rn(f, mu)));
            }
            m2[o] = q;
        } else {
            for (int t = 0; t < batch; t++) {
                const double f = (double)p[(size_t)t * in_fstride];
                const double n = (double)(n0 + t), np1 = (double)(n0 + t + 1);
                mu = __dadd_rn(__ddiv_rn(__dmul_rn(mu, n), np1), __ddiv_rn(f, np1));
            }
        }
        mean[o] = mu;
    }
}

extern "C" int va_mean_update_f64(va_ctx *ctx, va_stream stream,
                                  const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                  double *mean, double *m2, size_t pitch_e,
                                  int row_elems, int h, int batch, long long n0) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && mean, "va_mean_update_f64: null pointer");
    VA_REQUIRE(ctx, row_elems > 0 && h > 0 && batch > 0 && n0 >= 0, "va_mean_update_f64: bad size");
    const long long items = (long long)row_elems * h;
    VA_REQUIRE(ctx, items < (1ll << 31), "va_mean_update_f64: frame too large");
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = mean_update_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, mean, m2, pitch_e, row_elems, h, batch, n0);
    return VA_OK;
}


// ---------------------------------------------------------------------------------
// FilterResize beyond the 1/2 case (video/filters.py:308-315): INTER_AREA with integer scale factors
// and INTER_NEAREST.  One thread per output byte (plain gathers; these are not on the bench path).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resize_area_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                   uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                   int ow, int oh, int cs, int batch, int kx, int ky, float scale) {
    const unsigned rowb = (unsigned)(ow * cs);
    const unsigned long long total = (unsigned long long)rowb * oh * batch;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned xb = (unsigned)(i % rowb);
        const unsigned long long rest = i / rowb;
        const unsigned y = (unsigned)(rest % oh), b = (unsigned)(rest / oh);
        const unsigned x = xb / cs, c = xb - x * cs;
        const uint8_t *p = in + (size_t)b * in_fstride + (size_t)(y * ky) * in_pitch + (size_t)(x * kx) * cs + c;
        int sum = 0;
        for (int j = 0; j < ky; j++, p += in_pitch)
            for (int k = 0; k < kx; k++) sum += p[k * cs];
        // OpenCV: 2 x 2 has its own integer form, everything else goes through float(sum) * scale, cvRound
        const int v = (kx == 2 && ky == 2) ? (sum + 2) >> 2 : __float2int_rn(__fmul_rn((float)sum, scale));
        out[(size_t)b * out_fstride + (size_t)y * out_pitch + xb] = (uint8_t)v;
    }
}

extern "C" int va_resize_area_u8(va_ctx *ctx, va_stream stream,
                                 const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                 uint8_t *out, size_t out_pitch, size_t out_fstride,
                                 int w, int h, int channels, int batch, int kx, int ky) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_resize_area_u8: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && (channels == 1 || channels == 3), "va_resize_area_u8: bad size");
    VA_REQUIRE(ctx, kx >= 1 && ky >= 1 && kx * ky > 1 && (long long)kx * ky <= 8000000, "va_resize_area_u8: bad factors %d x %d", kx, ky);   // sums stay in 31 bits
    VA_REQUIRE(ctx, w % kx == 0 && h % ky == 0, "va_resize_area_u8: %dx%d is not a multiple of %dx%d", w, h, kx, ky);
    const int ow = w / kx, oh = h / ky;
    VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)ow * channels, "va_resize_area_u8: pitch smaller than a row");
    const long long items = (long long)ow * channels * oh * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 16);
    auto kfn = resize_area_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, ow, oh, channels, batch,
              kx, ky, 1.f / (float)(kx * ky));
    return VA_OK;
}

// thread = 4 consecutive output bytes of one row (one 4-byte store), grid = (groups of a row, rows, frames): no 64-bit
// index arithmetic, the row's source pointer once per thread (one thread per output byte with a flat 64-bit index: 0.150 ms
// per 32 frames 1080p -> 720p)
template <int CS>
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                      uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                      int w, int h, int ow, int oh, double ifx, double ify, int vec) {
    const unsigned rowb = (unsigned)(ow * CS);
    const unsigned xb0 = 4u * (blockIdx.x * blockDim.x + threadIdx.x);
    const unsigned y = blockIdx.y, b = blockIdx.z;
    if (xb0 >= rowb) return;
    const int sy = min((int)floor(__dmul_rn((double)y, ify)), h - 1);
    const uint8_t *srow = in + (size_t)b * in_fstride + (size_t)sy * in_pitch;
    uint8_t *orow = out + (size_t)b * out_fstride + (size_t)y * out_pitch;
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned xb = xb0 + k;
        if (xb < rowb) {
            const unsigned x = xb / CS, c = xb - x * CS;
            const int sx = min((int)floor(__dmul_rn((double)x, ifx)), w - 1);
            v |= (unsigned)srow[(size_t)sx * CS + c] << (8 * k);
        }
    }
    if (vec && xb0 + 4 <= rowb) {
        *reinterpret_cast<unsigned *>(orow + xb0) = v;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (xb0 + k < rowb) orow[xb0 + k] = (uint8_t)(v >> (8 * k));
    }
}

extern "C" int va_resize_nearest_u8(va_ctx *ctx, va_stream stream,
                                    const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                    uint8_t *out, size_t out_pitch, size_t out_fstride,
                                    int w, int h, int dw, int dh, int channels, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_resize_nearest_u8: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && dw > 0 && dh > 0 && batch > 0 && (channels == 1 || channels == 3), "va_resize_nearest_u8: bad size");
    VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)dw * channels, "va_resize_nearest_u8: pitch smaller than a row");
    VA_REQUIRE(ctx, dh <= 65535 && batch <= 65535, "va_resize_nearest_u8: more than 65535 rows or frames");
    // cv::resize: inv_scale = dsize / ssize, ifx = 1 / inv_scale (doubles), x_ofs[x] = min(cvFloor(x * ifx), ssize - 1)
    const double ifx = 1.0 / ((double)dw / (double)w), ify = 1.0 / ((double)dh / (double)h);
    const int vec = va_aligned(out, 4) && out_pitch % 4 == 0 && out_fstride % 4 == 0;
    const dim3 grid(va_div_up(va_div_up(dw * channels, 4), 256), dh, batch);
    if (channels == 1) {
        auto kfn = resize_nearest_kernel<1>;
        VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, dw, dh, ifx, ify, vec);
    } else {
        auto kfn = resize_nearest_kernel<3>;
        VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, dw, dh, ifx, ify, vec);
    }
    return VA_OK;
}


// ---------------------------------------------------------------------------------
// INTER_AREA shrinking by non-integer factors (OpenCV's computeResizeAreaTab + resizeArea_<uchar, float>):
// destination cell dx covers source [dx s, dx s + s), s = 1 / (dsize / ssize) in doubles; the cell's table is
// an optional partial pixel on the left (if it covers more than 1e-3), the whole pixels with weight
// float(1 / cell width), an optional partial pixel on the right.  A destination value is
//     sum over rows in table order of  beta * (sum over columns in table order of  S * alpha)
// in float32, every product and every sum rounded on its own (this is what the x86 build of cv2 4.13 does;
// checked bit for bit), then rounded half to even.  One thread per output byte; the table of its cell is
// recomputed in registers with the same double operations, so nothing is staged.
// ---------------------------------------------------------------------------------
struct AreaCell {
    int s0;                  // first source index of the table
    int n;                   // entries
    float a_first, a_mid, a_last;
    int has_first, has_last;
};

__device__ __forceinline__ AreaCell area_cell(int d, double scale, int ssize) {
    AreaCell c;
    const double f1 = __dmul_rn((double)d, scale);
    const double f2 = __dadd_rn(f1, scale);
    const double cell = fmin(scale, (double)ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    c.has_first = (double)s1 - f1 > 1e-3;
    c.has_last = f2 - (double)s2 > 1e-3;
    c.a_first = (float)__ddiv_rn((double)s1 - f1, cell);
    c.a_mid = (float)__ddiv_rn(1.0, cell);
    c.a_last = (float)__ddiv_rn(fmin(fmin(f2 - (double)s2, 1.0), cell), cell);
    c.s0 = c.has_first ? s1 - 1 : s1;
    c.n = (s2 - s1) + c.has_first + c.has_last;
    return c;
}

__device__ __forceinline__ float area_weight(const AreaCell &c, int i) {
    if (i == 0 && c.has_first) return c.a_first;
    if (i == c.n - 1 && c.has_last) return c.a_last;
    return c.a_mid;
}

// tables: one cell per destination column (entries 0 .. ow - 1) and per destination row (entries ow .. ow + oh - 1),
// computed once per launch by resize_area_tables_kernel with the double operations above (FP64 is slow on this part:
// doing them per output byte was what bounded the old one-thread-per-byte kernel)
__global__ void __launch_bounds__(256)
resize_area_tables_kernel(AreaCell *__restrict__ tab, int w, int h, int ow, int oh, double scale_x, double scale_y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ow) tab[i] = area_cell(i, scale_x, w);
    else if (i < ow + oh) tab[i] = area_cell(i - ow, scale_y, h);
}

// one thread per 4 consecutive output bytes of a row (one 32-bit store); grid = (row words, rows, frames)
__global__ void __launch_bounds__(128)
resize_area_any_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                       uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                       int ow, int oh, int cs, const AreaCell *__restrict__ tab, int vec_out) {
    const int rowb = ow * cs;
    const int xb0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (xb0 >= rowb) return;
    const int y = blockIdx.y, b = blockIdx.z;
    const AreaCell cy = tab[ow + y];
    const uint8_t *frame = in + (size_t)b * in_fstride + (size_t)cy.s0 * in_pitch;
    unsigned packed = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int xb = xb0 + q;
        if (xb >= rowb) break;
        const int x = cs == 1 ? xb : xb / 3, c = cs == 1 ? 0 : xb - 3 * x;
        const AreaCell cx = tab[x];
        const uint8_t *p = frame + (size_t)cx.s0 * cs + c;
        float sum = 0.f;
        for (int j = 0; j < cy.n; j++, p += in_pitch) {
            float buf = 0.f;
            for (int k = 0; k < cx.n; k++) buf = __fadd_rn(buf, __fmul_rn((float)p[k * cs], area_weight(cx, k)));
            const float t = __fmul_rn(area_weight(cy, j), buf);
            sum = j == 0 ? t : __fadd_rn(sum, t);
        }
        const int v = __float2int_rn(sum);
        packed |= (unsigned)min(max(v, 0), 255) << (8 * q);
    }
    uint8_t *o = out + (size_t)b * out_fstride + (size_t)y * out_pitch + xb0;
    if (vec_out && xb0 + 4 <= rowb) {
        *reinterpret_cast<unsigned *>(o) = packed;
    } else {
        for (int q = 0; q < 4 && xb0 + q < rowb; q++) o[q] = (uint8_t)(packed >> (8 * q));
    }
}

// scratch for the per-launch coefficient tables (ctx-owned, grown on demand, ordered across streams by event 4)
static int resize_tables(va_ctx *ctx, const char *name, size_t bytes, void **tab) {
    if (ctx->rs_tab_bytes < bytes) {
        if (ctx->rs_tab) {
            VA_CUDA(ctx, cudaDeviceSynchronize());
            cudaFree(ctx->rs_tab);
            ctx->rs_tab = nullptr;
            ctx->rs_tab_bytes = 0;
        }
        const size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        if (cudaMalloc(&ctx->rs_tab, want) != cudaSuccess) {
            cudaGetLastError();
            VA_FAIL(ctx, VA_ERR_NOMEM, "%s: cannot allocate the coefficient tables", name);
        }
        ctx->rs_tab_bytes = want;
    }
    *tab = ctx->rs_tab;
    return VA_OK;
}

static int resize_linear_launch(va_ctx *ctx, va_stream stream, const char *name,
                                const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                uint8_t *out, size_t out_pitch, size_t out_fstride,
                                int w, int h, int dw, int dh, int channels, int batch,
                                double sx, double sy, int area_mode, double isx, double isy);
template <int MODE>
static int resize_staged_launch(va_ctx *ctx, va_stream stream, const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                uint8_t *out, size_t out_pitch, size_t out_fstride,
                                int w, int h, int dw, int dh, int channels, int batch, double sx, double sy, const void *tab);

// cv::resize's test for the integer-factor INTER_AREA path (both |scale - round(scale)| < DBL_EPSILON)
static bool resize_area_is_fast(int w, int h, int dw, int dh, double *sx, double *sy, int *kx, int *ky) {
    *sx = 1.0 / ((double)dw / (double)w);
    *sy = 1.0 / ((double)dh / (double)h);
    *kx = (int)nearbyint(*sx);
    *ky = (int)nearbyint(*sy);
    return fabs(*sx - *kx) < 2.220446049250313e-16 && fabs(*sy - *ky) < 2.220446049250313e-16;
}

extern "C" int va_resize_area_any_u8(va_ctx *ctx, va_stream stream,
                                     const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                     uint8_t *out, size_t out_pitch, size_t out_fstride,
                                     int w, int h, int dw, int dh, int channels, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_resize_area_any_u8: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && dw > 0 && dh > 0 && batch > 0 && (channels == 1 || channels == 3), "va_resize_area_any_u8: bad size");
    double sx, sy;
    int kx, ky;
    if (dw > w || dh > h) {        // cv::resize: INTER_AREA needs both factors >= 1, otherwise it interpolates linearly (area rule)
        VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)dw * channels, "va_resize_area_any_u8: pitch smaller than a row");
        const double isx = (double)dw / (double)w, isy = (double)dh / (double)h;
        sx = 1.0 / isx;
        sy = 1.0 / isy;
        return resize_linear_launch(ctx, stream, "va_resize_area_any_u8", in, in_pitch, in_fstride, out, out_pitch, out_fstride,
                                    w, h, dw, dh, channels, batch, sx, sy, 1, isx, isy);
    }
    if (resize_area_is_fast(w, h, dw, dh, &sx, &sy, &kx, &ky)) {
        if (kx == 1 && ky == 1)
            return va_copy2d_u8(ctx, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w * channels, h, batch);
        if (kx == 2 && ky == 2)
            return va_resize_half_u8(ctx, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, channels, batch);
        return va_resize_area_u8(ctx, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, channels, batch, kx, ky);
    }
    VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)dw * channels, "va_resize_area_any_u8: pitch smaller than a row");
    VA_REQUIRE(ctx, dh <= 65535 && batch <= 65535, "va_resize_area_any_u8: too many rows or frames for one launch");
    void *tab;
    { const int rc = resize_tables(ctx, "va_resize_area_any_u8", (size_t)(dw + dh) * sizeof(AreaCell), &tab); if (rc != VA_OK) return rc; }
    VA_REQUIRE(ctx, va_scratch_acquire(ctx, stream, 4) == 0, "va_resize_area_any_u8: cannot order the table scratch");
    { auto kfn = resize_area_tables_kernel;
      VA_LAUNCH(ctx, kfn, va_div_up(dw + dh, 256), 256, 0, stream, (AreaCell *)tab, w, h, dw, dh, sx, sy); }
    const int rs = resize_staged_launch<0>(ctx, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, dw, dh, channels, batch,
                                           sx, sy, tab);
    if (rs != VA_OK && rs != VA_ERR_UNSUPPORTED) return rs;
    if (rs == VA_ERR_UNSUPPORTED) {
      auto kfn = resize_area_any_kernel;
      const int words = va_div_up(dw * channels, 4);
      const dim3 grid(va_div_up(words, 128), dh, batch);
      const int vec_out = va_aligned(out, 4) && out_pitch % 4 == 0 && out_fstride % 4 == 0;
      VA_LAUNCH(ctx, kfn, grid, 128, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, dw, dh, channels,
                (const AreaCell *)tab, vec_out); }
    VA_REQUIRE(ctx, va_scratch_release(ctx, stream, 4) == 0, "va_resize_area_any_u8: cannot order the table scratch");
    return VA_OK;
}

// ---------------------------------------------------------------------------------
// INTER_LINEAR (OpenCV's 8-bit fixed-point path: 11-bit coefficients, HResizeLinear then VResizeLinear):
//     fx = float((dx + .5) scale - .5), sx = floor(fx), fx -= sx; columns left of the image use sx = 0, fx = 0,
//     columns at or beyond the last source pixel use that pixel alone; a = (rint((1 - fx) 2048), rint(fx 2048))
//     rows: the same without the reset, row indices clipped instead
//     H(row) = S[sx] a0 + S[sx + 1] a1;   out = (((b0 (H0 >> 4)) >> 16) + ((b1 (H1 >> 4)) >> 16) + 2) >> 2
// Shrinking by exactly 2 x 2 is INTER_AREA in OpenCV ((a + b + c + d + 2) >> 2): the host dispatches it.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void linear_coef(int d, double scale, float &f, int &s) {
    f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    s = (int)floorf(f);
    f = __fadd_rn(f, -(float)s);
}
// cv::resize with INTER_AREA when either direction enlarges: the linear path with its own coefficient rule,
//     s = floor(d scale);  f = float((d + 1) - (s + 1) inv_scale);  f = f <= 0 ? 0 : f - floor(f)
__device__ __forceinline__ void linear_coef_area(int d, double scale, double inv_scale, float &f, int &s) {
    s = (int)floor(__dmul_rn((double)d, scale));
    f = (float)__dadd_rn((double)(d + 1), -__dmul_rn((double)(s + 1), inv_scale));
    f = f <= 0.f ? 0.f : __fadd_rn(f, -floorf(f));
}

struct LinCell { int s0, s1, a0, a1; };      // two source indices (clipped) and their 11-bit weights

// tables: columns 0 .. ow - 1 (s1 = s0 + 1 or s0 at the last pixel, where its weight is not used), rows ow .. ow + oh - 1
__global__ void __launch_bounds__(256)
resize_linear_tables_kernel(LinCell *__restrict__ tab, int w, int h, int ow, int oh, double scale_x, double scale_y,
                            int area_mode, double inv_scale_x, double inv_scale_y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ow + oh) return;
    const bool col = i < ow;
    const int d = col ? i : i - ow;
    float f;
    int sidx;
    if (area_mode) linear_coef_area(d, col ? scale_x : scale_y, col ? inv_scale_x : inv_scale_y, f, sidx);
    else linear_coef(d, col ? scale_x : scale_y, f, sidx);
    LinCell c;
    if (col) {
        if (sidx < 0) { sidx = 0; f = 0.f; }
        if (sidx >= w - 1) { sidx = w - 1; f = 0.f; }
        c.s0 = sidx;
        c.s1 = sidx < w - 1 ? sidx + 1 : -1;             // -1: no second tap (the last pixel alone)
    } else {
        c.s0 = min(max(sidx, 0), h - 1);
        c.s1 = min(max(sidx + 1, 0), h - 1);
    }
    c.a0 = __float2int_rn(__fmul_rn(__fadd_rn(1.f, -f), 2048.f));
    c.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    tab[i] = c;
}

__global__ void __launch_bounds__(128)
resize_linear_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                     uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                     int ow, int oh, int cs, const LinCell *__restrict__ tab, int vec_out) {
    const int rowb = ow * cs;
    const int xb0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (xb0 >= rowb) return;
    const int y = blockIdx.y, b = blockIdx.z;
    const LinCell cy = tab[ow + y];
    const uint8_t *f0 = in + (size_t)b * in_fstride + (size_t)cy.s0 * in_pitch;
    const uint8_t *f1 = in + (size_t)b * in_fstride + (size_t)cy.s1 * in_pitch;
    unsigned packed = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int xb = xb0 + q;
        if (xb >= rowb) break;
        const int x = cs == 1 ? xb : xb / 3, c = cs == 1 ? 0 : xb - 3 * x;
        const LinCell cx = tab[x];
        const size_t o0 = (size_t)cx.s0 * cs + c, o1 = (size_t)(cx.s1 < 0 ? cx.s0 : cx.s1) * cs + c;
        const int h0 = f0[o0] * cx.a0 + (cx.s1 >= 0 ? f0[o1] * cx.a1 : 0);
        const int h1 = f1[o0] * cx.a0 + (cx.s1 >= 0 ? f1[o1] * cx.a1 : 0);
        const int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
        packed |= (unsigned)(v & 0xff) << (8 * q);
    }
    uint8_t *o = out + (size_t)b * out_fstride + (size_t)y * out_pitch + xb0;
    if (vec_out && xb0 + 4 <= rowb) {
        *reinterpret_cast<unsigned *>(o) = packed;
    } else {
        for (int q = 0; q < 4 && xb0 + q < rowb; q++) o[q] = (uint8_t)(packed >> (8 * q));
    }
}

// ---------------------------------------------------------------------------------
// Staged, separable evaluation of INTER_AREA (float tables) and INTER_LINEAR for whole output tiles: a CTA owns
// RS_TH output rows x RS_TWB output bytes.  It stages the source rows / columns the tile's table cells reach into
// shared memory with 16-byte loads, evaluates the horizontal pass ONCE per source row (H[j][xb], the `buf` / `h`
// value of the per-byte kernels above, same operations in the same order), then combines the rows of H per
// output row.  Source bytes are read from global memory once per tile instead of once per output byte and tap,
// and the horizontal pass is shared by the output rows that use the same source row.
// ---------------------------------------------------------------------------------
#define RS_TH 8
#define RS_TWB 512
#define RS_THREADS 256

struct RsTile { int rows_max, cols_max; };     // shared-memory extent in source rows / source bytes (host-computed bound)

template <int MODE>     // 0: INTER_AREA (AreaCell, float), 1: INTER_LINEAR (LinCell, int)
__global__ void __launch_bounds__(RS_THREADS)
resize_staged_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                     uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                     int w, int h, int ow, int oh, int cs, const void *__restrict__ tabv, int vec_in, int vec_out, RsTile tile) {
    VA_DYN_SMEM(unsigned char, smem);
    const AreaCell *ta = reinterpret_cast<const AreaCell *>(tabv);
    const LinCell *tl = reinterpret_cast<const LinCell *>(tabv);
    const int tid = threadIdx.x;
    const int rowb = ow * cs, srcb = w * cs;
    const int xb0 = blockIdx.x * RS_TWB, y0 = blockIdx.y * RS_TH, b = blockIdx.z;
    const int nxb = min(RS_TWB, rowb - xb0), ny = min(RS_TH, oh - y0);
    // source extent of the tile (cells are monotonic in the destination index)
    const int xf = xb0 / cs, xl = (xb0 + nxb - 1) / cs;
    int c_lo, c_hi, r_lo, r_hi;
    if (MODE == 0) {
        c_lo = ta[xf].s0 * cs; c_hi = (ta[xl].s0 + ta[xl].n) * cs;
        r_lo = ta[ow + y0].s0; r_hi = ta[ow + y0 + ny - 1].s0 + ta[ow + y0 + ny - 1].n;
    } else {
        c_lo = tl[xf].s0 * cs; c_hi = (max(tl[xl].s0, tl[xl].s1) + 1) * cs;
        r_lo = min(tl[ow + y0].s0, tl[ow + y0].s1);
        r_hi = max(tl[ow + y0 + ny - 1].s0, tl[ow + y0 + ny - 1].s1) + 1;
        // clipped row indices are monotonic too; an enlarging resize may repeat rows
    }
    const int c_al = c_lo & ~15;                         // staged columns start 16-byte aligned (relative to the row)
    const int ncols = c_hi - c_al, nrows = r_hi - r_lo;
    const int spitch = (tile.cols_max + 31) & ~15;       // bytes per staged row
    unsigned char *S = smem;                             // [rows_max][spitch]
    float *Hf = reinterpret_cast<float *>(smem + (((size_t)tile.rows_max * spitch + 15) & ~(size_t)15));   // [rows_max][RS_TWB]
    int *Hi = reinterpret_cast<int *>(Hf);
    if (nrows > tile.rows_max || ncols > spitch) return;   // cannot happen for the bound the host computed; never write out of bounds
    const uint8_t *src = in + (size_t)b * in_fstride + (size_t)r_lo * in_pitch + c_al;
    // ---- stage
    const int chunks = (ncols + 15) >> 4;
    for (int it = tid; it < nrows * chunks; it += RS_THREADS) {
        const int r = it / chunks, ch = it - r * chunks;
        const uint8_t *g = src + (size_t)r * in_pitch + 16 * ch;
        unsigned char *d = S + (size_t)r * spitch + 16 * ch;
        if (vec_in && c_al + 16 * ch + 16 <= srcb) {
            *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(g);
        } else {
            for (int k = 0; k < 16; k++) d[k] = (c_al + 16 * ch + k < srcb) ? g[k] : (unsigned char)0;
        }
    }
    __syncthreads();
    // ---- horizontal pass: H[j][xb] for every staged source row; a thread keeps the cell of its column and walks the rows
    for (int q = tid; q < nxb; q += RS_THREADS) {
        const int xb = xb0 + q;
        const int x = cs == 1 ? xb : xb / 3, c = cs == 1 ? 0 : xb - 3 * x;
        if (MODE == 0) {
            const AreaCell cx = ta[x];
            const unsigned char *p = S - c_al + c + cx.s0 * cs;
            const int nk = cx.n;
            const float w0 = area_weight(cx, 0), wl = area_weight(cx, nk - 1), wm = cx.a_mid;
            for (int j = 0; j < nrows; j++, p += spitch) {
                float buf = __fadd_rn(0.f, __fmul_rn((float)p[0], w0));
                for (int k = 1; k < nk - 1; k++) buf = __fadd_rn(buf, __fmul_rn((float)p[k * cs], wm));
                if (nk > 1) buf = __fadd_rn(buf, __fmul_rn((float)p[(nk - 1) * cs], wl));
                Hf[j * RS_TWB + q] = buf;
            }
        } else {
            const LinCell cx = tl[x];
            const unsigned char *p0 = S - c_al + c + cx.s0 * cs;
            const unsigned char *p1 = S - c_al + c + (cx.s1 >= 0 ? cx.s1 : cx.s0) * cs;
            const int a1 = cx.s1 >= 0 ? cx.a1 : 0;
            for (int j = 0; j < nrows; j++, p0 += spitch, p1 += spitch) Hi[j * RS_TWB + q] = p0[0] * cx.a0 + p1[0] * a1;
        }
    }
    __syncthreads();
    // ---- vertical pass: 4 output bytes per thread and row
    // a thread owns a word column (4 bytes) and walks every second row of the tile: no index divisions in the loop
    constexpr int WPT = RS_TWB / 4;                       // word columns of a full tile
    const int q0 = 4 * (tid % WPT);
    for (int yy = tid / WPT; yy < ny && q0 < nxb; yy += RS_THREADS / WPT) {
        const int y = y0 + yy;
        unsigned packed = 0;
        if (MODE == 0) {
            const AreaCell cy = ta[ow + y];
            const int j0 = cy.s0 - r_lo;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (q0 + e >= nxb) break;
                float sum = 0.f;
                for (int j = 0; j < cy.n; j++) {
                    const float t = __fmul_rn(area_weight(cy, j), Hf[(j0 + j) * RS_TWB + q0 + e]);
                    sum = j == 0 ? t : __fadd_rn(sum, t);
                }
                const int v = __float2int_rn(sum);
                packed |= (unsigned)min(max(v, 0), 255) << (8 * e);
            }
        } else {
            const LinCell cy = tl[ow + y];
            const int ja = cy.s0 - r_lo, jb = cy.s1 - r_lo;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (q0 + e >= nxb) break;
                const int h0 = Hi[ja * RS_TWB + q0 + e], h1 = Hi[jb * RS_TWB + q0 + e];
                const int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
                packed |= (unsigned)(v & 0xff) << (8 * e);
            }
        }
        uint8_t *o = out + (size_t)b * out_fstride + (size_t)y * out_pitch + xb0 + q0;
        if (vec_out && q0 + 4 <= nxb) {
            *reinterpret_cast<unsigned *>(o) = packed;
        } else {
            for (int e = 0; e < 4 && q0 + e < nxb; e++) o[e] = (uint8_t)(packed >> (8 * e));
        }
    }
}

// launches the staged kernel when the tile's source extent fits in shared memory; returns VA_ERR_UNSUPPORTED otherwise
// (strong shrink factors: the per-byte kernels take over)
template <int MODE>
static int resize_staged_launch(va_ctx *ctx, va_stream stream, const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                uint8_t *out, size_t out_pitch, size_t out_fstride,
                                int w, int h, int dw, int dh, int channels, int batch, double sx, double sy, const void *tab) {
    if (getenv("VA_RESIZE_STAGED") && atoi(getenv("VA_RESIZE_STAGED")) == 0) return VA_ERR_UNSUPPORTED;
    if (dh > 65535 * RS_TH || batch > 65535) return VA_ERR_UNSUPPORTED;
    RsTile tile;
    tile.rows_max = (int)ceil(RS_TH * (sy > 1.0 ? sy : 1.0)) + 3;
    tile.cols_max = ((int)ceil((RS_TWB / channels + 2) * (sx > 1.0 ? sx : 1.0)) + 3) * channels + 16;
    const size_t spitch = ((size_t)tile.cols_max + 31) & ~(size_t)15;
    const size_t smem = (((size_t)tile.rows_max * spitch + 15) & ~(size_t)15) + (size_t)tile.rows_max * RS_TWB * 4;
    if (smem > 100 * 1024) return VA_ERR_UNSUPPORTED;
    auto kfn = resize_staged_kernel<MODE>;
    if (smem > 48 * 1024) VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(va_div_up(dw * channels, RS_TWB), va_div_up(dh, RS_TH), batch);
    const int vec_in = va_aligned(in, 16) && in_pitch % 16 == 0 && in_fstride % 16 == 0;
    const int vec_out = va_aligned(out, 4) && out_pitch % 4 == 0 && out_fstride % 4 == 0;
    VA_LAUNCH(ctx, kfn, grid, RS_THREADS, smem, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, dw, dh, channels,
              tab, vec_in, vec_out, tile);
    return VA_OK;
}

static int resize_linear_launch(va_ctx *ctx, va_stream stream, const char *name,
                                const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                uint8_t *out, size_t out_pitch, size_t out_fstride,
                                int w, int h, int dw, int dh, int channels, int batch,
                                double sx, double sy, int area_mode, double isx, double isy) {
    VA_REQUIRE(ctx, dh <= 65535 && batch <= 65535, "%s: too many rows or frames for one launch", name);
    void *tab;
    { const int rc = resize_tables(ctx, name, (size_t)(dw + dh) * sizeof(LinCell), &tab); if (rc != VA_OK) return rc; }
    VA_REQUIRE(ctx, va_scratch_acquire(ctx, stream, 4) == 0, "%s: cannot order the table scratch", name);
    { auto kfn = resize_linear_tables_kernel;
      VA_LAUNCH(ctx, kfn, va_div_up(dw + dh, 256), 256, 0, stream, (LinCell *)tab, w, h, dw, dh, sx, sy, area_mode, isx, isy); }
    const int rs = resize_staged_launch<1>(ctx, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, dw, dh, channels, batch,
                                           sx, sy, tab);
    if (rs != VA_OK && rs != VA_ERR_UNSUPPORTED) return rs;
    if (rs == VA_ERR_UNSUPPORTED) {
      auto kfn = resize_linear_kernel;
      const int words = va_div_up(dw * channels, 4);
      const dim3 grid(va_div_up(words, 128), dh, batch);
      const int vec_out = va_aligned(out, 4) && out_pitch % 4 == 0 && out_fstride % 4 == 0;
      VA_LAUNCH(ctx, kfn, grid, 128, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, dw, dh, channels,
                (const LinCell *)tab, vec_out); }
    VA_REQUIRE(ctx, va_scratch_release(ctx, stream, 4) == 0, "%s: cannot order the table scratch", name);
    return VA_OK;
}

extern "C" int va_resize_linear_u8(va_ctx *ctx, va_stream stream,
                                   const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                   uint8_t *out, size_t out_pitch, size_t out_fstride,
                                   int w, int h, int dw, int dh, int channels, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_resize_linear_u8: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && dw > 0 && dh > 0 && batch > 0 && (channels == 1 || channels == 3), "va_resize_linear_u8: bad size");
    double sx, sy;
    int kx, ky;
    if (resize_area_is_fast(w, h, dw, dh, &sx, &sy, &kx, &ky) && kx == 2 && ky == 2)
        return va_resize_half_u8(ctx, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, channels, batch);
    VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)dw * channels, "va_resize_linear_u8: pitch smaller than a row");
    return resize_linear_launch(ctx, stream, "va_resize_linear_u8", in, in_pitch, in_fstride, out, out_pitch, out_fstride,
                                w, h, dw, dh, channels, batch, sx, sy, 0, 0.0, 0.0);
}

// ---------------------------------------------------------------------------------
// INTER_CUBIC (what 'auto' picks when enlarging, filters.py:282-284).  OpenCV's own 8-bit path
// (HResizeCubic + VResizeCubic, a = -0.75): coefficients from float32 polynomials scaled to 11 bits,
//     H(row) = sum_j S[clamp(sx - 1 + j)] a_j      (int32)
// and the column pass in float32 -- t = H3 b3; t = H2 b2 + t; t = H1 b1 + t; t = H0 b0 + t with
// b_k = float(beta_k) * 2^-22, every operation rounded on its own, round half to even -- for the first
// 8 floor(row bytes / 8) bytes of a row (its 8-lane vector body) and (sum + 2^21) >> 22 for the tail.
// NB: the cv2 wheel routes INTER_CUBIC through Intel IPP when IPP is enabled (its default); IPP's
// arithmetic is not published and differs from OpenCV's own by at most 1 LSB on a few per cent of the
// pixels.  This kernel is bit-exact against cv2 with cv2.ipp.setUseIPP(False).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coef(int d, double scale, int &s, int (&a)[4]) {
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    s = (int)floorf(f);
    f = __fadd_rn(f, -(float)s);
    const float A = -0.75f;
    const float x1 = __fadd_rn(f, 1.f), xm = __fadd_rn(1.f, -f);
    float c[4];
    c[0] = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(A, x1), -5.f * A), x1), 8.f * A), x1), -4.f * A);
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(A + 2.f, f), -(A + 3.f)), f), f), 1.f);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(A + 2.f, xm), -(A + 3.f)), xm), xm), 1.f);
    c[3] = __fadd_rn(__fadd_rn(__fadd_rn(1.f, -c[0]), -c[1]), -c[2]);
#pragma unroll
    for (int k = 0; k < 4; k++) a[k] = __float2int_rn(__fmul_rn(c[k], 2048.f));
}

__global__ void __launch_bounds__(256)
resize_cubic_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                    uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                    int w, int h, int ow, int oh, int cs, int batch, double scale_x, double scale_y) {
    // grid = (bytes of a row, rows, frames): no 64-bit index arithmetic per output byte
    const unsigned rowb = (unsigned)(ow * cs);
    const unsigned vec_end = rowb & ~7u;
    const unsigned xb = blockIdx.x * blockDim.x + threadIdx.x;
    if (xb < rowb) {
        const unsigned y = blockIdx.y, b = blockIdx.z;
        const unsigned x = xb / cs, c = xb - x * cs;
        int sx, sy, a[4], bt[4];
        cubic_coef((int)x, scale_x, sx, a);
        cubic_coef((int)y, scale_y, sy, bt);
        const uint8_t *f = in + (size_t)b * in_fstride + c;
        int hs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint8_t *row = f + (size_t)min(max(sy - 1 + k, 0), h - 1) * in_pitch;
            int v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) v += row[(size_t)min(max(sx - 1 + j, 0), w - 1) * cs] * a[j];
            hs[k] = v;
        }
        int v;
        if (xb < vec_end) {
            const float sc = 1.f / (2048.f * 2048.f);
            float t = __fmul_rn((float)hs[3], __fmul_rn((float)bt[3], sc));
            t = __fadd_rn(__fmul_rn((float)hs[2], __fmul_rn((float)bt[2], sc)), t);
            t = __fadd_rn(__fmul_rn((float)hs[1], __fmul_rn((float)bt[1], sc)), t);
            t = __fadd_rn(__fmul_rn((float)hs[0], __fmul_rn((float)bt[0], sc)), t);
            v = __float2int_rn(t);
        } else {
            v = (hs[0] * bt[0] + hs[1] * bt[1] + hs[2] * bt[2] + hs[3] * bt[3] + (1 << 21)) >> 22;
        }
        out[(size_t)b * out_fstride + (size_t)y * out_pitch + xb] = (uint8_t)min(max(v, 0), 255);
    }
}

extern "C" int va_resize_cubic_u8(va_ctx *ctx, va_stream stream,
                                  const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                  uint8_t *out, size_t out_pitch, size_t out_fstride,
                                  int w, int h, int dw, int dh, int channels, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_resize_cubic_u8: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && dw > 0 && dh > 0 && batch > 0 && (channels == 1 || channels == 3), "va_resize_cubic_u8: bad size");
    VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)dw * channels, "va_resize_cubic_u8: pitch smaller than a row");
    const double sx = 1.0 / ((double)dw / (double)w), sy = 1.0 / ((double)dh / (double)h);
    VA_REQUIRE(ctx, dh <= 65535 && batch <= 65535, "va_resize_cubic_u8: more than 65535 rows or frames");
    const dim3 grid(va_div_up(dw * channels, 256), dh, batch);
    auto kfn = resize_cubic_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, dw, dh, channels, batch,
              sx, sy);
    return VA_OK;
}

// =================================================================================
// VideoComposer.highlight_mask (video/io/composer.py:131-154): in the pixels of a mask
//     frame[mask, channel] = strength + (255 - strength) / 255 * frame[mask, channel]      (float64, cast to uint8)
// i.e. a 256-entry table applied where the mask bit is set (the table is the reference's expression
// evaluated on 0..255 by the host).  channel = -1 touches every channel, 0..2 one channel of an
// interleaved frame.  The mask comes as the packed bits K3 / K4 leave on the device, so an annotated
// output video needs no mask round trip through the host.  One thread per 4 output bytes.
// =================================================================================
__global__ void __launch_bounds__(256)
highlight_mask_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                      const uint32_t *__restrict__ mask, size_t mask_pitch_w, size_t mask_fstride_w,
                      uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                      int w, int h, int cs, int batch, int channel, int vec, const __grid_constant__ Lut256 lut) {
    __shared__ unsigned char s[256];
    s[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const unsigned rowb = (unsigned)(w * cs);
    const unsigned quads = (rowb + 3) >> 2;
    const unsigned long long total = (unsigned long long)quads * h * batch;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned q = (unsigned)(i % quads);
        const unsigned long long rest = i / quads;
        const unsigned y = (unsigned)(rest % h), b = (unsigned)(rest / h);
        const uint8_t *src = in + (size_t)b * in_fstride + (size_t)y * in_pitch + 4 * q;
        uint8_t *dst = out + (size_t)b * out_fstride + (size_t)y * out_pitch + 4 * q;
        const uint32_t *mrow = mask + (size_t)b * mask_fstride_w + (size_t)y * mask_pitch_w;
        const unsigned nb = min(4u, rowb - 4 * q);
        unsigned word = 0;
        if (vec) word = *reinterpret_cast<const unsigned *>(src);
        else for (unsigned k = 0; k < nb; k++) word |= (unsigned)src[k] << (8 * k);
        unsigned res = 0;
#pragma unroll
        for (unsigned k = 0; k < 4; k++) {
            const unsigned xb = 4 * q + k;
            const unsigned px = cs == 3 ? xb / 3 : xb, c = xb - px * cs;
            unsigned v = (word >> (8 * k)) & 0xff;
            if (k < nb && (channel < 0 || (int)c == channel) && ((mrow[px >> 5] >> (px & 31)) & 1)) v = s[v];
            res |= v << (8 * k);
        }
        if (vec) *reinterpret_cast<unsigned *>(dst) = res;
        else for (unsigned k = 0; k < nb; k++) dst[k] = (uint8_t)(res >> (8 * k));
    }
}

extern "C" int va_highlight_mask_u8(va_ctx *ctx, va_stream stream,
                                    const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                    const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                                    uint8_t *out, size_t out_pitch, size_t out_fstride,
                                    int w, int h, int channels, int batch, int channel, const uint8_t *lut256) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && mask && lut256, "va_highlight_mask_u8: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && (channels == 1 || channels == 3), "va_highlight_mask_u8: bad size");
    VA_REQUIRE(ctx, channel >= -1 && channel < channels && (channels == 3 || channel <= 0),
               "va_highlight_mask_u8: highlighting a specific channel is only supported for color videos");
    VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)w * channels && mask_pitch_w >= (size_t)(w + 31) / 32,
               "va_highlight_mask_u8: pitch smaller than a row");
    Lut256 lut;
    memcpy(lut.v, lut256, 256);
    const int vec = (w * channels) % 4 == 0 && va_aligned(in, 4) && va_aligned(out, 4) && in_pitch % 4 == 0 && out_pitch % 4 == 0 &&
                    in_fstride % 4 == 0 && out_fstride % 4 == 0;
    const long long items = (long long)((w * channels + 3) / 4) * h * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = highlight_mask_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, mask, mask_pitch_w, mask_fstride_w, out, out_pitch, out_fstride,
              w, h, channels, batch, channels == 1 ? -1 : channel, vec, lut);
    return VA_OK;
}

// ---------------------------------------------------------------------------------
// INTER_LANCZOS4 (OpenCV's 8-bit fixed-point path, HResizeLanczos4 + VResizeLanczos4, no vector body):
//     fx = float((dx + .5) scale - .5), sx = floor(fx), fx -= sx;  8 coefficients from interpolateLanczos4 (sines and
//     cosines in doubles, normalised in float32), scaled to 11 bits;  H(row) = sum_j S[clamp(sx - 3 + j)] a_j  (int32);
//     out = saturate((sum_k b_k H(clamp(sy - 3 + k)) + 2^21) >> 22)
// The coefficient tables are computed on the host with the same libm calls OpenCV makes (so they agree to the last
// bit) and travel to the device in a stream-ordered allocation.  Bit-exact against cv2 (IPP does not take Lanczos).
// ---------------------------------------------------------------------------------
static void lanczos4_table(int dsize, int ssize, std::vector<int> &ofs, std::vector<short> &coef) {
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    ofs.resize(dsize);
    coef.resize((size_t)dsize * 8);
    for (int d = 0; d < dsize; d++) {
        volatile float fx = (float)((d + 0.5) * scale - 0.5);
        const int s = (int)std::floor(fx);
        fx = fx - (float)s;
        const float x = fx;
        float c[8];
        {
            // interpolateLanczos4 of OpenCV 4.x: float sums, double sines; a tap that falls on the sample itself gets
            // the weight 1e30 and takes everything after the normalisation
            volatile float sum = 0.f;
            volatile float x3 = x + 3;
            const double y0 = -x3 * 3.1415926535897932384626433832795 * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
            for (int i = 0; i < 8; i++) {
                volatile float yi = x3 - i;
                if (std::fabs(yi) >= 1e-6f) {
                    const double y = -yi * 3.1415926535897932384626433832795 * 0.25;
                    c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
                } else {
                    c[i] = 1e30f;
                }
                sum = sum + c[i];
            }
            volatile float inv = 1.f / sum;
            for (int i = 0; i < 8; i++) { volatile float t = c[i] * inv; c[i] = t; }
        }
        ofs[d] = s;
        for (int i = 0; i < 8; i++) {
            volatile float t = c[i] * 2048.f;
            const long v = std::lrintf(t);                         // saturate_cast<short>(float) = cvRound, saturated
            coef[(size_t)d * 8 + i] = (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v);
        }
    }
}

__global__ void __launch_bounds__(256)
resize_lanczos4_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                       uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                       int w, int h, int ow, int oh, int cs, int batch,
                       const int *__restrict__ xofs, const short *__restrict__ alpha,
                       const int *__restrict__ yofs, const short *__restrict__ beta) {
    // grid = (bytes of a row, rows, frames): no 64-bit index arithmetic per output byte
    const unsigned rowb = (unsigned)(ow * cs);
    const unsigned xb = blockIdx.x * blockDim.x + threadIdx.x;
    if (xb < rowb) {
        const unsigned y = blockIdx.y, b = blockIdx.z;
        const unsigned x = xb / cs, c = xb - x * cs;
        const int sx = xofs[x], sy = yofs[y];
        const short *a = alpha + 8 * (size_t)x, *bt = beta + 8 * (size_t)y;
        const uint8_t *f = in + (size_t)b * in_fstride + c;
        int acc = 1 << 21;
        for (int k = 0; k < 8; k++) {
            const uint8_t *row = f + (size_t)min(max(sy - 3 + k, 0), h - 1) * in_pitch;
            int v = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) v += row[(size_t)min(max(sx - 3 + j, 0), w - 1) * cs] * a[j];
            acc += v * bt[k];
        }
        out[(size_t)b * out_fstride + (size_t)y * out_pitch + xb] = (uint8_t)min(max(acc >> 22, 0), 255);
    }
}

extern "C" int va_resize_lanczos4_u8(va_ctx *ctx, va_stream stream,
                                     const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                     uint8_t *out, size_t out_pitch, size_t out_fstride,
                                     int w, int h, int dw, int dh, int channels, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_resize_lanczos4_u8: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && dw > 0 && dh > 0 && batch > 0 && (channels == 1 || channels == 3), "va_resize_lanczos4_u8: bad size");
    VA_REQUIRE(ctx, in_pitch >= (size_t)w * channels && out_pitch >= (size_t)dw * channels, "va_resize_lanczos4_u8: pitch smaller than a row");
    std::vector<int> xo, yo;
    std::vector<short> xa, ya;
    lanczos4_table(dw, w, xo, xa);
    lanczos4_table(dh, h, yo, ya);
    // one stream-ordered allocation: [xofs | yofs | alpha | beta]
    const size_t n_ofs = (size_t)dw + dh, n_coef = 8 * n_ofs;
    const size_t bytes = n_ofs * sizeof(int) + n_coef * sizeof(short);
    std::vector<unsigned char> host(bytes);
    memcpy(host.data(), xo.data(), dw * sizeof(int));
    memcpy(host.data() + dw * sizeof(int), yo.data(), dh * sizeof(int));
    memcpy(host.data() + n_ofs * sizeof(int), xa.data(), 8 * (size_t)dw * sizeof(short));
    memcpy(host.data() + n_ofs * sizeof(int) + 8 * (size_t)dw * sizeof(short), ya.data(), 8 * (size_t)dh * sizeof(short));
    unsigned char *dev = nullptr;
    VA_CUDA(ctx, cudaMallocAsync((void **)&dev, bytes, (cudaStream_t)stream));
    VA_CUDA(ctx, cudaMemcpyAsync(dev, host.data(), bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));   // pageable source: staged before the call returns
    const int *d_xofs = reinterpret_cast<const int *>(dev), *d_yofs = d_xofs + dw;
    const short *d_alpha = reinterpret_cast<const short *>(dev + n_ofs * sizeof(int)), *d_beta = d_alpha + 8 * (size_t)dw;
    VA_REQUIRE(ctx, dh <= 65535 && batch <= 65535, "va_resize_lanczos4_u8: more than 65535 rows or frames");
    const dim3 grid(va_div_up(dw * channels, 256), dh, batch);
    auto kfn = resize_lanczos4_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, dw, dh, channels, batch,
              d_xofs, d_alpha, d_yofs, d_beta);
    VA_CUDA(ctx, cudaFreeAsync(dev, (cudaStream_t)stream));
    return VA_OK;
}

// =================================================================================
// BASELINE.json configs[3]: concurrent camera streams batched on the leading axis of one launch, every
// stream with its own crop rectangle position (FilterCrop, video/filters.py:158-248, one rectangle size for
// the batch) followed by FilterMonochrome (video/filters.py:359-374):
//     out[s] = mono(in[s][top_s : top_s + h, left_s : left_s + w])          xy = {left_0, top_0, left_1, top_1, ...}
// A crop is addressing only, but per-stream offsets cannot be folded into one base pointer + frame stride, so
// they come as a device table.  One thread per 4 output pixels (12 source bytes, word loads when the
// stream's offset leaves them aligned); the arithmetic is K1's (va_luma_x4).
// =================================================================================
__global__ void __launch_bounds__(256)
luma_crop_multi_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                       uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                       int w, int h, int batch, int mode, const int *__restrict__ xy) {
    const unsigned quads = (unsigned)((w + 3) >> 2);
    const unsigned long long total = (unsigned long long)quads * h * batch;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned q = (unsigned)(i % quads);
        const unsigned long long rest = i / quads;
        const unsigned y = (unsigned)(rest % h), s = (unsigned)(rest / h);
        const int left = xy[2 * s], top = xy[2 * s + 1];
        const uint8_t *p = in + (size_t)s * in_fstride + (size_t)(top + (int)y) * in_pitch + (size_t)(left + 4 * (int)q) * 3;
        uint8_t *o = out + (size_t)s * out_fstride + (size_t)y * out_pitch + 4 * q;
        const int np = min(4, w - 4 * (int)q);
        unsigned wd[3] = {0, 0, 0};
        if (np == 4 && (reinterpret_cast<size_t>(p) & 3) == 0) {
            const unsigned *pw = reinterpret_cast<const unsigned *>(p);
            wd[0] = __ldg(pw); wd[1] = __ldg(pw + 1); wd[2] = __ldg(pw + 2);
        } else {
            for (int k = 0; k < 3 * np; k++) wd[k >> 2] |= (unsigned)__ldg(p + k) << (8 * (k & 3));
        }
        const unsigned r = va_luma_x4(wd[0], wd[1], wd[2], mode);
        if (np == 4 && (reinterpret_cast<size_t>(o) & 3) == 0) *reinterpret_cast<unsigned *>(o) = r;
        else for (int k = 0; k < np; k++) o[k] = (uint8_t)(r >> (8 * k));
    }
}

// w % 16 == 0, word-aligned source rows, 16-byte aligned destination rows: one thread per 16 output pixels.  The 48
// source bytes start at any byte offset (3 * left_s): 13 aligned words, realigned by funnel shifts.
__global__ void __launch_bounds__(256)
luma_crop_multi16_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                         uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                         int w, int h, int batch, int mode, const int *__restrict__ xy) {
    const unsigned groups = (unsigned)(w >> 4);
    const unsigned long long total = (unsigned long long)groups * h * batch;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned g = (unsigned)(i % groups);
        const unsigned long long rest = i / groups;
        const unsigned y = (unsigned)(rest % h), s = (unsigned)(rest / h);
        const int left = xy[2 * s], top = xy[2 * s + 1];
        const uint8_t *p = in + (size_t)s * in_fstride + (size_t)(top + (int)y) * in_pitch + (size_t)(left + 16 * (int)g) * 3;
        const unsigned mis = (unsigned)(reinterpret_cast<size_t>(p) & 3);
        const unsigned *pw = reinterpret_cast<const unsigned *>(p - mis);
        unsigned a[13];
#pragma unroll
        for (int k = 0; k < 12; k++) a[k] = __ldg(pw + k);
        a[12] = mis ? __ldg(pw + 12) : 0u;
        const unsigned sh = 8 * mis;
        unsigned r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned w0 = __funnelshift_r(a[3 * k], a[3 * k + 1], sh);
            const unsigned w1 = __funnelshift_r(a[3 * k + 1], a[3 * k + 2], sh);
            const unsigned w2 = __funnelshift_r(a[3 * k + 2], a[3 * k + 3], sh);
            r[k] = va_luma_x4(w0, w1, w2, mode);
        }
        *reinterpret_cast<uint4 *>(out + (size_t)s * out_fstride + (size_t)y * out_pitch + 16 * g) = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

// The whole front of configs[3] in one pass: crop position per stream + monochrome + static mask (one per stream, or
// one for all with mask_fstride = 0, or none) + threshold -> packed bits.  w % 32 == 0; a lane pair assembles one
// 32-bit mask word.  3N bytes of RGB and N bytes of mask in, N/8 out -- instead of three kernels moving 8.1N.
__global__ void __launch_bounds__(256)
streams_threshold_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                         const uint8_t *__restrict__ smask, size_t smask_pitch, size_t smask_fstride,
                         uint32_t *__restrict__ bits, size_t bits_pitch_w, size_t bits_fstride_w,
                         int w, int h, int batch, int mode, int thr, const int *__restrict__ xy) {
    const unsigned groups = (unsigned)(w >> 4);
    const unsigned long long total = (unsigned long long)groups * h * batch;
    const unsigned lane = threadIdx.x & 31;
    for (unsigned long long base = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) - lane; base < total;
         base += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = base + lane;
        unsigned m16 = 0;
        unsigned g = 0, y = 0, s = 0;
        if (i < total) {
            g = (unsigned)(i % groups);
            const unsigned long long rest = i / groups;
            y = (unsigned)(rest % h);
            s = (unsigned)(rest / h);
            const int left = xy[2 * s], top = xy[2 * s + 1];
            const uint8_t *p = in + (size_t)s * in_fstride + (size_t)(top + (int)y) * in_pitch + (size_t)(left + 16 * (int)g) * 3;
            const unsigned mis = (unsigned)(reinterpret_cast<size_t>(p) & 3);
            const unsigned *pw = reinterpret_cast<const unsigned *>(p - mis);
            unsigned a[13];
#pragma unroll
            for (int k = 0; k < 12; k++) a[k] = __ldg(pw + k);
            a[12] = mis ? __ldg(pw + 12) : 0u;
            const unsigned sh = 8 * mis;
            uint4 mk = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
            if (smask) mk = __ldg(reinterpret_cast<const uint4 *>(smask + (size_t)s * smask_fstride + (size_t)y * smask_pitch + 16 * g));
            const unsigned mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned w0 = __funnelshift_r(a[3 * k], a[3 * k + 1], sh);
                const unsigned w1 = __funnelshift_r(a[3 * k + 1], a[3 * k + 2], sh);
                const unsigned w2 = __funnelshift_r(a[3 * k + 2], a[3 * k + 3], sh);
                const unsigned r = va_luma_x4(w0, w1, w2, mode);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int v = ((mw[k] >> (8 * b)) & 0xffu) ? (int)((r >> (8 * b)) & 0xffu) : 0;     // apply-mask
                    m16 |= (v > thr ? 1u : 0u) << (4 * k + b);
                }
            }
        }
        const unsigned other = __shfl_down_sync(0xffffffffu, m16, 1);
        if (i < total && !(g & 1))
            bits[(size_t)s * bits_fstride_w + (size_t)y * bits_pitch_w + (g >> 1)] = m16 | (other << 16);
    }
}

extern "C" int va_streams_threshold_bits(va_ctx *ctx, va_stream stream,
                                         const uint8_t *in, size_t in_pitch, size_t in_fstride, int in_w, int in_h,
                                         const uint8_t *smask, size_t smask_pitch, size_t smask_fstride,
                                         uint32_t *bits, size_t bits_pitch_w, size_t bits_fstride_w,
                                         int w, int h, int batch, int mode, int thr, const int32_t *xy) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && bits && xy, "va_streams_threshold_bits: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && w <= in_w && h <= in_h, "va_streams_threshold_bits: bad size");
    VA_REQUIRE(ctx, mode >= -1 && mode <= 2, "va_streams_threshold_bits: unsupported conversion method to monochrome: %d", mode);
    VA_REQUIRE(ctx, in_pitch >= (size_t)3 * in_w && bits_pitch_w >= (size_t)(w + 31) / 32 && (!smask || smask_pitch >= (size_t)w),
               "va_streams_threshold_bits: pitch smaller than a row");
    if (!(w % 32 == 0 && va_aligned(in, 4) && in_pitch % 4 == 0 && in_fstride % 4 == 0 &&
          (!smask || (va_aligned(smask, 16) && smask_pitch % 16 == 0 && smask_fstride % 16 == 0))))
        VA_FAIL(ctx, VA_ERR_UNSUPPORTED, "va_streams_threshold_bits: needs w %% 32 == 0, word-aligned frames and 16-byte aligned "
                                         "mask rows (use va_luma_crop_multi_u8 + va_apply_mask_u8 + va_threshold_bits)");
    const long long items = (long long)(w / 16) * h * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 16);
    auto kfn = streams_threshold_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, smask, smask_pitch, smask_fstride, bits, bits_pitch_w,
              bits_fstride_w, w, h, batch, mode, thr, (const int *)xy);
    return VA_OK;
}

extern "C" int va_luma_crop_multi_u8(va_ctx *ctx, va_stream stream,
                                     const uint8_t *in, size_t in_pitch, size_t in_fstride, int in_w, int in_h,
                                     uint8_t *out, size_t out_pitch, size_t out_fstride,
                                     int w, int h, int batch, int mode, const int32_t *xy) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && xy, "va_luma_crop_multi_u8: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && w <= in_w && h <= in_h, "va_luma_crop_multi_u8: bad size");
    VA_REQUIRE(ctx, mode >= -1 && mode <= 2, "va_luma_crop_multi_u8: unsupported conversion method to monochrome: %d", mode);
    VA_REQUIRE(ctx, in_pitch >= (size_t)3 * in_w && out_pitch >= (size_t)w, "va_luma_crop_multi_u8: pitch smaller than a row");
    if (w % 16 == 0 && va_aligned(in, 4) && in_pitch % 4 == 0 && in_fstride % 4 == 0 && va_aligned(out, 16) &&
        out_pitch % 16 == 0 && out_fstride % 16 == 0) {
        const long long items16 = (long long)(w / 16) * h * batch;
        const int grid16 = va_grid(ctx, (items16 + 255) / 256, 16);
        auto k16 = luma_crop_multi16_kernel;
        VA_LAUNCH(ctx, k16, grid16, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, batch, mode,
                  (const int *)xy);
        return VA_OK;
    }
    const long long items = (long long)((w + 3) / 4) * h * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 16);
    auto kfn = luma_crop_multi_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w, h, batch, mode,
              (const int *)xy);
    return VA_OK;
}
