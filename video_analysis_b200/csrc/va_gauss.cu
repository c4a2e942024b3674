// va_gauss.cu -- K2: bit-exact cv2.GaussianBlur on uint8 (video/filters.py:392).
//
// OpenCV's 8-bit Gaussian is fixed-point integer arithmetic:
//   K = 8-bit quantised taps (sum 256), out = (sum_ky sum_kx K[ky] K[kx] src + 32768) >> 16
// with BORDER_REFLECT_101.  The double sum is exact in 32 bits, so any exact
// evaluation order is bit-identical; we run the two 1-D passes inside one CTA and
// keep the 16-bit row-pass image in shared memory (algorithmic HBM bytes: 2N, or
// 4N for the fused RGB -> luma -> blur variant which never writes the luma frame).
//
// Fast path (1 channel, taps <= 255, radius <= 63), per 128 x TH tile:
//   stage   u8 tile + halo -> smem (cp.async 16 B for interior chunks, reflected
//           bytes at the image border; the fused variant computes luma on the fly)
//   rows    dp4a: 4 output pixels x 2 rows per thread against pre-shifted tap words,
//           results packed as (row 2q, row 2q+1) u16 pairs
//   cols    dp2a on those pairs: 4 columns x 2 output rows per thread
// Generic path (3 interleaved channels, tap 256, radius <= 127): same structure with
// scalar multiply-adds.
#include <cmath>

#include "va_device.cuh"

#define GAUSS_MAX_TAPS 255
#define GAUSS_FAST_MAX_R 63
#define GAUSS_TW 128
#define GAUSS_THREADS 256
#define GAUSS_NW_MAX 36
#define GAUSS_NP_MAX 64

// ---------------------------------------------------------------------------------
// host: OpenCV's 8-bit kernel (getGaussianKernel + error-diffused quantisation)
// ---------------------------------------------------------------------------------
int va_gauss_build_taps(double sigma, int *taps, int capacity) {
    if (!(sigma > 0)) return VA_ERR_INVALID;
    int ksize = (int)std::nearbyint(sigma * 6 + 1) | 1;     // cvRound == round-half-even
    if (ksize > capacity) return VA_ERR_CAPACITY;
    // cv::getGaussianKernel(ksize, sigma, CV_64F): exp(-x^2 / (2 sigma^2)), normalised
    double k64[2 * 127 + 1 + 512];
    if (ksize > (int)(sizeof(k64) / sizeof(k64[0]))) return VA_ERR_CAPACITY;
    const double scale2x = -0.5 / (sigma * sigma);
    double sum = 0;
    for (int i = 0; i < ksize; i++) {
        const double x = i - (ksize - 1) * 0.5;
        k64[i] = std::exp(scale2x * x * x);
        sum += k64[i];
    }
    const double inv = 1.0 / sum;
    for (int i = 0; i < ksize; i++) k64[i] *= inv;
    // 8 fractional bits, error diffused from the ends inwards, centre takes the rest
    double err = 0;
    int acc = 0;
    for (int i = 0; i < ksize / 2; i++) {
        const double adj = k64[i] * 256 + err;
        const int q = (int)std::nearbyint(adj);
        taps[i] = taps[ksize - 1 - i] = q;
        err = adj - q;
        acc += 2 * q;
    }
    taps[ksize / 2] = 256 - acc;
    return ksize;
}

extern "C" int va_gauss_taps(double sigma, int *taps, int capacity) {
    if (!taps) return VA_ERR_INVALID;
    return va_gauss_build_taps(sigma, taps, capacity);
}

struct GaussFast {
    int r, o, nw, np;
    unsigned cw[4][GAUSS_NW_MAX];   // row pass: tap bytes for output pixel i, staged word j
    unsigned cp[2][GAUSS_NP_MAX];   // column pass: (K[2j-i], K[2j+1-i]) in the low 16 bits
};
struct GaussGeneric {
    int ksize;
    short k[GAUSS_MAX_TAPS];
};

// ---------------------------------------------------------------------------------
// fast kernel
// ---------------------------------------------------------------------------------
template <bool FUSE_LUMA>
__global__ void __launch_bounds__(GAUSS_THREADS)
gauss_fast_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                  uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                  int w, int h, int mode, int TH, int tiles_x, int tiles_y, int n_tiles,
                  int vec_in, const __grid_constant__ GaussFast g) {
    VA_DYN_SMEM(uint8_t, smem);
    const int tid = threadIdx.x;
    const int r = g.r;
    const int R16 = (r + 15) & ~15;
    const int SW = GAUSS_TW + 2 * R16 + 16;        // staged bytes per row
    const int R = TH + 2 * r;                      // staged rows (even)
    uint8_t *s8 = smem;
    unsigned *hp = reinterpret_cast<unsigned *>(smem + (size_t)R * SW);   // [R/2][TW] u16 pairs
    const bool out_words = (((uintptr_t)out | out_pitch | out_fstride) & 3) == 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y);
        const int rem = tile - b * tiles_x * tiles_y;
        const int tyi = rem / tiles_x;
        const int tx0 = (rem - tyi * tiles_x) * GAUSS_TW;
        const int ty0 = tyi * TH;
        const uint8_t *fin = in + (size_t)b * in_fstride;

        // ---- stage the tile (+ halo)
        if (!FUSE_LUMA) {
            const int chunks = SW >> 4;
            for (int it = tid; it < R * chunks; it += GAUSS_THREADS) {
                const int tr = it / chunks, c = it - tr * chunks;
                const int gy = va_reflect101(ty0 + tr - r, h);
                const uint8_t *rp = fin + (size_t)gy * in_pitch;
                const int gx0 = tx0 - R16 + 16 * c;
                uint8_t *d = s8 + (size_t)tr * SW + 16 * c;
                if (vec_in && gx0 >= 0 && gx0 + 16 <= w) {
                    va_cp_async16(d, rp + gx0);
                } else {
                    unsigned v[4] = {0, 0, 0, 0};
                    for (int i = 0; i < 16; i++)
                        v[i >> 2] |= (unsigned)rp[va_reflect101(gx0 + i, w)] << (8 * (i & 3));
                    *reinterpret_cast<uint4 *>(d) = make_uint4(v[0], v[1], v[2], v[3]);
                }
            }
            va_cp_async_wait_all();
        } else {
            const int groups = SW >> 2;
            for (int it = tid; it < R * groups; it += GAUSS_THREADS) {
                const int tr = it / groups, u = it - tr * groups;
                const int gy = va_reflect101(ty0 + tr - r, h);
                const uint8_t *rp = fin + (size_t)gy * in_pitch;
                const int gx0 = tx0 - R16 + 4 * u;
                unsigned res;
                if (vec_in && gx0 >= 0 && gx0 + 4 <= w) {
                    const unsigned *p = reinterpret_cast<const unsigned *>(rp + 3 * (size_t)gx0);
                    res = va_luma_x4(__ldg(p), __ldg(p + 1), __ldg(p + 2), mode);
                } else {
                    res = 0;
                    for (int i = 0; i < 4; i++)
                        res |= va_luma_px(rp + 3 * (size_t)va_reflect101(gx0 + i, w), mode) << (8 * i);
                }
                *reinterpret_cast<unsigned *>(s8 + (size_t)tr * SW + 4 * u) = res;
            }
        }
        __syncthreads();

        // ---- row pass: item = (row pair q, 4-pixel group gx)
        {
            const int groups = GAUSS_TW >> 2;
            const int wofs = (R16 >> 2) - ((r + 3) >> 2);     // first staged word of group 0
            for (int it = tid; it < (R >> 1) * groups; it += GAUSS_THREADS) {
                const int q = it / groups, gx = it - q * groups;
                const unsigned *r0 = reinterpret_cast<const unsigned *>(s8 + (size_t)(2 * q) * SW) + wofs + gx;
                const unsigned *r1 = reinterpret_cast<const unsigned *>(s8 + (size_t)(2 * q + 1) * SW) + wofs + gx;
                unsigned a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
                for (int j = 0; j < g.nw; j++) {
                    const unsigned x0 = r0[j], x1 = r1[j];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const unsigned c = g.cw[i][j];
                        a0[i] = __dp4a(x0, c, a0[i]);
                        a1[i] = __dp4a(x1, c, a1[i]);
                    }
                }
                *reinterpret_cast<uint4 *>(hp + (size_t)q * GAUSS_TW + 4 * gx) =
                    make_uint4(a0[0] | (a1[0] << 16), a0[1] | (a1[1] << 16), a0[2] | (a1[2] << 16), a0[3] | (a1[3] << 16));
            }
        }
        __syncthreads();

        // ---- column pass: item = (output row pair yp, 4-column group gx)
        {
            const int groups = GAUSS_TW >> 2;
            for (int it = tid; it < (TH >> 1) * groups; it += GAUSS_THREADS) {
                const int yp = it / groups, gx = it - yp * groups;
                const int x = tx0 + 4 * gx;
                const int y = ty0 + 2 * yp;
                if (x >= w || y >= h) continue;
                unsigned acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
                const uint4 *col = reinterpret_cast<const uint4 *>(hp + (size_t)yp * GAUSS_TW + 4 * gx);
                for (int j = 0; j < g.np; j++) {
                    const uint4 v = col[(size_t)j * (GAUSS_TW >> 2)];
                    const unsigned c0 = g.cp[0][j], c1 = g.cp[1][j];
                    acc[0][0] = __dp2a_lo(v.x, c0, acc[0][0]); acc[1][0] = __dp2a_lo(v.x, c1, acc[1][0]);
                    acc[0][1] = __dp2a_lo(v.y, c0, acc[0][1]); acc[1][1] = __dp2a_lo(v.y, c1, acc[1][1]);
                    acc[0][2] = __dp2a_lo(v.z, c0, acc[0][2]); acc[1][2] = __dp2a_lo(v.z, c1, acc[1][2]);
                    acc[0][3] = __dp2a_lo(v.w, c0, acc[0][3]); acc[1][3] = __dp2a_lo(v.w, c1, acc[1][3]);
                }
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    if (y + i >= h) break;
                    const unsigned res = ((acc[i][0] + 32768u) >> 16) | (((acc[i][1] + 32768u) >> 16) << 8) |
                                         (((acc[i][2] + 32768u) >> 16) << 16) | (((acc[i][3] + 32768u) >> 16) << 24);
                    uint8_t *op = out + (size_t)b * out_fstride + (size_t)(y + i) * out_pitch + x;
                    if (out_words && x + 4 <= w) {
                        *reinterpret_cast<unsigned *>(op) = res;
                    } else {
                        for (int k = 0; k < 4 && x + k < w; k++) op[k] = (uint8_t)(res >> (8 * k));
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// generic kernel: any channel count, taps up to 256, scalar arithmetic
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(GAUSS_THREADS)
gauss_generic_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                     uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                     int w, int h, int cs, int T, int tiles_x, int tiles_y, int n_tiles, const __grid_constant__ GaussGeneric g) {
    VA_DYN_SMEM(uint8_t, smem);
    const int tid = threadIdx.x;
    const int r = g.ksize >> 1;
    const int SWp = T + 2 * r;                 // staged pixels per row
    const int R = T + 2 * r;                   // staged rows
    const int SWb = SWp * cs;                  // staged bytes per row
    const int HW = T * cs;                     // row-pass values per row
    uint8_t *s8 = smem;
    unsigned short *hs = reinterpret_cast<unsigned short *>(smem + (((size_t)R * SWb + 15) & ~(size_t)15));

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y);
        const int rem = tile - b * tiles_x * tiles_y;
        const int tyi = rem / tiles_x;
        const int tx0 = (rem - tyi * tiles_x) * T;
        const int ty0 = tyi * T;
        const uint8_t *fin = in + (size_t)b * in_fstride;
        for (int it = tid; it < R * SWb; it += GAUSS_THREADS) {
            const int tr = it / SWb, xb = it - tr * SWb;
            const int px = xb / cs, c = xb - px * cs;
            const int gy = va_reflect101(ty0 + tr - r, h);
            const int gx = va_reflect101(tx0 + px - r, w);
            s8[it] = fin[(size_t)gy * in_pitch + (size_t)gx * cs + c];
        }
        __syncthreads();
        for (int it = tid; it < R * HW; it += GAUSS_THREADS) {
            const int tr = it / HW, xb = it - tr * HW;
            const uint8_t *p = s8 + (size_t)tr * SWb + xb;
            unsigned acc = 0;
            for (int k = 0; k < g.ksize; k++) acc += (unsigned)g.k[k] * p[k * cs];
            hs[it] = (unsigned short)acc;
        }
        __syncthreads();
        for (int it = tid; it < T * HW; it += GAUSS_THREADS) {
            const int ty = it / HW, xb = it - ty * HW;
            const int y = ty0 + ty, x = tx0 + xb / cs;
            if (y >= h || x >= w) continue;
            const unsigned short *p = hs + (size_t)ty * HW + xb;
            unsigned acc = 32768u;
            for (int k = 0; k < g.ksize; k++) acc += (unsigned)g.k[k] * p[(size_t)k * HW];
            out[(size_t)b * out_fstride + (size_t)y * out_pitch + (size_t)tx0 * cs + xb] = (uint8_t)(acc >> 16);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------
static int gauss_launch(va_ctx *ctx, va_stream stream, const char *name, bool fuse,
                        const uint8_t *in, size_t in_pitch, size_t in_fstride,
                        uint8_t *out, size_t out_pitch, size_t out_fstride,
                        int w, int h, int channels, int batch, int mode, double sigma) {
    VA_REQUIRE(ctx, in && out && in != out, "%s: null or aliased pointers", name);
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "%s: bad size %dx%dx%d", name, w, h, batch);
    VA_REQUIRE(ctx, channels == 1 || channels == 3, "%s: channels must be 1 or 3", name);
    VA_REQUIRE(ctx, sigma > 0, "%s: sigma must be positive", name);
    VA_REQUIRE(ctx, mode >= -1 && mode <= 2, "%s: unsupported conversion method to monochrome: %d", name, mode);
    int taps[GAUSS_MAX_TAPS];
    const int ksize = va_gauss_build_taps(sigma, taps, GAUSS_MAX_TAPS);
    if (ksize < 0) VA_FAIL(ctx, VA_ERR_UNSUPPORTED, "%s: sigma %g needs more than %d taps", name, sigma, GAUSS_MAX_TAPS);
    const int r = ksize / 2;
    int kmax = 0;
    for (int i = 0; i < ksize; i++) kmax = taps[i] > kmax ? taps[i] : kmax;

    if (ksize == 1 && !fuse) {   // identity
        return va_copy2d_u8(ctx, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride, w * channels, h, batch);
    }

    const bool fast = (fuse || channels == 1) && kmax <= 255 && r <= GAUSS_FAST_MAX_R && r >= 1;
    if (fast) {
        GaussFast g;
        memset(&g, 0, sizeof(g));
        g.r = r;
        g.o = (4 - (r & 3)) & 3;
        g.nw = (g.o + 2 * r + 4 + 3) / 4;
        g.np = r + 1;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < g.nw; j++) {
                unsigned wd = 0;
                for (int bb = 0; bb < 4; bb++) {
                    const int k = 4 * j + bb - g.o - i;
                    if (k >= 0 && k < ksize) wd |= (unsigned)taps[k] << (8 * bb);
                }
                g.cw[i][j] = wd;
            }
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < g.np; j++) {
                const int k0 = 2 * j - i, k1 = 2 * j + 1 - i;
                unsigned wd = 0;
                if (k0 >= 0 && k0 < ksize) wd |= (unsigned)taps[k0];
                if (k1 >= 0 && k1 < ksize) wd |= (unsigned)taps[k1] << 8;
                g.cp[i][j] = wd;
            }
        const int TH = r <= 16 ? 64 : 96;
        const int R16 = (r + 15) & ~15;
        const int SW = GAUSS_TW + 2 * R16 + 16;
        const int R = TH + 2 * r;
        const size_t smem = (size_t)R * SW + (size_t)(R / 2) * GAUSS_TW * 4;
        const int tiles_x = va_div_up(w, GAUSS_TW), tiles_y = va_div_up(h, TH);
        const int n_tiles = tiles_x * tiles_y * batch;
        const int ctas = smem <= 36 * 1024 ? 6 : (smem <= 56 * 1024 ? 4 : (smem <= 110 * 1024 ? 2 : 1));
        const int grid = va_grid(ctx, n_tiles, ctas);
        int vec_in;
        if (fuse) {
            vec_in = va_aligned(in, 4) && in_pitch % 4 == 0 && in_fstride % 4 == 0;
            auto kfn = gauss_fast_kernel<true>;
            if (smem > 48 * 1024)
                VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VA_LAUNCH(ctx, kfn, grid, GAUSS_THREADS, smem, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
                      w, h, mode, TH, tiles_x, tiles_y, n_tiles, vec_in, g);
        } else {
            vec_in = va_aligned(in, 16) && in_pitch % 16 == 0 && in_fstride % 16 == 0;
            auto kfn = gauss_fast_kernel<false>;
            if (smem > 48 * 1024)
                VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VA_LAUNCH(ctx, kfn, grid, GAUSS_THREADS, smem, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
                      w, h, mode, TH, tiles_x, tiles_y, n_tiles, vec_in, g);
        }
        return VA_OK;
    }

    if (fuse) VA_FAIL(ctx, VA_ERR_UNSUPPORTED, "%s: sigma %g is outside the fused kernel's range", name, sigma);
    VA_REQUIRE(ctx, r <= 127, "%s: radius %d too large", name, r);
    GaussGeneric gg;
    memset(&gg, 0, sizeof(gg));
    gg.ksize = ksize;
    for (int i = 0; i < ksize; i++) gg.k[i] = (short)taps[i];
    int T = 64;
    size_t smem = 0;
    for (; T >= 8; T >>= 1) {
        const size_t R = T + 2 * r;
        smem = ((R * R * channels + 15) & ~(size_t)15) + R * T * channels * 2;
        if (smem <= 96 * 1024 || (T == 8 && smem <= 200 * 1024)) break;
    }
    if (T < 8) VA_FAIL(ctx, VA_ERR_UNSUPPORTED, "%s: sigma %g with %d channels does not fit in shared memory", name, sigma, channels);
    const int tiles_x = va_div_up(w, T), tiles_y = va_div_up(h, T);
    const int n_tiles = tiles_x * tiles_y * batch;
    auto kfn = gauss_generic_kernel;
    if (smem > 48 * 1024)
        VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = va_grid(ctx, n_tiles, smem <= 48 * 1024 ? 4 : 2);
    VA_LAUNCH(ctx, kfn, grid, GAUSS_THREADS, smem, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
              w, h, channels, T, tiles_x, tiles_y, n_tiles, gg);
    return VA_OK;
}

extern "C" int va_gauss_u8(va_ctx *ctx, va_stream stream,
                           const uint8_t *in, size_t in_pitch, size_t in_fstride,
                           uint8_t *out, size_t out_pitch, size_t out_fstride,
                           int w, int h, int channels, int batch, double sigma) {
    VA_CHECK_CTX(ctx);
    return gauss_launch(ctx, stream, "va_gauss_u8", false, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
                        w, h, channels, batch, VA_MONO_MEAN, sigma);
}

extern "C" int va_luma_gauss_u8(va_ctx *ctx, va_stream stream,
                                const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                uint8_t *out, size_t out_pitch, size_t out_fstride,
                                int w, int h, int batch, int mode, double sigma) {
    VA_CHECK_CTX(ctx);
    return gauss_launch(ctx, stream, "va_luma_gauss_u8", true, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
                        w, h, 1, batch, mode, sigma);
}
