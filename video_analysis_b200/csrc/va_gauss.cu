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
#include <cstdlib>

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

int va_gauss_mma_launch(va_ctx *ctx, va_stream stream, const char *name, bool fuse,
                        const uint8_t *in, size_t in_pitch, size_t in_fstride,
                        uint8_t *out, size_t out_pitch, size_t out_fstride,
                        int w, int h, int batch, int mode, const int *taps, int ksize);

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
// fast kernel.  RT > 0: radius known at compile time (tap loops fully unrolled, taps read
// straight from the constant bank, index divisions by constants); RT == 0: any radius.
// Work split: lane = 4-pixel column group (32 groups = 128 pixels), warp = row (pair).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int gauss_reflect_fast(int i, int n) {
    if ((unsigned)i >= (unsigned)n) {
        i = i < 0 ? -i : 2 * n - 2 - i;
        if ((unsigned)i >= (unsigned)n) i = va_reflect101(i, n);
    }
    return i;
}

// one bounce is enough when the halo is shorter than the image (r < n); `tiny` selects the
// general form for images smaller than the halo
__device__ __forceinline__ int gauss_reflect_row(int i, int n, bool tiny) {
    if (tiny) return va_reflect101(i, n);
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    return i < 0 ? 0 : i;          // rows further than a halo below the image are never used: any valid row will do
}

template <int RT, bool FUSE_LUMA>
__global__ void __launch_bounds__(GAUSS_THREADS, (RT > 0 && RT <= 8) ? 6 : 1)
gauss_fast_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                  uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                  int w, int h, int mode, int TH, int vec_in, int dbg, const __grid_constant__ GaussFast g) {
    // grid = (tiles in x, tiles in y, frames): one 128 x TH tile per CTA
    VA_DYN_SMEM(uint8_t, smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = RT ? RT : g.r;
    const int R16 = (r + 15) & ~15;
    const int SW = GAUSS_TW + 2 * R16 + 16;                 // staged bytes per row
    const int O = (4 - (r & 3)) & 3;
    const int NW = (O + 2 * r + 4 + 3) >> 2;                // staged words feeding one 4-pixel group
    const int NP = r + 1;                                   // row pairs feeding one output row pair
    const int WOFS = (R16 >> 2) - ((r + 3) >> 2);           // first staged word needed by group 0
    const int NWORDS = 32 + NW;                             // staged words the row pass reads per row
    const int R = TH + 2 * r;                               // staged rows (even)
    uint8_t *s8 = smem;
    unsigned *hp = reinterpret_cast<unsigned *>(smem + (size_t)R * SW);   // [R/2][TW] u16 pairs
    const bool out_words = (((uintptr_t)out | out_pitch | out_fstride) & 3) == 0;

    const int tx0 = blockIdx.x * GAUSS_TW;
    const int ty0 = blockIdx.y * TH;
    const uint8_t *fin = in + (size_t)blockIdx.z * in_fstride;
    // image columns the row pass really reads for this tile (anything else is staged as 0)
    const int need_lo = tx0 - r, need_hi = tx0 + GAUSS_TW + r;

    // ---- stage the tile (+ halo): only the staged words the row pass will read
    if (dbg & 1) {
    } else if (!FUSE_LUMA) {
        const int c0 = WOFS >> 2;                              // first / last+1 16-byte chunk
        const int c1 = (WOFS + NWORDS + 3) >> 2;
        const int NCH = c1 - c0;
        for (int it = tid; it < R * NCH; it += GAUSS_THREADS) {
            const int tr = it / NCH, c = c0 + (it - tr * NCH);
            const int gy = gauss_reflect_fast(ty0 + tr - r, h);
            const uint8_t *rp = fin + (size_t)gy * in_pitch;
            const int gx0 = tx0 - R16 + 16 * c;
            uint8_t *d = s8 + tr * SW + 16 * c;
            if (vec_in && gx0 >= 0 && gx0 + 16 <= w) {
                va_cp_async16(d, rp + gx0);
            } else {
                unsigned v[4] = {0, 0, 0, 0};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int gx = gx0 + i;
                    if (gx >= need_lo && gx < need_hi)
                        v[i >> 2] |= (unsigned)rp[gauss_reflect_fast(gx, w)] << (8 * (i & 3));
                }
                *reinterpret_cast<uint4 *>(d) = make_uint4(v[0], v[1], v[2], v[3]);
            }
        }
        va_cp_async_wait_all();
    } else {
        // 8 pixels (24 bytes of RGB) per item -> two staged words.
        // interior (the tile's own 128 columns): half a warp covers one row with 16 coalesced
        // items, so there is no index arithmetic per item; three passes of loads are in flight
        const int IW = R16 >> 2;                                  // staged word of tile column 0
        const int half = lane >> 4, li = lane & 15;
        const bool tile_inside = vec_in && tx0 + GAUSS_TW <= w;   // warp-uniform
        const int gxi = tx0 + 8 * li;
        const bool tiny = h <= r;                                 // image shorter than the halo
        const uint8_t *colp = fin + 3 * (size_t)gxi;              // this lane's column in every row
        const unsigned pitch32 = (unsigned)in_pitch;
        uint8_t *sdst = s8 + 4 * (IW + 2 * li);
        for (int tr0 = 2 * warp + half; tr0 < R; tr0 += 3 * 2 * (GAUSS_THREADS / 32)) {
            uint2 q[3][3];
            if (tile_inside) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int tr = tr0 + k * 2 * (GAUSS_THREADS / 32);
                    if (tr < R) {
                        const int gy = gauss_reflect_row(ty0 + tr - r, h, tiny);
                        const uint2 *p = reinterpret_cast<const uint2 *>(colp + (size_t)((unsigned)gy * pitch32));
                        q[k][0] = __ldg(p); q[k][1] = __ldg(p + 1); q[k][2] = __ldg(p + 2);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int tr = tr0 + k * 2 * (GAUSS_THREADS / 32);
                if (tr >= R) break;
                unsigned lo, hi;
                if (tile_inside) {
                    lo = va_luma_x4(q[k][0].x, q[k][0].y, q[k][1].x, mode);
                    hi = va_luma_x4(q[k][1].y, q[k][2].x, q[k][2].y, mode);
                } else {
                    const uint8_t *rp = fin + (size_t)gauss_reflect_fast(ty0 + tr - r, h) * in_pitch;
                    lo = hi = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int gx = gxi + i;
                        if (gx < need_hi) {
                            const unsigned v = va_luma_px(rp + 3 * (size_t)gauss_reflect_fast(gx, w), mode) << (8 * (i & 3));
                            if (i < 4) lo |= v; else hi |= v;
                        }
                    }
                }
                *reinterpret_cast<uint2 *>(sdst + tr * SW) = make_uint2(lo, hi);
            }
        }
        // halo: HS items of 8 pixels on either side of every staged row
        const int HS = (r + 7) >> 3;
        for (int it = tid; it < R * 2 * HS; it += GAUSS_THREADS) {
            const int tr = it / (2 * HS), k = it - tr * 2 * HS;
            const bool right = k >= HS;
            const int kk = right ? k - HS : k;
            const int gx0 = right ? tx0 + GAUSS_TW + 8 * kk : tx0 - 8 * (kk + 1);
            const int word = right ? IW + 32 + 2 * kk : IW - 2 * (kk + 1);
            const uint8_t *rp = fin + (size_t)gauss_reflect_fast(ty0 + tr - r, h) * in_pitch;
            unsigned lo, hi;
            if (vec_in && gx0 >= 0 && gx0 + 8 <= w) {
                const uint2 *p = reinterpret_cast<const uint2 *>(rp + 3 * (size_t)gx0);
                const uint2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
                lo = va_luma_x4(q0.x, q0.y, q1.x, mode);
                hi = va_luma_x4(q1.y, q2.x, q2.y, mode);
            } else {
                lo = hi = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int gx = gx0 + i;
                    if (gx >= need_lo && gx < need_hi) {
                        const unsigned v = va_luma_px(rp + 3 * (size_t)gauss_reflect_fast(gx, w), mode) << (8 * (i & 3));
                        if (i < 4) lo |= v; else hi |= v;
                    }
                }
            }
            *reinterpret_cast<uint2 *>(s8 + tr * SW + 4 * word) = make_uint2(lo, hi);
        }
    }
    __syncthreads();

    // ---- row pass: warp = row pair q, lane = 4-pixel group
    if (!(dbg & 2))
    for (int q = warp; q < (R >> 1); q += GAUSS_THREADS / 32) {
        const unsigned *r0 = reinterpret_cast<const unsigned *>(s8 + (2 * q) * SW) + WOFS + lane;
        const unsigned *r1 = reinterpret_cast<const unsigned *>(s8 + (2 * q + 1) * SW) + WOFS + lane;
        unsigned a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < NW; j++) {
            const unsigned x0 = r0[j], x1 = r1[j];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned c = g.cw[i][j];
                a0[i] = __dp4a(x0, c, a0[i]);
                a1[i] = __dp4a(x1, c, a1[i]);
            }
        }
        *reinterpret_cast<uint4 *>(hp + q * GAUSS_TW + 4 * lane) =
            make_uint4(__byte_perm(a0[0], a1[0], 0x5410), __byte_perm(a0[1], a1[1], 0x5410),
                       __byte_perm(a0[2], a1[2], 0x5410), __byte_perm(a0[3], a1[3], 0x5410));
    }
    __syncthreads();

    // ---- column pass: warp = output row pair yp, lane = 4-column group
    const int x = tx0 + 4 * lane;
    if (x >= w || (dbg & 4)) return;
    const bool vec_store = out_words && x + 4 <= w;
    uint8_t *op = out + (size_t)blockIdx.z * out_fstride + (size_t)(ty0 + 2 * warp) * out_pitch + x;
    const size_t ostep = (size_t)(GAUSS_THREADS / 32) * 2 * out_pitch;
    for (int yp = warp; yp < (TH >> 1); yp += GAUSS_THREADS / 32, op += ostep) {
        const int y = ty0 + 2 * yp;
        if (y >= h) break;
        unsigned acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[i][k] = 32768u;
        const uint4 *col = reinterpret_cast<const uint4 *>(hp + yp * GAUSS_TW + 4 * lane);
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const uint4 v = col[j * (GAUSS_TW >> 2)];
            const unsigned c0 = g.cp[0][j], c1 = g.cp[1][j];
            acc[0][0] = __dp2a_lo(v.x, c0, acc[0][0]); acc[1][0] = __dp2a_lo(v.x, c1, acc[1][0]);
            acc[0][1] = __dp2a_lo(v.y, c0, acc[0][1]); acc[1][1] = __dp2a_lo(v.y, c1, acc[1][1]);
            acc[0][2] = __dp2a_lo(v.z, c0, acc[0][2]); acc[1][2] = __dp2a_lo(v.z, c1, acc[1][2]);
            acc[0][3] = __dp2a_lo(v.w, c0, acc[0][3]); acc[1][3] = __dp2a_lo(v.w, c1, acc[1][3]);
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (y + i >= h) break;
            // byte 2 of every accumulator (sum + 32768 < 2^24) is the rounded result
            const unsigned res = __byte_perm(__byte_perm(acc[i][0], acc[i][1], 0x0062),
                                             __byte_perm(acc[i][2], acc[i][3], 0x0062), 0x5410);
            uint8_t *o = op + (i ? out_pitch : 0);
            if (vec_store) {
                *reinterpret_cast<unsigned *>(o) = res;
            } else {
                for (int k = 0; k < 4 && x + k < w; k++) o[k] = (uint8_t)(res >> (8 * k));
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// streaming kernel (radius <= 9: sigma <= 3; wider windows cost too many registers and the tile kernel
// wins again; 16-byte aligned frames, w % 16 == 0): a warp owns a strip of 128
// columns and walks down a segment of rows.  Lane t owns columns 4t .. 4t+3 for the whole walk and
// keeps the last RT + 1 row pairs of the row pass in registers, so the 16-bit
// intermediate image never exists in memory and no row of the segment is staged twice.
// Warps are independent (each stages its own 128 columns plus a 16-column halo on either side), so
// there is no CTA barrier.  A step is RT + 1 row pairs = one turn of the register window, so every
// slot index is static.  Per pair of step s:
//   stage     one 16-pixel item of step s + 1: raw bytes (cp.async'ed into a private ring slot a few
//             pairs ago) -> luma words in the row buffer of step s + 1; the slot is refilled at once
//   rows      dp4a on 2 rows x 4 pixels from the row buffer of step s -> window slot
//   cols      dp2a over the window -> 2 output rows x 4 pixels, one 4-byte store each
// and one __syncwarp per step.
// ---------------------------------------------------------------------------------
#define GS_MAX_THREADS 256
#define GS_MAX_R 16
#define GS_DEPTH 3                       // raw ring slots per thread
#define GS_HW 4                          // halo words (16 pixels) on either side of a staged row

// luma of four interleaved RGB pixels with the channel pick expressed as two PRMT selectors
__device__ __forceinline__ unsigned gs_luma_x4(unsigned w0, unsigned w1, unsigned w2, bool mean, unsigned sel_a, unsigned sel_b) {
    if (mean) return va_mean3_x4(w0, w1, w2);
    return __byte_perm(__byte_perm(w0, w1, sel_a), w2, sel_b);
}

template <int RT, bool FUSE_LUMA>
__global__ void __launch_bounds__(GS_MAX_THREADS, RT <= 8 ? 3 : 2)
gauss_stream_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                    uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                    int w, int h, int mode, int SH, const __grid_constant__ GaussFast g) {
    // grid = (groups of strips in x, segments in y, frames); every warp works on its own
    constexpr int r = RT, NP = RT + 1;
    constexpr int O = (4 - (r & 3)) & 3;
    constexpr int NW = (O + 2 * r + 4 + 3) >> 2;             // staged words feeding one 4-pixel group
    constexpr int WOFS = GS_HW - ((r + 3) >> 2);             // first staged word needed by group 0
    constexpr int NQ = FUSE_LUMA ? 3 : 1;                    // 16-byte chunks per 16-pixel item
    constexpr int ROWS = 2 * NP;                             // image rows per step (one turn of the window)
    constexpr int IPT = (ROWS + 3) / 4;                      // interior items per lane and step
    constexpr int HPT = (2 * ROWS + 31) / 32;                // halo items per lane and step
    constexpr int NI = IPT + HPT;
    constexpr int D = GS_DEPTH;                              // raw ring slots per lane
    constexpr int RW = 32 + 2 * GS_HW;                       // words per staged row
    static_assert(NI <= NP, "one item per pair");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x;
    VA_DYN_SMEM(uint4, smem);
    uint4 *rb = smem + tid;                                  // [D][NQ][NT] private raw bytes
    unsigned *lbuf = reinterpret_cast<unsigned *>(smem + D * NQ * NT) + warp * (2 * ROWS * RW);   // [warp][2][ROWS][RW] luma words
    const int x0 = (blockIdx.x * (NT >> 5) + warp) * 128;
    if (x0 >= w) return;
    const int y0 = blockIdx.y * SH;
    const int rows = min(SH, h - y0);
    const int npairs = ((rows + 1) >> 1) + r;
    const int nsteps = (npairs + NP - 1) / NP;
    const uint8_t *fin = in + (size_t)blockIdx.z * in_fstride;
    const unsigned pitch32 = (unsigned)in_pitch;
    const bool mean = mode < 0;
    const unsigned sel_a = mode == 0 ? 0x0630u : mode == 1 ? 0x0741u : 0x0052u;
    const unsigned sel_b = mode == 0 ? 0x5210u : mode == 1 ? 0x6210u : 0x7410u;

    // staging duty of this lane: NI items of 16 pixels per step -- IPT of the strip's own 128 columns (a
    // quarter warp covers one row; rows frow, frow + 4, ...) and HPT of the 16-column halos on either side.
    // Items travel through a private ring of D raw slots: item n is converted while items n + 1 ..
    // n + D - 1 are in flight, and its slot is refilled with item n + D at once.
    // Columns outside the image (w % 16 == 0, so an item is inside or outside as a whole): the item
    // next to the left / right border is the mirrored neighbour item (BORDER_REFLECT_101, radius < 16),
    // anything further out is never read and staged as 0.
    enum { IN = 0, MIRROR_L = 1, MIRROR_R = 2, ZERO = 3 };
    const int frow = lane >> 3;
    const int fcol = lane & 7;
    const int fgx = x0 + 16 * fcol;
    const int fkind = fgx < w ? IN : fgx == w ? MIRROR_R : ZERO;
    const uint8_t *fcolp = fin + (FUSE_LUMA ? 3 : 1) * (size_t)(fkind == MIRROR_R ? w - 16 : fgx);
    const bool hside = lane & 1;
    const int hgx = hside ? x0 + 128 : x0 - 16;
    const int hkind = hgx < 0 ? MIRROR_L : hgx < w ? IN : hgx == w ? MIRROR_R : ZERO;
    const uint8_t *hcolp = fin + (FUSE_LUMA ? 3 : 1) * (size_t)(hkind == MIRROR_L ? 0 : hkind == MIRROR_R ? w - 16 : hgx);

    auto issue = [&](int st, int i, int slot) {             // i static
        if (st < nsteps) {
            const bool halo = i >= IPT;
            const int row = halo ? (lane >> 1) + 16 * (i - IPT) : frow + 4 * i;
            if ((halo ? hkind : fkind) != ZERO && row < ROWS) {
                const unsigned gy = (unsigned)gauss_reflect_row(y0 - r + ROWS * st + row, h, false);
                const uint8_t *p = (halo ? hcolp : fcolp) + (size_t)(gy * pitch32);
#pragma unroll
                for (int k = 0; k < NQ; k++) va_cp_async16(rb + (slot * NQ + k) * NT, p + 16 * k);
            }
        }
        va_cp_async_commit();
    };
    auto convert = [&](int st, int i, int slot) {           // i static
        if (st >= nsteps) return;
        const bool halo = i >= IPT;
        const int row = halo ? (lane >> 1) + 16 * (i - IPT) : frow + 4 * i;
        if (row >= ROWS) return;
        const int kind = halo ? hkind : fkind;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (kind != ZERO) {
            const uint4 *q = rb + slot * NQ * NT;
            if (FUSE_LUMA) {
                const uint4 a = q[0], b = q[NT], c = q[2 * NT];
                val = make_uint4(gs_luma_x4(a.x, a.y, a.z, mean, sel_a, sel_b), gs_luma_x4(a.w, b.x, b.y, mean, sel_a, sel_b),
                                 gs_luma_x4(b.z, b.w, c.x, mean, sel_a, sel_b), gs_luma_x4(c.y, c.z, c.w, mean, sel_a, sel_b));
            } else {
                val = q[0];
            }
            if (kind == MIRROR_L)        // columns -16 .. -1 <- columns 16 .. 1 (column 16 is never read)
                val = make_uint4(__byte_perm(val.w, 0, 0x1234), __byte_perm(val.z, val.w, 0x1234),
                                 __byte_perm(val.y, val.z, 0x1234), __byte_perm(val.x, val.y, 0x1234));
            else if (kind == MIRROR_R)   // columns w .. w + 15 <- columns w - 2 .. w - 17 (the last one is never read)
                val = make_uint4(__byte_perm(val.z, val.w, 0x3456), __byte_perm(val.y, val.z, 0x3456),
                                 __byte_perm(val.x, val.y, 0x3456), __byte_perm(0, val.x, 0x3456));
        }
        unsigned *lb = lbuf + (st & 1) * ROWS * RW + row * RW;
        *reinterpret_cast<uint4 *>(lb + (halo ? (hside ? GS_HW + 32 : 0) : GS_HW + 4 * fcol)) = val;
    };

    const int x = x0 + 4 * lane;
    const bool active = x < w;
    uint8_t *orow = out + (size_t)blockIdx.z * out_fstride + (size_t)y0 * out_pitch + x;
    unsigned win[NP][4];
#pragma unroll
    for (int u = 0; u < NP; u++)
#pragma unroll
        for (int k = 0; k < 4; k++) win[u][k] = 0;

    // prologue: fill the ring, then stage all of step 0
#pragma unroll
    for (int n = 0; n < D; n++) issue(n / NI, n % NI, n);
#pragma unroll
    for (int i = 0; i < NI; i++) {
        va_cp_async_wait_group<D - 1>();
        convert(0, i, i % D);
        issue((i + D) / NI, (i + D) % NI, i % D);
    }
    __syncwarp();
    int bslot = NI % D;                                     // ring slot of item 0 of the step being staged
    for (int s = 0; s < nsteps; s++) {
        const unsigned *lb = lbuf + (s & 1) * ROWS * RW + WOFS + lane;
#pragma unroll
        for (int gg = 0; gg < NP; gg++) {              // pair j lives in window slot gg
            // ---- stage one item of step s + 1 (into the other row buffer), refill its ring slot
            if (gg < NI) {
                int slot = bslot + gg % D;
                slot = slot >= D ? slot - D : slot;
                va_cp_async_wait_group<D - 1>();
                convert(s + 1, gg, slot);
                issue(s + 1 + (gg + D) / NI, (gg + D) % NI, slot);
            }
            const int j = s * NP + gg;
            if (!active || j >= npairs) continue;
            // ---- row pass of pair j
            const unsigned *r0 = lb + 2 * gg * RW;
            const unsigned *r1 = r0 + RW;
            unsigned a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
#pragma unroll
            for (int jw = 0; jw < NW; jw++) {
                const unsigned xa = r0[jw], xb = r1[jw];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (4 * jw + 3 < O + i || 4 * jw > O + i + 2 * r) continue;   // no tap of pixel i in this word
                    const unsigned c = g.cw[i][jw];
                    a0[i] = __dp4a(xa, c, a0[i]);
                    a1[i] = __dp4a(xb, c, a1[i]);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) win[gg][k] = __byte_perm(a0[k], a1[k], 0x5410);
            if (s == 0 && gg < r) continue;
            // ---- column pass: output rows 2 (j - r) + {0, 1} of the segment from pairs j - r .. j,
            // i.e. window slots gg + 1, gg + 2, ... (mod NP)
            unsigned acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int k = 0; k < 4; k++) acc[i][k] = 32768u;
#pragma unroll
            for (int t = 0; t < NP; t++) {
                const unsigned c0 = g.cp[0][t], c1 = g.cp[1][t];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned v = win[(gg + 1 + t) % NP][k];
                    acc[0][k] = __dp2a_lo(v, c0, acc[0][k]);
                    acc[1][k] = __dp2a_lo(v, c1, acc[1][k]);
                }
            }
            // byte 2 of every accumulator (sum + 32768 < 2^24) is the rounded result
            *reinterpret_cast<unsigned *>(orow) =
                __byte_perm(__byte_perm(acc[0][0], acc[0][1], 0x0062), __byte_perm(acc[0][2], acc[0][3], 0x0062), 0x5410);
            if (2 * (j - r) + 1 < rows)
                *reinterpret_cast<unsigned *>(orow + out_pitch) =
                    __byte_perm(__byte_perm(acc[1][0], acc[1][1], 0x0062), __byte_perm(acc[1][2], acc[1][3], 0x0062), 0x5410);
            orow += 2 * out_pitch;
        }
        bslot = (bslot + NI) % D;
        __syncwarp();
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
    VA_REQUIRE(ctx, (unsigned long long)h * in_pitch < (1ull << 32), "%s: frame larger than 4 GiB", name);
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
        // tensor-core kernel (va_gauss_mma.cu) wherever its shape constraints hold; VA_GAUSS_MMA=0 keeps the dot-product kernels
        {
            const char *env = getenv("VA_GAUSS_MMA");
            // radius 9 (16 + 2r = 34: a third 16-row group for two rows) is the one case where the streaming dot-product
            // kernel is still ahead (0.159 against 0.183 ms per 64 frames of 1080p)
            if ((!env || atoi(env) != 0) && !(r == 9 && (!env || atoi(env) != 2))) {
                int rc = va_gauss_mma_launch(ctx, stream, name, fuse, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
                                             w, h, batch, mode, taps, ksize);
                if (rc != VA_ERR_UNSUPPORTED) return rc;
                if (fuse && r > 8 && w % 16 == 0) {
                    // the fused tensor-core kernel covers radius <= 8; beyond that the monochrome frames go through a
                    // stream-ordered temporary (N bytes per frame more traffic on a kernel that is far from HBM-bound)
                    uint8_t *tmp = nullptr;
                    const size_t tp = (size_t)w, tf = tp * (size_t)h;
                    if (cudaMallocAsync((void **)&tmp, tf * (size_t)batch, (cudaStream_t)stream) == cudaSuccess) {
                        rc = va_luma_u8(ctx, stream, in, in_pitch, in_fstride, tmp, tp, tf, w, h, batch, mode);
                        if (rc == VA_OK)
                            rc = va_gauss_mma_launch(ctx, stream, name, false, tmp, tp, tf, out, out_pitch, out_fstride,
                                                     w, h, batch, mode, taps, ksize);
                        cudaFreeAsync(tmp, (cudaStream_t)stream);
                        if (rc != VA_ERR_UNSUPPORTED) return rc;
                    } else {
                        cudaGetLastError();
                    }
                }
            }
        }
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
        // streaming kernel for small radii on aligned frames
        {
            const char *env = getenv("VA_GAUSS_STREAM");
            const int stream_mode = env ? atoi(env) : 1;          // 0: tile kernel only (tuning / A-B checks)
            const bool in16 = va_aligned(in, 16) && in_pitch % 16 == 0 && in_fstride % 16 == 0;
            const bool out4 = va_aligned(out, 4) && out_pitch % 4 == 0 && out_fstride % 4 == 0;
            if (stream_mode && r <= 9 && h >= r + 2 && w % 16 == 0 && in16 && out4 && batch <= 65535) {
                // warps per CTA: least idle warps at the right image edge, then the size closest to 3 warps
                // (warps are independent; small CTAs pack the SMs better)
                int NT = 96;
                long long best = 1ll << 60;
                for (int nt = 32; nt <= GS_MAX_THREADS; nt += 32) {
                    const long long cover = (long long)va_div_up(w, 4 * nt) * nt;
                    const int dist = nt > 96 ? nt - 96 : 96 - nt, bdist = NT > 96 ? NT - 96 : 96 - NT;
                    if (cover < best || (cover == best && dist < bdist)) { best = cover; NT = nt; }
                }
                if (getenv("VA_GS_NT")) NT = atoi(getenv("VA_GS_NT"));
                const int strips = va_div_up(w, 4 * NT);
                // segments in y: about 12 CTAs per SM (measured optimum at 1080p: fewer leave SMs idle at the
                // end, more re-stage too many halo rows -- each segment stages 2 r extra rows)
                int segs = va_div_up((long long)ctx->sm_count * 12, (long long)strips * batch);
                const int max_segs = h / (8 * r) > 0 ? h / (8 * r) : 1;
                if (segs > max_segs) segs = max_segs;
                if (getenv("VA_GS_SEGS")) segs = atoi(getenv("VA_GS_SEGS"));
                if (segs < 1) segs = 1;
                const int SH = 2 * va_div_up(h, 2 * segs);
                const int segs_y = va_div_up(h, SH);
                if (segs_y <= 65535) {
                    const int NQ = fuse ? 3 : 1;
                    const int ROWS = 2 * (r + 1);
                    const size_t smem = (size_t)GS_DEPTH * NQ * NT * 16 + (size_t)(NT / 32) * 2 * ROWS * (32 + 2 * GS_HW) * 4;
                    const dim3 grid(strips, segs_y, batch);
#define GS_GO(RT, FUSE)                                                                                       \
                    do {                                                                                      \
                        auto kfn = gauss_stream_kernel<RT, FUSE>;                                             \
                        if (smem > 48 * 1024)                                                                 \
                            VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                        VA_LAUNCH(ctx, kfn, grid, NT, smem, stream, in, in_pitch, in_fstride, out, out_pitch, \
                                  out_fstride, w, h, mode, SH, g);                                            \
                    } while (0)
#define GS_CASE(RT) case RT: if (fuse) GS_GO(RT, true); else GS_GO(RT, false); break;
                    switch (r) {
                        GS_CASE(1) GS_CASE(2) GS_CASE(3) GS_CASE(4) GS_CASE(5) GS_CASE(6) GS_CASE(7) GS_CASE(8)
                        GS_CASE(9)
                    }
#undef GS_CASE
#undef GS_GO
                    return VA_OK;
                }
            }
        }
        // tile height: minimise staged rows + idle warp rounds + rows wasted below the image
        const int R16 = (r + 15) & ~15;
        const int SW = GAUSS_TW + 2 * R16 + 16;
        int TH = 0;
        {
            const char *env = getenv("VA_GAUSS_TH");
            const int forced = env ? atoi(env) : 0;
            double best = 1e30;
            for (int th = 32; th <= 192; th += 2) {
                const int R = th + 2 * r;
                const size_t sm = (size_t)R * SW + (size_t)(R / 2) * GAUSS_TW * 4;
                if (sm > (r <= 16 ? 44 : 112) * 1024) break;   // keep >= 5 CTAs per SM for small radii
                const int ty = va_div_up(h, th);
                const double cost = ty * (11.0 * 16 * ((R + 15) / 16) + 10.0 * 16 * ((th + 15) / 16));   // staged rows + output rows, 16 per round
                if (cost < best) { best = cost; TH = th; }
            }
            if (forced >= 8 && forced % 2 == 0) TH = forced;
            if (TH == 0) TH = 32;
        }
        const int R = TH + 2 * r;
        const size_t smem = (size_t)R * SW + (size_t)(R / 2) * GAUSS_TW * 4;
        VA_REQUIRE(ctx, smem <= 220 * 1024, "%s: tile does not fit in shared memory", name);
        const int tiles_x = va_div_up(w, GAUSS_TW), tiles_y = va_div_up(h, TH);
        VA_REQUIRE(ctx, tiles_y <= 65535 && batch <= 65535, "%s: too many tiles for one launch", name);
        const dim3 grid(tiles_x, tiles_y, batch);
        const int dbg = getenv("VA_GAUSS_DBG") ? atoi(getenv("VA_GAUSS_DBG")) : 0;   // timing ablation only
        const int vec_in = fuse ? (va_aligned(in, 8) && in_pitch % 8 == 0 && in_fstride % 8 == 0)
                                : (va_aligned(in, 16) && in_pitch % 16 == 0 && in_fstride % 16 == 0);
#define GAUSS_GO(RT, FUSE)                                                                                       \
        do {                                                                                                     \
            auto kfn = gauss_fast_kernel<RT, FUSE>;                                                              \
            if (smem > 48 * 1024)                                                                                \
                VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            VA_LAUNCH(ctx, kfn, grid, GAUSS_THREADS, smem, stream, in, in_pitch, in_fstride, out, out_pitch,     \
                      out_fstride, w, h, mode, TH, vec_in, dbg, g);                                                   \
        } while (0)
#define GAUSS_CASE(RT) case RT: if (fuse) GAUSS_GO(RT, true); else GAUSS_GO(RT, false); break;
        switch (r) {
            GAUSS_CASE(3) GAUSS_CASE(4) GAUSS_CASE(5) GAUSS_CASE(6) GAUSS_CASE(8) GAUSS_CASE(9) GAUSS_CASE(12) GAUSS_CASE(15)
            default: if (fuse) GAUSS_GO(0, true); else GAUSS_GO(0, false); break;
        }
#undef GAUSS_CASE
#undef GAUSS_GO
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
