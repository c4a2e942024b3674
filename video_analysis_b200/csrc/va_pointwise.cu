// va_pointwise.cu -- K1 luma / channel pick (+crop addressing), strided copy,
// K2b 2x2 area resize, K6 apply-mask, threshold / pack / unpack of bit masks and
// the seeded synthetic-video generator.  All HBM-bound streaming kernels.
#include "va_device.cuh"

// =================================================================================
// K1: interleaved RGB u8 -> luma u8.  Rows are staged through shared memory with
// 16-byte cp.async so that global reads are full-sector regardless of where a
// crop rectangle starts; each thread then turns 12 staged bytes into 4 pixels
// with dp4a (sum of 3) + mulhi (exact /3).
// Algorithmic bytes: 3N read + N written = 4N.
// =================================================================================
#define LUMA_THREADS 256

__global__ void __launch_bounds__(LUMA_THREADS)
luma_rows_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                 uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                 int w, int h, int mode, int rows_per_tile, int tiles_per_frame, int n_tiles,
                 int srow /* smem bytes per staged row, multiple of 16 */) {
    VA_DYN_SMEM(uint8_t, smem);
    const int tid = threadIdx.x;
    const int row_bytes = 3 * w;
    const int groups = (w + 3) >> 2;            // 4-pixel groups per row
    const bool out_words = (((uintptr_t)out | out_pitch | out_fstride) & 3) == 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_frame;
        const int y0 = (tile - b * tiles_per_frame) * rows_per_tile;
        const int nrows = min(rows_per_tile, h - y0);
        const uint8_t *fin = in + (size_t)b * in_fstride;

        // ---- stage rows: body with aligned 16-byte async copies, ragged ends bytewise
        for (int r = 0; r < nrows; r++) {
            const uint8_t *src = fin + (size_t)(y0 + r) * in_pitch;
            const int a = (int)((uintptr_t)src & 15);       // misalignment of this row
            uint8_t *dst = smem + (size_t)r * srow;         // dst[a + i] <- src[i]
            const int first = a ? 16 - a : 0;               // bytes before the first aligned chunk
            const int body = row_bytes > first ? (row_bytes - first) >> 4 : 0;
            for (int c = tid; c < body; c += LUMA_THREADS)
                va_cp_async16(dst + a + first + 16 * c, src + first + 16 * c);
            const int tail0 = first + 16 * body;
            for (int i = tid; i < min(first, row_bytes); i += LUMA_THREADS) dst[a + i] = src[i];
            for (int i = tail0 + tid; i < row_bytes; i += LUMA_THREADS) dst[a + i] = src[i];
        }
        va_cp_async_wait_all();
        __syncthreads();

        // ---- 4 pixels per item
        for (int r = tid >> 5; r < nrows; r += LUMA_THREADS / 32)
        for (int g = tid & 31; g < groups; g += 32) {
            const uint8_t *src = fin + (size_t)(y0 + r) * in_pitch;
            const int a = (int)((uintptr_t)src & 15);
            const int off = r * srow + a + 12 * g;
            const unsigned *sw = reinterpret_cast<const unsigned *>(smem + (off & ~3));
            const unsigned sel = 0x3210u + 0x1111u * (off & 3);
            const unsigned s0 = sw[0], s1 = sw[1], s2 = sw[2], s3 = sw[3];
            const unsigned w0 = __byte_perm(s0, s1, sel);
            const unsigned w1 = __byte_perm(s1, s2, sel);
            const unsigned w2 = __byte_perm(s2, s3, sel);
            const unsigned res = va_luma_x4(w0, w1, w2, mode);
            uint8_t *orow = out + (size_t)b * out_fstride + (size_t)(y0 + r) * out_pitch;
            const int x = 4 * g;
            if (out_words && x + 4 <= w) {
                *reinterpret_cast<unsigned *>(orow + x) = res;
            } else {
                for (int i = 0; i < 4 && x + i < w; i++) orow[x + i] = (uint8_t)(res >> (8 * i));
            }
        }
        __syncthreads();
    }
}

// Fast path (everything 16-byte aligned, width a multiple of 16): one warp converts 512
// pixels per step.  The 1536 input bytes are fetched with three perfectly coalesced
// 16-byte loads per lane, bounced through a per-warp shared-memory slab so that every lane
// ends up with the 48 contiguous bytes of its own 16 pixels (conflict-free: 48-byte stride),
// and leave as one 16-byte store per lane.  No block-wide barrier.
#define LUMA_WARPS 8
__global__ void __launch_bounds__(LUMA_WARPS * 32)
luma_fast_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                 uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                 int w, int h, long long rows, int chunks_per_row, int mode) {
    __shared__ uint4 stage[LUMA_WARPS][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *st = stage[warp];
    const unsigned total = (unsigned)(rows * chunks_per_row);
    for (unsigned item = blockIdx.x * LUMA_WARPS + warp; item < total;
         item += gridDim.x * LUMA_WARPS) {
        const unsigned row = item / (unsigned)chunks_per_row;
        const int px0 = (int)(item - row * chunks_per_row) * 512;
        const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
        const int npx = min(512, w - px0);                 // multiple of 16
        const int n16 = (npx * 3) >> 4;
        const uint8_t *src = in + (size_t)b * in_fstride + (size_t)y * in_pitch + 3 * (size_t)px0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int idx = lane + 32 * k;
            if (idx < n16) st[idx] = va_ld_stream16(src + 16 * idx);
        }
        __syncwarp();
        if (16 * lane < npx) {
            const uint4 q0 = st[3 * lane], q1 = st[3 * lane + 1], q2 = st[3 * lane + 2];
            uint4 r;
            r.x = va_luma_x4(q0.x, q0.y, q0.z, mode);
            r.y = va_luma_x4(q0.w, q1.x, q1.y, mode);
            r.z = va_luma_x4(q1.z, q1.w, q2.x, mode);
            r.w = va_luma_x4(q2.y, q2.z, q2.w, mode);
            va_st_stream16(out + (size_t)b * out_fstride + (size_t)y * out_pitch + px0 + 16 * lane, r);
        }
        __syncwarp();
    }
}

extern "C" int va_luma_u8(va_ctx *ctx, va_stream stream,
                          const uint8_t *in, size_t in_pitch, size_t in_fstride,
                          uint8_t *out, size_t out_pitch, size_t out_fstride,
                          int w, int h, int batch, int mode) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out, "va_luma_u8: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "va_luma_u8: bad size %dx%dx%d", w, h, batch);
    VA_REQUIRE(ctx, mode >= -1 && mode <= 2, "va_luma_u8: unsupported conversion method to monochrome: %d", mode);
    VA_REQUIRE(ctx, in_pitch >= (size_t)3 * w && out_pitch >= (size_t)w, "va_luma_u8: pitch smaller than a row");
    if (w % 16 == 0 && va_aligned(in, 16) && va_aligned(out, 16) && in_pitch % 16 == 0 && in_fstride % 16 == 0 &&
        out_pitch % 16 == 0 && out_fstride % 16 == 0) {
        long long rows = (long long)h * batch;
        int fw = w, fh = h;
        size_t fin_pitch = in_pitch, fout_pitch = out_pitch;
        // a dense batch is one long row: no ragged chunk at every row end
        if (in_pitch == (size_t)3 * w && out_pitch == (size_t)w && in_fstride == in_pitch * h &&
            out_fstride == out_pitch * h && (long long)w * h * batch < (1ll << 30)) {
            fw = w * h * batch; fh = 1; rows = 1;
            fin_pitch = (size_t)3 * fw; fout_pitch = (size_t)fw;
        }
        const int chunks = va_div_up(fw, 512);
        const long long items = rows * chunks;
        const int grid = va_grid(ctx, (items + LUMA_WARPS - 1) / LUMA_WARPS, 8);
        auto kfast = luma_fast_kernel;
        VA_LAUNCH(ctx, kfast, grid, LUMA_WARPS * 32, 0, stream, in, fin_pitch, in_fstride, out, fout_pitch, out_fstride,
                  fw, fh, rows, chunks, mode);
        return VA_OK;
    }
    const int srow = ((3 * w + 15 + 16 + 15) / 16) * 16;   // misalignment slack + one word of over-read
    int rpt = 12288 / srow;
    if (rpt < 1) rpt = 1;
    if (rpt > 16) rpt = 16;
    if (rpt > h) rpt = h;
    const int tiles_per_frame = va_div_up(h, rpt);
    const int n_tiles = tiles_per_frame * batch;
    const size_t smem = (size_t)rpt * srow;
    VA_REQUIRE(ctx, smem <= 200 * 1024, "va_luma_u8: row of %d pixels does not fit in shared memory", w);
    auto kfn = luma_rows_kernel;
    if (smem > 48 * 1024)
        VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = va_grid(ctx, n_tiles, 8);
    VA_LAUNCH(ctx, kfn, grid, LUMA_THREADS, smem, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
              w, h, mode, rpt, tiles_per_frame, n_tiles, srow);
    return VA_OK;
}

// =================================================================================
// strided 2-D copy (crop of frames that keep their channels)
// =================================================================================
__global__ void __launch_bounds__(256)
copy2d_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
              uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
              int row_bytes, int h, int batch, int vec) {
    const long long rows = (long long)h * batch;
    if (vec) {
        const int chunks = row_bytes >> 4;
        const unsigned total = (unsigned)(rows * chunks);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += gridDim.x * blockDim.x) {
            const unsigned row = i / (unsigned)chunks;
            const int c = (int)(i - row * chunks);
            const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
            const uint4 v = va_ld_stream16(in + (size_t)b * in_fstride + (size_t)y * in_pitch + 16 * c);
            va_st_stream16(out + (size_t)b * out_fstride + (size_t)y * out_pitch + 16 * c, v);
        }
    } else {
        const unsigned total = (unsigned)(rows * row_bytes);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += gridDim.x * blockDim.x) {
            const unsigned row = i / (unsigned)row_bytes;
            const int c = (int)(i - row * row_bytes);
            const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
            out[(size_t)b * out_fstride + (size_t)y * out_pitch + c] =
                in[(size_t)b * in_fstride + (size_t)y * in_pitch + c];
        }
    }
}

extern "C" int va_copy2d_u8(va_ctx *ctx, va_stream stream,
                            const uint8_t *in, size_t in_pitch, size_t in_fstride,
                            uint8_t *out, size_t out_pitch, size_t out_fstride,
                            int row_bytes, int h, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out, "va_copy2d_u8: null pointer");
    VA_REQUIRE(ctx, row_bytes > 0 && h > 0 && batch > 0, "va_copy2d_u8: bad size");
    const int vec = (row_bytes % 16 == 0) && va_aligned(in, 16) && va_aligned(out, 16) &&
                    in_pitch % 16 == 0 && out_pitch % 16 == 0 && in_fstride % 16 == 0 && out_fstride % 16 == 0;
    const long long items = (long long)h * batch * (vec ? row_bytes / 16 : row_bytes);
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = copy2d_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
              row_bytes, h, batch, vec);
    return VA_OK;
}

// =================================================================================
// K2b: exact 2x2 area average (cv2.resize INTER_AREA by 1/2): (a+b+c+d+2) >> 2
// Algorithmic bytes: N read + N/4 written.
// =================================================================================
__global__ void __launch_bounds__(256)
resize_half_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                   uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                   int ow, int oh, int channels, int batch, int vec) {
    // one item = 4 output bytes of one output row (channels == 1, vec) or 1 output byte
    if (vec) {
        const int groups = ow >> 2;
        const unsigned total = (unsigned)((long long)groups * oh * batch);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += gridDim.x * blockDim.x) {
            const unsigned row = i / (unsigned)groups;
            const int g = (int)(i - row * groups);
            const int b = (int)(row / (unsigned)oh), y = (int)(row - (unsigned)b * (unsigned)oh);
            const uint8_t *p = in + (size_t)b * in_fstride + (size_t)(2 * y) * in_pitch + 8 * g;
            const uint2 r0 = *reinterpret_cast<const uint2 *>(p);
            const uint2 r1 = *reinterpret_cast<const uint2 *>(p + in_pitch);
            // vertical sums of byte pairs as 16-bit lanes (max 510), then horizontal pairs
            const unsigned e0 = (r0.x & 0x00FF00FFu) + (r1.x & 0x00FF00FFu);            // bytes 0,2
            const unsigned o0 = ((r0.x >> 8) & 0x00FF00FFu) + ((r1.x >> 8) & 0x00FF00FFu);  // bytes 1,3
            const unsigned e1 = (r0.y & 0x00FF00FFu) + (r1.y & 0x00FF00FFu);
            const unsigned o1 = ((r0.y >> 8) & 0x00FF00FFu) + ((r1.y >> 8) & 0x00FF00FFu);
            const unsigned s0 = ((e0 + o0 + 0x00020002u) >> 2) & 0x00FF00FFu;   // outputs 0,1 in 16-bit lanes
            const unsigned s1 = ((e1 + o1 + 0x00020002u) >> 2) & 0x00FF00FFu;   // outputs 2,3
            const unsigned res = __byte_perm(s0, s1, 0x6420);
            *reinterpret_cast<unsigned *>(out + (size_t)b * out_fstride + (size_t)y * out_pitch + 4 * g) = res;
        }
    } else {
        const int orow = ow * channels;
        const unsigned total = (unsigned)((long long)orow * oh * batch);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += gridDim.x * blockDim.x) {
            const unsigned row = i / (unsigned)orow;
            const int xb = (int)(i - row * orow);
            const int x = xb / channels, c = xb - x * channels;
            const int b = (int)(row / (unsigned)oh), y = (int)(row - (unsigned)b * (unsigned)oh);
            const uint8_t *p = in + (size_t)b * in_fstride + (size_t)(2 * y) * in_pitch + (size_t)(2 * x) * channels + c;
            const unsigned s = (unsigned)p[0] + p[channels] + p[in_pitch] + p[in_pitch + channels] + 2u;
            out[(size_t)b * out_fstride + (size_t)y * out_pitch + xb] = (uint8_t)(s >> 2);
        }
    }
}

extern "C" int va_resize_half_u8(va_ctx *ctx, va_stream stream,
                                 const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                 uint8_t *out, size_t out_pitch, size_t out_fstride,
                                 int w, int h, int channels, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out, "va_resize_half_u8: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && (channels == 1 || channels == 3), "va_resize_half_u8: bad size");
    VA_REQUIRE(ctx, w % 2 == 0 && h % 2 == 0, "va_resize_half_u8: %dx%d is not even", w, h);
    const int ow = w / 2, oh = h / 2;
    const int vec = channels == 1 && ow % 4 == 0 && va_aligned(in, 8) && in_pitch % 8 == 0 && in_fstride % 8 == 0 &&
                    va_aligned(out, 4) && out_pitch % 4 == 0 && out_fstride % 4 == 0;
    const long long items = vec ? (long long)(ow / 4) * oh * batch : (long long)ow * channels * oh * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = resize_half_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, out, out_pitch, out_fstride,
              ow, oh, channels, batch, vec);
    return VA_OK;
}

// =================================================================================
// K6: apply-mask, out = mask ? in : 0.  2N + N/8.. bytes; mask is u8 (H, W)
// =================================================================================
__global__ void __launch_bounds__(256)
apply_mask_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                  const uint8_t *__restrict__ mask, size_t mask_pitch, size_t mask_fstride,
                  uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                  int w, int h, int channels, int batch, int vec) {
    if (vec) {   // channels == 1: 16 pixels per item
        const int chunks = w >> 4;
        const unsigned total = (unsigned)((long long)chunks * h * batch);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += gridDim.x * blockDim.x) {
            const unsigned row = i / (unsigned)chunks;
            const int c = (int)(i - row * chunks);
            const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
            uint4 v = va_ld_stream16(in + (size_t)b * in_fstride + (size_t)y * in_pitch + 16 * c);
            const uint4 m = *reinterpret_cast<const uint4 *>(mask + (size_t)b * mask_fstride + (size_t)y * mask_pitch + 16 * c);
            // per-byte mask != 0 -> 0xFF: ((m | (0x80 - m)... ) use carry-free trick on 7 low bits
            #define NZ(x) ((((((x) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | (x)) & 0x80808080u) >> 7) * 0xFFu
            v.x &= NZ(m.x); v.y &= NZ(m.y); v.z &= NZ(m.z); v.w &= NZ(m.w);
            #undef NZ
            va_st_stream16(out + (size_t)b * out_fstride + (size_t)y * out_pitch + 16 * c, v);
        }
    } else {
        const int rowb = w * channels;
        const unsigned total = (unsigned)((long long)rowb * h * batch);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += gridDim.x * blockDim.x) {
            const unsigned row = i / (unsigned)rowb;
            const int xb = (int)(i - row * rowb);
            const int x = xb / channels;
            const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
            const uint8_t m = mask[(size_t)b * mask_fstride + (size_t)y * mask_pitch + x];
            const uint8_t v = in[(size_t)b * in_fstride + (size_t)y * in_pitch + xb];
            out[(size_t)b * out_fstride + (size_t)y * out_pitch + xb] = m ? v : (uint8_t)0;
        }
    }
}

extern "C" int va_apply_mask_u8(va_ctx *ctx, va_stream stream,
                                const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                const uint8_t *mask, size_t mask_pitch, size_t mask_fstride,
                                uint8_t *out, size_t out_pitch, size_t out_fstride,
                                int w, int h, int channels, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && mask, "va_apply_mask_u8: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && (channels == 1 || channels == 3), "va_apply_mask_u8: bad size");
    const int vec = channels == 1 && w % 16 == 0 && va_aligned(in, 16) && va_aligned(out, 16) && va_aligned(mask, 16) &&
                    in_pitch % 16 == 0 && out_pitch % 16 == 0 && mask_pitch % 16 == 0 &&
                    in_fstride % 16 == 0 && out_fstride % 16 == 0 && mask_fstride % 16 == 0;
    const long long items = vec ? (long long)(w / 16) * h * batch : (long long)w * channels * h * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = apply_mask_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, mask, mask_pitch, mask_fstride,
              out, out_pitch, out_fstride, w, h, channels, batch, vec);
    return VA_OK;
}

// =================================================================================
// u8 -> packed bits (threshold or nonzero) and back.  One warp handles 512 pixels
// of a row per step: lane loads 16 pixels, builds 16 mask bits, lane pairs merge
// into one 32-bit word.
// =================================================================================
__device__ __forceinline__ uint4 va_load16_u8(const uint8_t *row, int x, int w, bool vec) {
    if (vec && x + 16 <= w) return va_ld_stream16(row + x);
    unsigned v[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++)
        if (x + i < w) v[i >> 2] |= (unsigned)row[x + i] << (8 * (i & 3));
    return make_uint4(v[0], v[1], v[2], v[3]);
}
// bit i of the result = byte i of `x` > thr (unsigned compare), 4 bytes
__device__ __forceinline__ unsigned va_gt_mask4(unsigned x, int thr) {
    unsigned m = 0;
    m |= ((int)(x & 0xFF) > thr) ? 1u : 0u;
    m |= ((int)((x >> 8) & 0xFF) > thr) ? 2u : 0u;
    m |= ((int)((x >> 16) & 0xFF) > thr) ? 4u : 0u;
    m |= ((int)(x >> 24) > thr) ? 8u : 0u;
    return m;
}

__global__ void __launch_bounds__(256)
threshold_bits_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                      uint32_t *__restrict__ mask, size_t mask_pitch_w, size_t mask_fstride_w,
                      int w, int h, int batch, int thr, int vec) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int chunks = (w + 511) >> 9;                       // 512-pixel chunks per row
    const unsigned total = (unsigned)((long long)chunks * h * batch);
    for (unsigned i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < total;
         i += gridDim.x * warps_per_block) {
        const unsigned row = i / (unsigned)chunks;
        const int c = (int)(i - row * chunks);
        const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
        const uint8_t *rp = in + (size_t)b * in_fstride + (size_t)y * in_pitch;
        const int x = c * 512 + lane * 16;
        const uint4 v = va_load16_u8(rp, x, w, vec != 0);
        unsigned m = va_gt_mask4(v.x, thr) | (va_gt_mask4(v.y, thr) << 4) | (va_gt_mask4(v.z, thr) << 8) |
                     (va_gt_mask4(v.w, thr) << 12);
        if (x + 16 > w) m &= (x < w) ? ((1u << (w - x)) - 1u) : 0u;
        const unsigned hi = __shfl_down_sync(0xffffffffu, m, 1);
        if (!(lane & 1) && x < w)
            mask[(size_t)b * mask_fstride_w + (size_t)y * mask_pitch_w + (x >> 5)] = m | (hi << 16);
    }
}

static int launch_threshold(va_ctx *ctx, va_stream stream, const char *name,
                            const uint8_t *in, size_t in_pitch, size_t in_fstride,
                            uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                            int w, int h, int batch, int thr) {
    VA_REQUIRE(ctx, in && mask, "%s: null pointer", name);
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "%s: bad size", name);
    VA_REQUIRE(ctx, mask_pitch_w >= (size_t)((w + 31) / 32), "%s: mask pitch smaller than a row", name);
    const int vec = va_aligned(in, 16) && in_pitch % 16 == 0 && in_fstride % 16 == 0;
    const long long warps = (long long)((w + 511) / 512) * h * batch;
    const int grid = va_grid(ctx, (warps + 7) / 8, 8);
    auto kfn = threshold_bits_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, mask, mask_pitch_w, mask_fstride_w,
              w, h, batch, thr, vec);
    return VA_OK;
}

extern "C" int va_threshold_bits(va_ctx *ctx, va_stream stream,
                                 const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                 uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                                 int w, int h, int batch, int thr) {
    VA_CHECK_CTX(ctx);
    return launch_threshold(ctx, stream, "va_threshold_bits", in, in_pitch, in_fstride, mask, mask_pitch_w,
                            mask_fstride_w, w, h, batch, thr);
}

extern "C" int va_pack_bits_u8(va_ctx *ctx, va_stream stream,
                               const uint8_t *in, size_t in_pitch, size_t in_fstride,
                               uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                               int w, int h, int batch) {
    VA_CHECK_CTX(ctx);
    return launch_threshold(ctx, stream, "va_pack_bits_u8", in, in_pitch, in_fstride, mask, mask_pitch_w,
                            mask_fstride_w, w, h, batch, 0);
}

__global__ void __launch_bounds__(256)
unpack_bits_kernel(const uint32_t *__restrict__ mask, size_t mask_pitch_w, size_t mask_fstride_w,
                   uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                   int w, int h, int batch, int vec) {
    // one item = 16 pixels
    const int chunks = (w + 15) >> 4;
    const unsigned total = (unsigned)((long long)chunks * h * batch);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += gridDim.x * blockDim.x) {
        const unsigned row = i / (unsigned)chunks;
        const int c = (int)(i - row * chunks);
        const int b = (int)(row / (unsigned)h), y = (int)(row - (unsigned)b * (unsigned)h);
        const unsigned word = mask[(size_t)b * mask_fstride_w + (size_t)y * mask_pitch_w + (c >> 1)];
        const unsigned m = (word >> ((c & 1) * 16)) & 0xFFFFu;
        unsigned v[4];
        for (int q = 0; q < 4; q++) {
            const unsigned nib = (m >> (4 * q)) & 0xFu;
            // spread 4 bits to 4 bytes of 0x00 / 0xFF
            v[q] = (((nib * 0x00204081u) & 0x01010101u) * 0xFFu);
        }
        uint8_t *orow = out + (size_t)b * out_fstride + (size_t)y * out_pitch;
        const int x = 16 * c;
        if (vec && x + 16 <= w) {
            va_st_stream16(orow + x, make_uint4(v[0], v[1], v[2], v[3]));
        } else {
            for (int k = 0; k < 16 && x + k < w; k++) orow[x + k] = (uint8_t)(v[k >> 2] >> (8 * (k & 3)));
        }
    }
}

extern "C" int va_unpack_bits_u8(va_ctx *ctx, va_stream stream,
                                 const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                                 uint8_t *out, size_t out_pitch, size_t out_fstride,
                                 int w, int h, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, mask && out, "va_unpack_bits_u8: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "va_unpack_bits_u8: bad size");
    const int vec = va_aligned(out, 16) && out_pitch % 16 == 0 && out_fstride % 16 == 0;
    const long long items = (long long)((w + 15) / 16) * h * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = unpack_bits_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, mask, mask_pitch_w, mask_fstride_w, out, out_pitch, out_fstride,
              w, h, batch, vec);
    return VA_OK;
}

// =================================================================================
// seeded synthetic video: integer hash, identical to oracle/synth.py
// =================================================================================
#define VA_MAX_BLOBS 32
struct SynthBlobs { int n; int v[VA_MAX_BLOBS][5]; };

__device__ __forceinline__ unsigned va_mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
static unsigned host_mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ int va_posmod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

__global__ void __launch_bounds__(256)
synth_rgb_kernel(uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                 int w, int h, int t0, int batch, unsigned seed, unsigned kbase, SynthBlobs blobs) {
    // one item = one pixel (3 bytes)
    const unsigned total = (unsigned)((long long)w * h * batch);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += gridDim.x * blockDim.x) {
        const int b = (int)(i / (unsigned)(w * h));
        const int p = (int)(i - (unsigned)b * (unsigned)(w * h));
        const int y = p / w, x = p - y * w;
        const int t = t0 + b;
        const unsigned kt = va_mix32(seed ^ ((unsigned)(t + 1) * 0x9E3779B9u));
        int add = 0;
        for (int k = 0; k < blobs.n; k++) {
            const int cx = va_posmod(blobs.v[k][0] * 16 + blobs.v[k][2] * t, 16 * w) >> 4;
            const int cy = va_posmod(blobs.v[k][1] * 16 + blobs.v[k][3] * t, 16 * h) >> 4;
            const int dx = x - cx, dy = y - cy, r = blobs.v[k][4];
            if (dx * dx + dy * dy <= r * r) add = 90;
        }
        uint8_t *o = out + (size_t)b * out_fstride + (size_t)y * out_pitch + 3 * (size_t)x;
        for (int c = 0; c < 3; c++) {
            const unsigned idx = (unsigned)p * 3u + (unsigned)c;
            const int base = 60 + (int)__umulhi(va_mix32(idx ^ kbase), 60u);
            const int noise = (int)__umulhi(va_mix32(idx + kt), 17u) - 8;
            int v = base + noise + add;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            o[c] = (uint8_t)v;
        }
    }
}

extern "C" int va_synth_rgb(va_ctx *ctx, va_stream stream,
                            uint8_t *out, size_t out_pitch, size_t out_fstride,
                            int w, int h, int t0, int batch, uint32_t seed,
                            const int32_t *blobs, int n_blobs) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, out, "va_synth_rgb: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "va_synth_rgb: bad size");
    VA_REQUIRE(ctx, n_blobs >= 0 && n_blobs <= VA_MAX_BLOBS && (n_blobs == 0 || blobs), "va_synth_rgb: 0..%d blobs", VA_MAX_BLOBS);
    SynthBlobs sb;
    memset(&sb, 0, sizeof(sb));
    sb.n = n_blobs;
    for (int k = 0; k < n_blobs; k++)
        for (int j = 0; j < 5; j++) sb.v[k][j] = blobs[5 * k + j];
    const unsigned kbase = host_mix32(seed * 0x9E3779B9u + 0x01234567u);
    const long long items = (long long)w * h * batch;
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = synth_rgb_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, out, out_pitch, out_fstride, w, h, t0, batch, seed, kbase, sb);
    return VA_OK;
}
