// va_gauss_mma.cu -- K2 on the integer tensor cores: bit-exact cv2.GaussianBlur on uint8 (video/filters.py:392)
// as two products with banded Toeplitz matrices, staged by TMA.
//
// OpenCV's 8-bit Gaussian is  out = (sum_ky K[ky] * Hrow[y + ky - r][x] + 32768) >> 16,  Hrow = sum_kx K[kx] * src[.][x + kx - r]
// (taps K are 8-bit, sum 256).  Both passes are products with a banded Toeplitz matrix of the
// taps, and  mma.sync.m16n8k32.s32.u8.u8.s32  is exact, so the result is bit-identical however the sums are grouped:
//
//   row pass     C1[x_out (16)][row (8)]   = T1[x_out][k]  *  src[row][window start + k]          k = 32 KS source columns
//   column pass  C2[y_out (16)][x (8)]     = T2[y_out][k]  *  byte plane of Hrow[16-row groups][x] k = 16 G input rows,
//                twice: low and high byte of the 16-bit Hrow,  out = byte 2 of ((C2_hi << 8) + C2_lo + 32768)
//
// The constant Toeplitz fragments live in registers (built once per warp from the tap table).  Because the matrix is
// ours, the k index of a fragment is a free permutation: B fragments are whatever 8 consecutive source bytes a lane
// reads with one LDS.64 (row pass), and whatever four rows a lane already holds in its row-pass accumulators (column
// pass), so the 16-bit intermediate goes from accumulator to operand through a per-warp shared-memory ring without any
// transposition.
//
// A warp owns a strip of 128 columns and walks down a segment of rows, 8 rows per half-step; warps share nothing.
// Staging: one lane issues a TMA box load (cp.async.bulk.tensor) of the next 8 rows of strip + halo, two half-steps
// ahead, and the warp waits on the stage's mbarrier.  Rows above / below the image are fetched as single-row boxes at
// their BORDER_REFLECT_101 row; columns outside the image arrive as TMA zero fill and are overwritten with the mirrored
// luma bytes in shared memory (strips at the image edge only).  The fused variant converts RGB -> luma from the raw
// stage into the row buffer; the plain variant lets TMA write the row buffer itself.
//
// Algorithmic HBM bytes: 2N (4N fused).  Not handled here (gauss_launch falls back to the dot-product kernels):
// rows that are not 16-byte aligned, widths that are not a multiple of 16, images smaller than the radius, radius > 56.
#include <cstring>

#include "va_mma.cuh"

#define GM_OUT_PITCH(TILES) (16 * (TILES) + 16)   // bytes per row of the output tile: rows land in different banks
#define GM_MAX_G 8
#define GM_MAX_STAGES 4
#define GMC_MAX_STAGES 16       // CTA-per-strip kernel: a stage is only 8 rows x (128 + 2 HL) bytes

struct GaussMma {
    int r;                  // radius, taps K[0 .. 2r]
    int SH;                 // rows per segment (multiple of 16)
    int strips, segs;       // task grid: strips x segments x frames
    int stages;             // TMA stages of 8 rows in flight per warp (<= GM_MAX_STAGES)
    int mode;               // fused: -1 mean, 0..2 channel
    unsigned char taps[128];
};

int va_tmap_encode(va_tmap *map, const void *base, size_t row_bytes, size_t rows, size_t frames, size_t pitch, size_t fstride,
                   unsigned box_w32, unsigned box_rows) {
#ifdef VA_EMU
    map->base = reinterpret_cast<const uint8_t *>(base);
    map->w32 = (unsigned)(row_bytes / 4);
    map->rows = (unsigned)rows;
    map->frames = (unsigned)frames;
    map->pitch = pitch;
    map->fstride = fstride;
    map->box_w32 = box_w32;
    map->box_rows = box_rows;
    return 0;
#else
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
        encode = (encode_fn)fn;
    }
    // a single frame still needs a non-zero, 16-byte-multiple stride for the (unused) third dimension
    const cuuint64_t dims[3] = {(cuuint64_t)(row_bytes / 4), (cuuint64_t)rows, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(frames > 1 ? fstride : pitch * rows)};
    const cuuint32_t box[3] = {box_w32, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? 0 : -1;
#endif
}

__device__ __forceinline__ unsigned gm_tap(const GaussMma &g, int idx) {
    return (idx >= 0 && idx <= 2 * g.r) ? g.taps[idx] : 0u;
}

// bytes of the shared memory one warp needs
// staged halo per side: a multiple of 16 pixels, so that the box of the fused variant (3 bytes per pixel) starts on a 16-byte
// boundary of the row -- a box starting 24 bytes before the strip (halo 8) faults with "illegal instruction" on the B200
__host__ __device__ inline int gm_hl(int G, bool fuse) { (void)fuse; const int KS = (G + 1) / 2; return KS == 1 ? 16 : KS <= 3 ? 48 : 80; }
// pitch of the luma rows the fused variant writes: >= LW and 32 (mod 64), which keeps the LDS.64 of the B fragments conflict-free
__host__ __device__ inline int gm_lp(int LW) { return ((LW - 32 + 63) / 64) * 64 + 32; }
__host__ __device__ inline size_t gm_warp_smem(bool fuse, int G, int tiles, int stages) {
    const size_t LW = (size_t)(16 * tiles + 2 * gm_hl(G, fuse));
    const size_t SP = ((fuse ? 3 * LW : LW) + 127) & ~(size_t)127;   // stage row pitch of a half-step fetched row by row
    size_t b = 0;
    b += (size_t)stages * 8 * SP;                   // TMA stages
    size_t ot = 16 * (size_t)GM_OUT_PITCH(tiles);   // output tile; the fused variant keeps its 8 luma rows in the same place
    if (fuse && 8 * (size_t)gm_lp((int)LW) > ot) ot = 8 * (size_t)gm_lp((int)LW);
    b += (ot + 15) & ~(size_t)15;
    if (G > 2) b += (size_t)G * tiles * 32 * 16;    // ring of row-pass results (16 rows x strip x 16 bit per group)
    return (b + 127) & ~(size_t)127;
}

// geometry that follows from the number of 16-row groups G of a column-pass block (16 + 2r <= 16 G)
template <bool FUSE, int G, int TILES> struct GmGeom {
    static constexpr int KS = (G + 1) / 2;                        // row pass: 16 + 2r <= 32 KS
    static constexpr int OFF = 16 * KS - 8;                       // row pass: output x = window start + OFF (>= r, multiple of 8)
    static constexpr int HL = KS == 1 ? 16 : KS <= 3 ? 48 : 80;   // staged halo (gm_hl): >= OFF, and 128 + 2 HL = 32 (mod 64) so that
                                                                  // the LDS.64 of the B fragments are bank-conflict free
    static constexpr int STRIP = 16 * TILES;                      // columns a warp owns
    static constexpr int LW = STRIP + 2 * HL;                     // staged pixels per row
    static constexpr int LP = ((LW - 32 + 63) / 64) * 64 + 32;    // fused: pitch of the converted luma rows (gm_lp)
};

// fetch 8 rows row by row at their BORDER_REFLECT_101 position (top / bottom of the image only; kept out of line)
static __device__ __noinline__ void gm_issue_rows(unsigned char *dst, const va_tmap *map1, uint64_t *bar, int x32, int ya, int h,
                                                  int frame, int sp) {
#pragma unroll 1
    for (int i = 0; i < 8; i++) va_tma_load_3d(dst + (size_t)i * sp, map1, bar, x32, va_reflect101(ya + i, h), frame);
}

template <bool FUSE, int G, int TILES, int MINB>
__global__ void __launch_bounds__(128, MINB)
gauss_mma_kernel(const __grid_constant__ va_tmap map8, const __grid_constant__ va_tmap map1,
                 uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                 int w, int h, int batch, const __grid_constant__ GaussMma gp) {
    typedef GmGeom<FUSE, G, TILES> Geo;
    constexpr int KS = Geo::KS, HL = Geo::HL, off = Geo::OFF, LW = Geo::LW, LP = Geo::LP;
    static_assert(!FUSE || (G == 2 && TILES == 8), "the fused conversion is laid out for 128-column strips, halo 16, window offset 8");
    constexpr int GM_STRIP = 16 * TILES, GM_TILES = TILES, OPITCH = GM_OUT_PITCH(TILES);
    constexpr int KF = G / 2;                       // column pass: full k32 steps (two groups each), then one k16 step if G is odd
    constexpr int C = FUSE ? 3 : 1;
    constexpr int BOXB = C * LW;                    // bytes per staged row
    constexpr int SP = (BOXB + 127) & ~127;         // pitch of the rows of a half-step that is fetched row by row (TMA
                                                    // destinations are 128-byte aligned); a box of 8 rows lands densely
    const int NS = gp.stages;                       // TMA stages of 8 rows in flight per warp
    constexpr bool REG_RING = (G == 2);             // two groups: the previous group stays in registers, no ring
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r = gp.r, SH = gp.SH;

    // ---- which strip / segment / frame ----------------------------------------------------------------
    const long long task = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long n_tasks = (long long)gp.strips * gp.segs * batch;
    if (task >= n_tasks) return;                    // warps are independent: no block-wide barrier below
    const int strip = (int)(task % gp.strips);
    const int seg = (int)((task / gp.strips) % gp.segs);
    const int frame = (int)(task / ((long long)gp.strips * gp.segs));
    const int x0 = strip * GM_STRIP, y0 = seg * SH;
    const int rows = min(SH, h - y0);
    const int nblocks = (rows + 15) >> 4;

    // ---- shared memory of this warp ---------------------------------------------------------------------
    VA_DYN_SMEM(unsigned char, smem_raw);
#ifdef VA_EMU
    unsigned char *sm0 = smem_raw;
#else
    unsigned char *sm0 = smem_raw + ((128u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u);
#endif
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm0) + NS * warp;                  // [warps][NS]
    unsigned char *wsm = sm0 + 256 + (size_t)warp * gm_warp_smem(FUSE, G, TILES, NS);   // 4 warps x NS x 8 bytes of barriers fit in 256
    unsigned char *stage0 = wsm;
    unsigned char *lbuf = wsm + NS * 8 * SP;                                         // FUSE: luma rows; shares its place with the output tile
    unsigned char *otile = lbuf;
    uint4 *ring = reinterpret_cast<uint4 *>(lbuf + (((FUSE && 8 * LP > 16 * OPITCH) ? 8 * LP : 16 * OPITCH) + 15 & ~15));   // G > 2 only

    // ---- constant Toeplitz fragments ------------------------------------------------------------------
    // row pass: fragment row m is output column pi(m) of the tile, pi(2u) = 4u, pi(2u+1) = 4u+1, pi(8+2u) = 4u+2, pi(9+2u) = 4u+3,
    // so that a lane's column-pass results are four neighbouring columns
    unsigned A1[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int m = g + ((q & 1) ? 8 : 0);
            const int xo = 4 * ((m & 7) >> 1) + (m & 1) + ((m & 8) ? 2 : 0);
            unsigned wd = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = 32 * ks + 8 * t + ((q & 2) ? 4 : 0) + j;            // source column relative to the window start
                wd |= gm_tap(gp, c - off - xo + r) << (8 * j);
            }
            A1[ks][q] = wd;
        }
    // column pass: physical k = 4t + j of a 16-row group is its local row rho(t, j) (what the row-pass fragments hold)
    unsigned A2[KF > 0 ? KF : 1][4], A2t[2] = {0, 0};
#pragma unroll
    for (int grp = 0; grp < G; grp++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int m = g + (q ? 8 : 0);
            unsigned wd = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int rho = j < 2 ? 2 * t + j : 8 + 2 * t + (j - 2);
                wd |= gm_tap(gp, 16 * grp + rho - m) << (8 * j);                  // input row 16 grp + rho, output row r + m
            }
            if (grp < 2 * KF) A2[grp >> 1][q + ((grp & 1) ? 2 : 0)] = wd;
            else A2t[q] = wd;
        }

    // ---- TMA producer (lane 0) ----------------------------------------------------------------------------
    if (lane == 0) {
        for (int i = 0; i < NS; i++) va_mbar_init(&bars[i], 1);
        va_mbar_fence_init();
    }
    __syncwarp();
    const int n_half = 2 * (nblocks + G - 1);
    const int x32 = ((x0 - HL) * C) >> 2;          // (x0 - HL) * C is a multiple of 16 (may be negative: arithmetic shift is exact)
    auto issue = [&](int hs, int st) {             // rows 8 hs .. 8 hs + 7 of the segment's padded row space into stage st
        if (lane != 0 || hs >= n_half) return;
        uint64_t *bar = &bars[st];
        unsigned char *dst = stage0 + (size_t)st * 8 * SP;
        const int ya = y0 - r + 8 * hs;
        va_mbar_expect_tx(bar, 8u * (unsigned)BOXB);
        if (ya >= 0 && ya + 7 < h) va_tma_load_3d(dst, &map8, bar, x32, ya, frame);
        else gm_issue_rows(dst, &map1, bar, x32, ya, h, frame, SP);
    };
    for (int i = 0; i < NS; i++) issue(i, i);

    const bool fix_l = x0 - HL < 0, fix_r = x0 + GM_STRIP + HL > w;
    const bool mean = gp.mode < 0;
    const unsigned sel_a = gp.mode == 0 ? 0x0630u : gp.mode == 1 ? 0x0741u : 0x0052u;
    const unsigned sel_b = gp.mode == 0 ? 0x5210u : gp.mode == 1 ? 0x6210u : 0x7410u;
    uint8_t *fout = out + (size_t)frame * out_fstride;
    unsigned hold[GM_TILES][2];
    uint4 prev[REG_RING ? GM_TILES : 1], cur[REG_RING ? GM_TILES : 1];
    int st = 0;                                    // stage of the next half-step
    unsigned phases = 0;                           // bit i: parity the next wait on stage i expects

    for (int S = 0; S < nblocks + G - 1; S++) {
        const int slot = S % G;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int hs = 2 * S + half;
            va_mbar_wait(&bars[st], (phases >> st) & 1u);
            phases ^= 1u << st;
            unsigned char *stg = stage0 + (size_t)st * 8 * SP;
            const int ya = y0 - r + 8 * hs;
            const int sp = (ya >= 0 && ya + 7 < h) ? BOXB : SP;      // row pitch of this stage (see issue())
            unsigned char *lrow = FUSE ? lbuf : stg;                // 8 rows of luma
            const int lp = FUSE ? LP : sp;
            if (FUSE) {
                // RGB -> luma of the 8 x 144 pixels the row pass reads (staged pixels 8 .. 151 of 160).  Two rounds of 16-pixel
                // items (three 16-byte loads, one 16-byte store; lane -> row lane / 8 (+ 4), item 1 + lane % 8: every address
                // static, conflict-free), then the two half items at the ends of the 8 rows as 32 pieces of 4 pixels: no lane
                // idles in any round (the loop over all 80 items of the stage ran 3 rounds of 32 lanes for 2.5 rounds of work)
                auto conv4 = [&](unsigned w0, unsigned w1, unsigned w2) {
                    return mean ? va_mean3_x4(w0, w1, w2) : __byte_perm(__byte_perm(w0, w1, sel_a), w2, sel_b);
                };
#pragma unroll
                for (int rd = 0; rd < 2; rd++) {
                    const int row = (lane >> 3) + 4 * rd, ci = 1 + (lane & 7);
                    const uint4 *q = reinterpret_cast<const uint4 *>(stg + row * sp + 48 * ci);
                    const uint4 a = q[0], b = q[1], c = q[2];
                    *reinterpret_cast<uint4 *>(lbuf + row * LP + 16 * ci) =
                        make_uint4(conv4(a.x, a.y, a.z), conv4(a.w, b.x, b.y), conv4(b.z, b.w, c.x), conv4(c.y, c.z, c.w));
                }
                {
                    const int row = lane >> 2, px = ((lane & 2) ? 136 : 8) + 4 * (lane & 3);     // 8, 12, 144, 148
                    const unsigned *q = reinterpret_cast<const unsigned *>(stg + row * sp + 3 * px);
                    *reinterpret_cast<unsigned *>(lbuf + row * LP + px) = conv4(q[0], q[1], q[2]);
                }
                __syncwarp();
                issue(hs + NS, st);                                 // the raw stage is free again
            }
            if (fix_l || fix_r) {
                // BORDER_REFLECT_101 in x: columns -k <- k and w - 1 + k <- w - 1 - k, k = 1 .. r (luma bytes, in place)
                for (int it = lane; it < 8 * r; it += 32) {
                    const int row = it / r, k = it - row * r + 1;
                    unsigned char *p = lrow + (size_t)row * lp + HL;
                    if (fix_l && x0 + k <= HL) p[-x0 - k] = p[-x0 + k];
                    if (fix_r) {
                        const int d = w - 1 - x0 + k;
                        if (d < GM_STRIP + HL) p[d] = p[w - 1 - x0 - k];
                    }
                }
                __syncwarp();
            }
            // ---- row pass: 8 rows x 128 columns
#pragma unroll
            for (int j = 0; j < GM_TILES; j++) {
                int c[4] = {0, 0, 0, 0};
                const unsigned char *bp = lrow + (size_t)g * lp + (HL - off) + 16 * j + 8 * t;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    const uint2 b = *reinterpret_cast<const uint2 *>(bp + 32 * ks);
                    va_imma_16832(c, A1[ks], b.x, b.y);
                }
                // (row 2t, row 2t + 1) of fragment row g and of fragment row g + 8: low bytes, then high bytes
                const unsigned pg = __byte_perm((unsigned)c[0], (unsigned)c[1], 0x5140);
                const unsigned pg8 = __byte_perm((unsigned)c[2], (unsigned)c[3], 0x5140);
                if (half == 0) {
                    hold[j][0] = pg;
                    hold[j][1] = pg8;
                } else {
                    // rows {2t, 2t+1, 8+2t, 9+2t} of this 16-row group: one operand word per byte plane
                    const uint4 v = make_uint4(__byte_perm(hold[j][0], pg, 0x5410), __byte_perm(hold[j][0], pg, 0x7632),
                                               __byte_perm(hold[j][1], pg8, 0x5410), __byte_perm(hold[j][1], pg8, 0x7632));
                    if constexpr (REG_RING) cur[j] = v;
                    else ring[(slot * GM_TILES + j) * 32 + lane] = v;
                }
            }
            if (!FUSE) {
                __syncwarp();
                issue(hs + NS, st);                                 // TMA wrote the luma rows itself: the stage is free now
            }
            st = st + 1 == NS ? 0 : st + 1;
        }
        if (S >= G - 1) {
            // ---- column pass: output rows y0 + 16 b .. + 15 from groups b .. b + G - 1
            const int b = S - (G - 1);
            __syncwarp();                                           // FUSE: every lane is done with the luma rows the tile overwrites
#pragma unroll
            for (int j = 0; j < GM_TILES; j++) {
                uint4 v[G];
                if constexpr (REG_RING) {
                    v[0] = prev[j];
                    v[G - 1] = cur[j];
                } else {
#pragma unroll
                    for (int i = 0; i < G; i++) {
                        int sl = b + i;
                        sl -= (sl / G) * G;
                        v[i] = ring[(sl * GM_TILES + j) * 32 + lane];
                    }
                }
                unsigned res[2][2];                                 // [fragment row g / g + 8][nh]: two neighbouring columns each
#pragma unroll
                for (int nh = 0; nh < 2; nh++) {
                    // high byte plane first; (hi << 8) + 32768 is then the accumulator the low plane adds to (one IMAD per
                    // value does the shift and the rounding, and no accumulator has to be initialised with a constant)
                    int hi[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int ks = 0; ks < KF; ks++)
                        va_imma_16832(hi, A2[ks], nh ? v[2 * ks].w : v[2 * ks].y, nh ? v[2 * ks + 1].w : v[2 * ks + 1].y);
                    if (G & 1) va_imma_16816(hi, A2t[0], A2t[1], nh ? v[G - 1].w : v[G - 1].y);
                    int lo[4] = {hi[0] * 256 + 32768, hi[1] * 256 + 32768, hi[2] * 256 + 32768, hi[3] * 256 + 32768};
#pragma unroll
                    for (int ks = 0; ks < KF; ks++)
                        va_imma_16832(lo, A2[ks], nh ? v[2 * ks].z : v[2 * ks].x, nh ? v[2 * ks + 1].z : v[2 * ks + 1].x);
                    if (G & 1) va_imma_16816(lo, A2t[0], A2t[1], nh ? v[G - 1].z : v[G - 1].x);
                    // byte 2 of (hi << 8) + lo + 32768 (< 2^24) is the rounded result
                    const unsigned v0 = (unsigned)lo[0], v1 = (unsigned)lo[1], v2 = (unsigned)lo[2], v3 = (unsigned)lo[3];
                    res[0][nh] = __byte_perm(v0, v1, 0x0062);
                    res[1][nh] = __byte_perm(v2, v3, 0x0062);
                }
                // fragment columns 2t, 2t + 1 are tile columns 4t, 4t + 1 (nh = 0) and 4t + 2, 4t + 3 (nh = 1): one word per row
                unsigned char *op = otile + g * OPITCH + 16 * j + 4 * t;
                *reinterpret_cast<unsigned *>(op) = __byte_perm(res[0][0], res[0][1], 0x5410);
                *reinterpret_cast<unsigned *>(op + 8 * OPITCH) = __byte_perm(res[1][0], res[1][1], 0x5410);
            }
            __syncwarp();
            // ---- 16 rows x 128 bytes -> global, 16 bytes per lane and round
#pragma unroll
            for (int i = 0; i < (16 * TILES) / 32; i++) {
                const int idx = lane + 32 * i, row = idx / TILES, c16 = idx % TILES;
                const int yl = 16 * b + row, x = x0 + 16 * c16;
                if (yl < rows && x < w) {
                    const uint4 val = *reinterpret_cast<const uint4 *>(otile + row * OPITCH + 16 * c16);
                    va_st_stream16(fout + (size_t)(y0 + yl) * out_pitch + x, val);
                }
            }
            __syncwarp();
        }
        if constexpr (REG_RING) {
#pragma unroll
            for (int j = 0; j < GM_TILES; j++) prev[j] = cur[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Large radii (G > 2 groups, sigma > 2.7): one CTA of 8 warps per 128-column strip.
//
// The warp-per-strip kernel above keeps a ring of 16 G rows x strip x 16 bit per WARP (14 KB at sigma = 15 with 64-column
// strips) and stages strip + 160 halo columns per warp: 6 warps fit on an SM and each runs its phases one after the other.
// Here the warps of a CTA share the staged source rows (one TMA box of 8 rows x (128 + 2 HL) bytes per half-step,
// issued by thread 0) and every warp owns TPW of the eight 16-column tiles: its row pass reads the shared rows, its results
// go to a lane-private ring of G x 512 bytes per tile that only the same warp reads again in the column pass -- no data
// crosses warps except the source rows and the output tile.  Shared memory per tile drops threefold (sigma = 15: 5.4
// instead of 17 KB), the staged halo is shared by 128 instead of 64 columns, 16 warps are resident instead of 6, and three
// CTA barriers per 16 rows replace the per-warp TMA bookkeeping (sigma 15, 32 frames: 0.293 -> 0.197 ms):  [wait stage | mirror fix (edge strips) | row pass | barrier | refill stage] x 2,  column pass -> output
// tile, barrier, 16 rows x 128 bytes to global memory.  Same arithmetic, same fragments, same bytes as above.
// ---------------------------------------------------------------------------------------------------------
#define GMC_WARPS 8            // 16-column tiles of a strip (the name is from the one-tile-per-warp layout)
__host__ __device__ inline size_t gmc_smem(int G, int stages) {
    const size_t LW = (size_t)(16 * GMC_WARPS + 2 * gm_hl(G, false));
    const size_t SP = (LW + 127) & ~(size_t)127;
    return 256 + (size_t)stages * 8 * SP + 16 * (size_t)GM_OUT_PITCH(GMC_WARPS) + (size_t)GMC_WARPS * G * 512 + 128;
}

// TPW = tiles per warp: 8 / TPW warps per CTA (a warp's fixed cost per 16 rows is spread over TPW tiles)
template <int G, int MINB, int TPW>
__global__ void __launch_bounds__(32 * GMC_WARPS / TPW, MINB)
gauss_mma_cta_kernel(const __grid_constant__ va_tmap map8, const __grid_constant__ va_tmap map1,
                     uint8_t *__restrict__ out, size_t out_pitch, size_t out_fstride,
                     int w, int h, int batch, const __grid_constant__ GaussMma gp) {
    typedef GmGeom<false, G, GMC_WARPS> Geo;
    constexpr int KS = Geo::KS, HL = Geo::HL, off = Geo::OFF, LW = Geo::LW;
    constexpr int STRIP = 16 * GMC_WARPS, OPITCH = GM_OUT_PITCH(GMC_WARPS);
    constexpr int KF = G / 2;
    constexpr int SP = (LW + 127) & ~127;
    constexpr int NT = 32 * GMC_WARPS / TPW;
    const int NS = gp.stages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r = gp.r, SH = gp.SH;

    const long long task = blockIdx.x;
    const int strip = (int)(task % gp.strips);
    const int seg = (int)((task / gp.strips) % gp.segs);
    const int frame = (int)(task / ((long long)gp.strips * gp.segs));
    const int x0 = strip * STRIP, y0 = seg * SH;
    const int rows = min(SH, h - y0);
    const int nblocks = (rows + 15) >> 4;

    VA_DYN_SMEM(unsigned char, smem_raw);
#ifdef VA_EMU
    unsigned char *sm0 = smem_raw;
#else
    unsigned char *sm0 = smem_raw + ((128u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u);
#endif
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm0);                               // [NS]
    unsigned char *stage0 = sm0 + 256;
    unsigned char *otile = stage0 + (size_t)NS * 8 * SP;
    uint4 *ring = reinterpret_cast<uint4 *>(otile + 16 * OPITCH) + (size_t)warp * TPW * G * 32;   // [G][TPW][32] of this warp

    // ---- constant Toeplitz fragments (as in gauss_mma_kernel) ----------------------------------------------
    unsigned A1[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int m = g + ((q & 1) ? 8 : 0);
            const int xo = 4 * ((m & 7) >> 1) + (m & 1) + ((m & 8) ? 2 : 0);
            unsigned wd = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = 32 * ks + 8 * t + ((q & 2) ? 4 : 0) + j;
                wd |= gm_tap(gp, c - off - xo + r) << (8 * j);
            }
            A1[ks][q] = wd;
        }
    unsigned A2[KF > 0 ? KF : 1][4], A2t[2] = {0, 0};
#pragma unroll
    for (int grp = 0; grp < G; grp++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int m = g + (q ? 8 : 0);
            unsigned wd = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int rho = j < 2 ? 2 * t + j : 8 + 2 * t + (j - 2);
                wd |= gm_tap(gp, 16 * grp + rho - m) << (8 * j);
            }
            if (grp < 2 * KF) A2[grp >> 1][q + ((grp & 1) ? 2 : 0)] = wd;
            else A2t[q] = wd;
        }

    // ---- TMA producer (thread 0) ------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < NS; i++) va_mbar_init(&bars[i], 1);
        va_mbar_fence_init();
    }
    __syncthreads();
    const int n_half = 2 * (nblocks + G - 1);
    const int x32 = (x0 - HL) >> 2;                // x0 - HL is a multiple of 16 (may be negative)
    auto issue = [&](int hs, int st) {             // rows 8 hs .. 8 hs + 7 of the segment's padded row space into stage st
        if (tid != 0 || hs >= n_half) return;
        uint64_t *bar = &bars[st];
        unsigned char *dst = stage0 + (size_t)st * 8 * SP;
        const int ya = y0 - r + 8 * hs;
        va_mbar_expect_tx(bar, 8u * (unsigned)LW);
        if (ya >= 0 && ya + 7 < h) va_tma_load_3d(dst, &map8, bar, x32, ya, frame);
        else gm_issue_rows(dst, &map1, bar, x32, ya, h, frame, SP);
    };
    for (int i = 0; i < NS; i++) issue(i, i);
    __syncthreads();

    const bool fix_l = x0 - HL < 0, fix_r = x0 + STRIP + HL > w;
    uint8_t *fout = out + (size_t)frame * out_fstride;
    unsigned hold0[TPW], hold1[TPW];
    int st = 0;
    unsigned phases = 0;

    for (int S = 0; S < nblocks + G - 1; S++) {
        const int slot = S % G;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int hs = 2 * S + half;
            va_mbar_wait(&bars[st], (phases >> st) & 1u);
            phases ^= 1u << st;
            unsigned char *stg = stage0 + (size_t)st * 8 * SP;
            const int ya = y0 - r + 8 * hs;
            const int sp = (ya >= 0 && ya + 7 < h) ? LW : SP;      // row pitch of this stage (see issue())
            if (fix_l || fix_r) {
                // BORDER_REFLECT_101 in x, the whole CTA on the 8 staged rows (edge strips only)
                for (int it = tid; it < 8 * r; it += NT) {
                    const int row = it / r, k = it - row * r + 1;
                    unsigned char *p = stg + (size_t)row * sp + HL;
                    if (fix_l && x0 + k <= HL) p[-x0 - k] = p[-x0 + k];
                    if (fix_r) {
                        const int d = w - 1 - x0 + k;
                        if (d < STRIP + HL) p[d] = p[w - 1 - x0 - k];
                    }
                }
                __syncthreads();
            }
            // ---- row pass of this warp's tiles: 8 rows x 16 columns each
#pragma unroll
            for (int jj = 0; jj < TPW; jj++) {
                int c[4] = {0, 0, 0, 0};
                const unsigned char *bp = stg + (size_t)g * sp + (HL - off) + 16 * (warp * TPW + jj) + 8 * t;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    const uint2 bq = *reinterpret_cast<const uint2 *>(bp + 32 * ks);
                    va_imma_16832(c, A1[ks], bq.x, bq.y);
                }
                const unsigned pg = __byte_perm((unsigned)c[0], (unsigned)c[1], 0x5140);
                const unsigned pg8 = __byte_perm((unsigned)c[2], (unsigned)c[3], 0x5140);
                if (half == 0) {
                    hold0[jj] = pg;
                    hold1[jj] = pg8;
                } else {
                    ring[(slot * TPW + jj) * 32 + lane] =
                        make_uint4(__byte_perm(hold0[jj], pg, 0x5410), __byte_perm(hold0[jj], pg, 0x7632),
                                   __byte_perm(hold1[jj], pg8, 0x5410), __byte_perm(hold1[jj], pg8, 0x7632));
                }
            }
            __syncthreads();                                        // every warp is done with the stage
            issue(hs + NS, st);
            st = st + 1 == NS ? 0 : st + 1;
        }
        if (S >= G - 1) {
            // ---- column pass of this warp's tiles: output rows y0 + 16 b .. + 15 from groups b .. b + G - 1
            const int b = S - (G - 1);
#pragma unroll
            for (int jj = 0; jj < TPW; jj++) {
                uint4 v[G];
#pragma unroll
                for (int i = 0; i < G; i++) {
                    int sl = b + i;
                    sl -= (sl / G) * G;
                    v[i] = ring[(sl * TPW + jj) * 32 + lane];
                }
                unsigned res[2][2];
#pragma unroll
                for (int nh = 0; nh < 2; nh++) {
                    // (splitting these into independent accumulator chains -- even / odd k steps, high / low plane -- was measured
                    // slower, 0.279 vs 0.270 ms per 32 frames at sigma 15: the kernel waits at its barriers, not for the tensor pipe)
                    int hi[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int ks = 0; ks < KF; ks++)
                        va_imma_16832(hi, A2[ks], nh ? v[2 * ks].w : v[2 * ks].y, nh ? v[2 * ks + 1].w : v[2 * ks + 1].y);
                    if (G & 1) va_imma_16816(hi, A2t[0], A2t[1], nh ? v[G - 1].w : v[G - 1].y);
                    int lo[4] = {hi[0] * 256 + 32768, hi[1] * 256 + 32768, hi[2] * 256 + 32768, hi[3] * 256 + 32768};
#pragma unroll
                    for (int ks = 0; ks < KF; ks++)
                        va_imma_16832(lo, A2[ks], nh ? v[2 * ks].z : v[2 * ks].x, nh ? v[2 * ks + 1].z : v[2 * ks + 1].x);
                    if (G & 1) va_imma_16816(lo, A2t[0], A2t[1], nh ? v[G - 1].z : v[G - 1].x);
                    res[0][nh] = __byte_perm((unsigned)lo[0], (unsigned)lo[1], 0x0062);
                    res[1][nh] = __byte_perm((unsigned)lo[2], (unsigned)lo[3], 0x0062);
                }
                unsigned char *op = otile + g * OPITCH + 16 * (warp * TPW + jj) + 4 * t;
                *reinterpret_cast<unsigned *>(op) = __byte_perm(res[0][0], res[0][1], 0x5410);
                *reinterpret_cast<unsigned *>(op + 8 * OPITCH) = __byte_perm(res[1][0], res[1][1], 0x5410);
            }
            __syncthreads();
            // ---- 16 rows x 128 bytes -> global: 16-byte stores (the next write to the tile comes two barriers later)
            for (int idx = tid; idx < 16 * GMC_WARPS; idx += NT) {
                const int row = idx >> 3, c16 = idx & 7;
                const int yl = 16 * b + row, x = x0 + 16 * c16;
                if (yl < rows && x < w) {
                    const uint4 val = *reinterpret_cast<const uint4 *>(otile + row * OPITCH + 16 * c16);
                    va_st_stream16(fout + (size_t)(y0 + yl) * out_pitch + x, val);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// launch; returns VA_ERR_UNSUPPORTED when the shape is outside this kernel's range (the caller falls back)
// ---------------------------------------------------------------------------------------------------------
int va_gauss_mma_launch(va_ctx *ctx, va_stream stream, const char *name, bool fuse,
                        const uint8_t *in, size_t in_pitch, size_t in_fstride,
                        uint8_t *out, size_t out_pitch, size_t out_fstride,
                        int w, int h, int batch, int mode, const int *taps, int ksize) {
    const int r = ksize / 2;
    if (r < 1 || r > 56 || w % 16 != 0 || w < r + 2 || h < r + 1) return VA_ERR_UNSUPPORTED;
    if (!va_aligned(in, 16) || in_pitch % 16 || in_fstride % 16 || !va_aligned(out, 16) || out_pitch % 16 || out_fstride % 16)
        return VA_ERR_UNSUPPORTED;
    for (int i = 0; i < ksize; i++)
        if (taps[i] > 255) return VA_ERR_UNSUPPORTED;
    const int C = fuse ? 3 : 1;
    if ((unsigned long long)h * in_pitch >= (1ull << 40)) return VA_ERR_UNSUPPORTED;
    const int G = (16 + 2 * r + 15) / 16;
    // strip width: 128 columns; 64 where the ring of row-pass results (16 G rows x strip x 16 bit per warp) would leave too few
    // warps per SM.  The fused variant exists for G == 2 (the chain's sigma); larger radii convert first (gauss_launch).
    if (fuse && G > 2) return VA_ERR_UNSUPPORTED;
    // large radii: one CTA per 128-column strip (gauss_mma_cta_kernel), two tiles per warp.  Measured per 32 frames of 1080p
    // against the warp-per-strip kernel, sigma 5 / 9 / 15 (G = 3 / 5 / 7): 0.121 / 0.201 / 0.293 ms there; here 0.150 / 0.238 /
    // 0.274 with one tile per warp (a warp's fixed cost per 16 rows on a single tile: twice the instructions, and the CTA
    // waits at its barriers), 0.098 / 0.150 / 0.197 with two, 0.097 / 0.153 / 0.231 with four.  The number of TMA stages
    // (4 ... 16) does not matter.  VA_GM_CTA = 0 forces the warp-per-strip kernel.
    const bool use_cta = getenv("VA_GM_CTA") ? atoi(getenv("VA_GM_CTA")) != 0 : true;
    if (!fuse && G > 2 && use_cta) {
        GaussMma gp;
        memset(&gp, 0, sizeof(gp));
        gp.r = r;
        gp.mode = mode;
        gp.stages = 4;
        if (getenv("VA_GMC_STAGES")) gp.stages = atoi(getenv("VA_GMC_STAGES"));
        if (gp.stages < 2) gp.stages = 2;
        if (gp.stages > GMC_MAX_STAGES) gp.stages = GMC_MAX_STAGES;
        for (int i = 0; i < ksize; i++) gp.taps[i] = (unsigned char)taps[i];
        const int LW = 16 * GMC_WARPS + 2 * gm_hl(G, false);
        if ((unsigned)(LW / 4) > 256u) return VA_ERR_UNSUPPORTED;
        gp.strips = va_div_up(w, 16 * GMC_WARPS);
        const size_t smem = gmc_smem(G, gp.stages);
        if (smem > 220 * 1024) return VA_ERR_UNSUPPORTED;
        int ctas_per_sm = (int)((size_t)(224 * 1024) / (smem + 1024));
        const int tpw = getenv("VA_GMC_TPW") ? atoi(getenv("VA_GMC_TPW")) : 2;            // tiles per warp: 1, 2 or 4
        const int minb = getenv("VA_GM_MINB") ? atoi(getenv("VA_GM_MINB")) : 3;       // registers: 3 CTAs of 256 threads at <= 85
        if (ctas_per_sm > (tpw > 1 ? 4 : minb)) ctas_per_sm = tpw > 1 ? 4 : minb;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        const long long slots = (long long)ctx->sm_count * ctas_per_sm;
        int segs = (int)va_div_up(4 * slots, (long long)gp.strips * batch);
        const int max_segs = h / (64 * (G - 1)) > 0 ? h / (64 * (G - 1)) : 1;             // re-staged rows <= 25 %
        if (segs > max_segs) segs = max_segs;
        if (getenv("VA_GM_SEGS")) segs = atoi(getenv("VA_GM_SEGS"));
        if (segs < 1) segs = 1;
        gp.SH = 16 * va_div_up(h, 16 * segs);
        gp.segs = va_div_up(h, gp.SH);
        va_tmap map8, map1;
        if (va_tmap_encode(&map8, in, (size_t)w, h, batch, in_pitch, in_fstride, (unsigned)(LW / 4), 8) != 0 ||
            va_tmap_encode(&map1, in, (size_t)w, h, batch, in_pitch, in_fstride, (unsigned)(LW / 4), 1) != 0)
            return VA_ERR_UNSUPPORTED;
        const long long grid = (long long)gp.strips * gp.segs * batch;
        if (grid > 0x7fffffffll) return VA_ERR_UNSUPPORTED;
#define GMC_GO(GG, MB, TP)                                                                                       \
        do {                                                                                                     \
            auto kfn = gauss_mma_cta_kernel<GG, MB, TP>;                                                         \
            if (smem > 48 * 1024)                                                                                \
                VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            VA_LAUNCH(ctx, kfn, (unsigned)grid, 32 * GMC_WARPS / TP, smem, stream, map8, map1, out, out_pitch, out_fstride, w, h, batch, gp); \
        } while (0)
#define GMC_CASE(GG) case GG: if (tpw == 4) GMC_GO(GG, 4, 4); else if (tpw == 2) GMC_GO(GG, 4, 2); else if (minb >= 3) GMC_GO(GG, 3, 1); else GMC_GO(GG, 2, 1); break;
        switch (G) {
            GMC_CASE(3) GMC_CASE(4) GMC_CASE(5) GMC_CASE(6) GMC_CASE(7) GMC_CASE(8)
            default: return VA_ERR_UNSUPPORTED;
        }
#undef GMC_CASE
#undef GMC_GO
        (void)name;
        return VA_OK;
    }
    int tiles = G <= 2 ? 8 : 4;
    if (getenv("VA_GM_TILES")) tiles = atoi(getenv("VA_GM_TILES")) == 8 ? 8 : 4;
    if (G == 2) tiles = 8;
    GaussMma gp;
    memset(&gp, 0, sizeof(gp));
    gp.r = r;
    gp.mode = mode;
    gp.stages = fuse ? 2 : 4;
    if (getenv("VA_GM_STAGES")) gp.stages = atoi(getenv("VA_GM_STAGES"));
    if (gp.stages < 2) gp.stages = 2;
    if (gp.stages > GM_MAX_STAGES) gp.stages = GM_MAX_STAGES;
    for (int i = 0; i < ksize; i++) gp.taps[i] = (unsigned char)taps[i];
    const int LW = 16 * tiles + 2 * gm_hl(G, fuse);
    if ((unsigned)(C * LW / 4) > 256u) return VA_ERR_UNSUPPORTED;
    gp.strips = va_div_up(w, 16 * tiles);
    // segments: enough warp tasks to fill the machine a few times over, but every segment re-stages 16 (G - 1) rows
    const size_t wsm = gm_warp_smem(fuse, G, tiles, gp.stages);
    int wpc = 4;
    while (wpc > 1 && 512 + (size_t)wpc * wsm > 110 * 1024) wpc >>= 1;      // at least two CTAs per SM
    if (512 + (size_t)wpc * wsm > 220 * 1024) return VA_ERR_UNSUPPORTED;
    const size_t smem = 512 + (size_t)wpc * wsm;
    int ctas_per_sm = (int)((size_t)(224 * 1024) / (smem + 1024));
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const long long slots = (long long)ctx->sm_count * ctas_per_sm * wpc;
    int segs = (int)va_div_up(4 * slots, (long long)gp.strips * batch);
    // re-staged rows <= 25 % (<= 12.5 % for the small radii: VGA, 256 frames, fused: 0.094 -> 0.089 ms with 3 instead of 6 segments)
    const int seg_rows = G == 2 ? 128 : 64 * (G - 1);
    const int max_segs = h / seg_rows > 0 ? h / seg_rows : 1;
    if (segs > max_segs) segs = max_segs;
    if (getenv("VA_GM_SEGS")) segs = atoi(getenv("VA_GM_SEGS"));
    if (segs < 1) segs = 1;
    gp.SH = 16 * va_div_up(h, 16 * segs);
    gp.segs = va_div_up(h, gp.SH);
    va_tmap map8, map1;
    if (va_tmap_encode(&map8, in, (size_t)w * C, h, batch, in_pitch, in_fstride, (unsigned)(C * LW / 4), 8) != 0 ||
        va_tmap_encode(&map1, in, (size_t)w * C, h, batch, in_pitch, in_fstride, (unsigned)(C * LW / 4), 1) != 0)
        return VA_ERR_UNSUPPORTED;
    const long long tasks = (long long)gp.strips * gp.segs * batch;
    const long long grid = va_div_up(tasks, wpc);
    if (grid > 0x7fffffffll) return VA_ERR_UNSUPPORTED;
#define GM_GO(FUSE, GG, TT, MB)                                                                                  \
    do {                                                                                                         \
        auto kfn = gauss_mma_kernel<FUSE, GG, TT, MB>;                                                             \
        if (smem > 48 * 1024)                                                                                    \
            VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        VA_LAUNCH(ctx, kfn, (unsigned)grid, 32 * wpc, smem, stream, map8, map1, out, out_pitch, out_fstride, w, h, batch, gp); \
    } while (0)
#define GM_CASE(GG) case GG: if (tiles == 8) GM_GO(false, GG, 8, 1); else GM_GO(false, GG, 4, 1); break;
    const int minb = getenv("VA_GM_MINB") ? atoi(getenv("VA_GM_MINB")) : 4;
    switch (G) {
        case 2:
            if (fuse) { if (minb == 4) GM_GO(true, 2, 8, 4); else GM_GO(true, 2, 8, 3); }
            else { if (minb == 4) GM_GO(false, 2, 8, 4); else GM_GO(false, 2, 8, 3); }
            break;
        GM_CASE(3) GM_CASE(4) GM_CASE(5) GM_CASE(6) GM_CASE(7) GM_CASE(8)
        default: return VA_ERR_UNSUPPORTED;
    }
#undef GM_CASE
#undef GM_GO
    (void)name;
    return VA_OK;
}
