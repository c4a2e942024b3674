// va_morph.cu -- K4: binary erode / dilate / open / close on packed bit masks.
//
// Replaces cv2.erode / cv2.dilate (video/analysis/image.py:248-256) and
// cv2.morphologyEx(MORPH_OPEN / MORPH_CLOSE) for structuring elements built by
// cv2.getStructuringElement(RECT | CROSS | ELLIPSE, (kx, ky)), default anchor
// (kx/2, ky/2) and OpenCV's default morphology border: pixels outside the image
// never win the min / max.  OpenCV evaluates BOTH erode and dilate as
//     dst(x, y) = op over {(i, j): se(i, j) != 0} of src(x + i - ax, y + j - ay),
// so no reflection of the element is applied (matters for even sizes).
//
// 32 pixels per word: a horizontal run [a, b] of the element becomes AND / OR of
// funnel-shifted words, evaluated by doubling in O(log(run length)).  Every row j of
// the element is one run (true for RECT, CROSS and ELLIPSE), so an op is
//     out(y) = AND/OR_j  hrun_j( in(y + j - ay) ).
// OPEN / CLOSE run both passes in one CTA on a band of rows; the intermediate stays
// in shared memory.  Algorithmic HBM bytes: N/8 in + N/8 out per frame.
// Odd square RECT elements up to 7x7 (the common case) take the register-only streaming kernel below.
#include <cmath>
#include <cstdlib>

#include "va_device.cuh"

#define MORPH_THREADS 256
#define MORPH_MAX_K 63
#define MORPH_BAND 32

struct MorphSE {
    int kx, ky, ax, ay;
    signed char a[MORPH_MAX_K];   // run of row j: offsets a[j]..b[j] relative to the anchor; a > b: empty row
    signed char b[MORPH_MAX_K];
};

// cv::getStructuringElement restated as per-row runs
static int morph_build_se(int shape, int kx, int ky, MorphSE *se) {
    if (kx < 1 || ky < 1 || kx > MORPH_MAX_K || ky > MORPH_MAX_K) return VA_ERR_UNSUPPORTED;
    if (shape != VA_SE_RECT && shape != VA_SE_CROSS && shape != VA_SE_ELLIPSE) return VA_ERR_INVALID;
    memset(se, 0, sizeof(*se));
    se->kx = kx; se->ky = ky; se->ax = kx / 2; se->ay = ky / 2;
    if (kx == 1 && ky == 1) shape = VA_SE_RECT;
    const int r = ky / 2, c = kx / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < ky; i++) {
        int j1 = 0, j2 = 0;
        if (shape == VA_SE_RECT || (shape == VA_SE_CROSS && i == r)) {
            j2 = kx;
        } else if (shape == VA_SE_CROSS) {
            j1 = c; j2 = c + 1;
        } else {
            const int dy = i - r;
            if (std::abs(dy) <= r) {
                const int dx = (int)std::lrint(c * std::sqrt((r * r - dy * dy) * inv_r2));   // cvRound
                j1 = c - dx > 0 ? c - dx : 0;
                j2 = c + dx + 1 < kx ? c + dx + 1 : kx;
            }
        }
        se->a[i] = (signed char)(j1 - se->ax);
        se->b[i] = (signed char)(j2 - 1 - se->ax);
    }
    return VA_OK;
}

// AND (ERODE) or OR over src(x + d), d in [a, b], for the 32 pixels of the middle word of
// the window (prev, cur, next).  |a|, |b| <= 31.
template <bool ERODE>
__device__ __forceinline__ unsigned morph_hrun(unsigned prev, unsigned cur, unsigned next, int a, int b) {
    // acc_L(x) = op_{i<L} src(x + i) by doubling on the 96-bit window, then read at x + a
    unsigned w0 = prev, w1 = cur, w2 = next;
    const int L = b - a + 1;
    int len = 1;
    while (len < L) {
        const int s = (2 * len <= L) ? len : L - len;
        const unsigned n0 = __funnelshift_r(w0, w1, s), n1 = __funnelshift_r(w1, w2, s);
        const unsigned n2 = ERODE ? ((w2 >> s) | ~(0xffffffffu >> s)) : (w2 >> s);
        if (ERODE) { w0 &= n0; w1 &= n1; w2 &= n2; } else { w0 |= n0; w1 |= n1; w2 |= n2; }
        len += s;
    }
    // bits [32 + a, 64 + a) of the window
    if (a == 0) return w1;
    if (a > 0) return __funnelshift_r(w1, w2, a);
    return __funnelshift_l(w0, w1, -a);
}

// one output word of a morphological pass over rows held in shared memory.
//   src rows: index sr = (image row) - (first staged row), words [0, wpw); rows outside the
//   staged band are outside the image (the caller sizes the halo) and never win.
// RECT: every row of the element has the same run, so the ky rows are combined first
// (3 words each) and the horizontal run is evaluated once.
// KSQ > 0: odd square K x K element known at compile time (3x3, 5x5, 7x7): everything unrolled.
// Every row the element touches is inside the staged band (the caller stages the halo, rows
// outside the image hold the identity), so no range checks are needed.
template <bool ERODE, bool RECT, int KSQ>
__device__ __forceinline__ unsigned morph_word(const unsigned *src, int src_rows, int wpw, int sr, int j,
                                               const MorphSE &se) {
    const unsigned ident = ERODE ? 0xffffffffu : 0u;
    const bool has_p = j > 0, has_n = j + 1 < wpw;
    if (KSQ > 0) {
        constexpr int H = KSQ / 2;
        const unsigned *row = src + (sr - H) * wpw + j;
        unsigned p = ident, c = ident, n = ident;
#pragma unroll
        for (int i = 0; i < KSQ; i++, row += wpw) {
            const unsigned vp = has_p ? row[-1] : ident, vc = row[0], vn = has_n ? row[1] : ident;
            if (ERODE) { p &= vp; c &= vc; n &= vn; } else { p |= vp; c |= vc; n |= vn; }
        }
        unsigned acc = c;
#pragma unroll
        for (int d = 1; d <= H; d++) {
            const unsigned r = __funnelshift_r(c, n, d), l = __funnelshift_l(p, c, d);
            if (ERODE) acc &= r & l; else acc |= r | l;
        }
        return acc;
    }
    if (RECT) {
        unsigned p = ident, c = ident, n = ident;
        int r0 = sr - se.ay, r1 = r0 + se.ky;
        r0 = r0 < 0 ? 0 : r0;
        r1 = r1 > src_rows ? src_rows : r1;
        const unsigned *row = src + r0 * wpw + j;
        for (int rr = r0; rr < r1; rr++, row += wpw) {
            const unsigned vp = has_p ? row[-1] : ident, vc = row[0], vn = has_n ? row[1] : ident;
            if (ERODE) { p &= vp; c &= vc; n &= vn; } else { p |= vp; c |= vc; n |= vn; }
        }
        return morph_hrun<ERODE>(p, c, n, se.a[0], se.b[0]);
    }
    unsigned acc = ident;
    for (int i = 0; i < se.ky; i++) {
        const int a = se.a[i], b = se.b[i];
        if (a > b) continue;
        const int rr = sr + i - se.ay;
        if (rr < 0 || rr >= src_rows) continue;
        const unsigned *row = src + rr * wpw + j;
        const unsigned v = morph_hrun<ERODE>(has_p ? row[-1] : ident, row[0], has_n ? row[1] : ident, a, b);
        if (ERODE) acc &= v; else acc |= v;
    }
    return acc;
}

template <bool RECT, int KSQ>
__device__ __forceinline__ unsigned morph_word_op(int erode, const unsigned *src, int src_rows, int wpw, int sr, int j,
                                                  const MorphSE &se) {
    return erode ? morph_word<true, RECT, KSQ>(src, src_rows, wpw, sr, j, se)
                 : morph_word<false, RECT, KSQ>(src, src_rows, wpw, sr, j, se);
}

// first  : 0 erode, 1 dilate;  second: -1 none, 0 erode, 1 dilate
// work split: warp = row of the band, lane = word of the row
template <bool RECT, int KSQ>
__global__ void __launch_bounds__(MORPH_THREADS)
morph_bits_kernel(const uint32_t *__restrict__ in, size_t in_pitch_w, size_t in_fstride_w,
                  uint32_t *__restrict__ out, size_t out_pitch_w, size_t out_fstride_w,
                  int w, int h, int bands_per_frame, int n_bands, int first, int second,
                  const __grid_constant__ MorphSE se) {
    VA_DYN_SMEM(unsigned, smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NWARP = MORPH_THREADS / 32;
    const int wpw = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    const int up = se.ay, dn = se.ky - 1 - se.ay;          // rows needed above / below per pass
    const int passes = second >= 0 ? 2 : 1;
    const int rows_a = MORPH_BAND + passes * (up + dn);    // staged input rows
    const int rows_b = MORPH_BAND + (up + dn);             // intermediate rows (two-pass only)
    unsigned *sa = smem;
    unsigned *sb = smem + rows_a * wpw;
    const unsigned id1 = first == 0 ? 0xffffffffu : 0u;
    const unsigned id2 = second == 0 ? 0xffffffffu : 0u;

    for (int band = blockIdx.x; band < n_bands; band += gridDim.x) {
        const int b = band / bands_per_frame;
        const int y0 = (band - b * bands_per_frame) * MORPH_BAND;
        const int nrows = min(MORPH_BAND, h - y0);
        const uint32_t *fin = in + (size_t)b * in_fstride_w;
        uint32_t *fout = out + (size_t)b * out_fstride_w;

        // ---- stage input rows [ya0, ya0 + rows_a); outside the image = identity of the first op
        const int ya0 = y0 - passes * up;
        for (int rr = warp; rr < rows_a; rr += NWARP) {
            const int y = ya0 + rr;
            const bool inside = y >= 0 && y < h;
            const uint32_t *grow = fin + (size_t)(inside ? y : 0) * in_pitch_w;
            for (int j = lane; j < wpw; j += 32) {
                unsigned v = id1;
                if (inside) {
                    v = grow[j];
                    if (j == wpw - 1) v = first == 0 ? (v | ~lastmask) : (v & lastmask);
                }
                sa[rr * wpw + j] = v;
            }
        }
        __syncthreads();

        if (passes == 1) {
            for (int rr = warp; rr < nrows; rr += NWARP) {
                uint32_t *orow = fout + (size_t)(y0 + rr) * out_pitch_w;
                for (int j = lane; j < wpw; j += 32) {
                    unsigned v = morph_word_op<RECT, KSQ>(first == 0, sa, rows_a, wpw, rr + up, j, se);
                    if (j == wpw - 1) v &= lastmask;
                    orow[j] = v;
                }
            }
        } else {
            // ---- first pass -> intermediate rows [yb0, yb0 + rows_b); rows outside the image and
            //      bits beyond w take the identity of the SECOND op
            const int yb0 = y0 - up;
            for (int rr = warp; rr < rows_b; rr += NWARP) {
                const int y = yb0 + rr;
                const bool inside = y >= 0 && y < h;
                for (int j = lane; j < wpw; j += 32) {
                    unsigned v = id2;
                    if (inside) {
                        v = morph_word_op<RECT, KSQ>(first == 0, sa, rows_a, wpw, rr + up, j, se);
                        if (j == wpw - 1) v = second == 0 ? (v | ~lastmask) : (v & lastmask);
                    }
                    sb[rr * wpw + j] = v;
                }
            }
            __syncthreads();
            for (int rr = warp; rr < nrows; rr += NWARP) {
                uint32_t *orow = fout + (size_t)(y0 + rr) * out_pitch_w;
                for (int j = lane; j < wpw; j += 32) {
                    unsigned v = morph_word_op<RECT, KSQ>(second == 0, sb, rows_b, wpw, rr + up, j, se);
                    if (j == wpw - 1) v &= lastmask;
                    orow[j] = v;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// streaming kernel for odd square RECT elements (3x3, 5x5, 7x7): a warp owns a strip of 64 words and
// walks down a segment of rows with everything in registers -- lane = two adjacent words, horizontal
// neighbours by shuffle, the last K horizontally reduced rows in a register window.  Erosion is dilation of the
// complement, so both passes of an OPEN / CLOSE are dilations on data XORed with an all-ones / zero
// mask; "outside the image never wins" and the bits beyond the row end are then plain zeros.
// The second pass consumes the first one's rows as they appear (lag K/2 rows).  A pass spoils one
// word at either end of the strip, so strips overlap by `passes` words on each side.
// ---------------------------------------------------------------------------------
#define MORPH_S_THREADS 128
#define MORPH_S_PREFETCH 8

// horizontal dilation of the two adjacent words (a, b) of every lane; words of neighbouring lanes
// arrive by shuffle (the end lanes get their own word back: they are halo lanes, see above)
template <int K>
__device__ __forceinline__ void morph_dilate_h2(unsigned &a, unsigned &b) {
    const unsigned p = __shfl_up_sync(0xffffffffu, b, 1), n = __shfl_down_sync(0xffffffffu, a, 1);
    unsigned ra = a, rb = b;
#pragma unroll
    for (int d = 1; d <= K / 2; d++) {
        ra |= __funnelshift_r(a, b, d) | __funnelshift_l(p, a, d);
        rb |= __funnelshift_r(b, n, d) | __funnelshift_l(a, b, d);
    }
    a = ra;
    b = rb;
}

// FAST: rows of an even number of words, 8-byte aligned, two passes -- every lane's word pair is then either inside the
// row or outside it and either written or not, so all lanes load their pair unconditionally (lanes outside the row from a
// clamped position: their bits are masked away) and a row costs one predicated load and one predicated store instead of
// the nested branches of the ragged case (30 % of the instructions of the general kernel are branches and predicates)
template <int K, int PASSES, bool FAST>
__global__ void __launch_bounds__(MORPH_S_THREADS)
morph_stream_kernel(const uint32_t *__restrict__ in, size_t in_pitch_w, size_t in_fstride_w,
                    uint32_t *__restrict__ out, size_t out_pitch_w, size_t out_fstride_w,
                    int w, int h, int strips, int segs, int SH, int n_warps, int first, int second, int vec) {
    constexpr int H = K / 2;
    constexpr int USABLE = 64 - 2 * PASSES;                    // words of a strip that are written
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * (MORPH_S_THREADS / 32) + (threadIdx.x >> 5);
    if (wid >= n_warps) return;
    const int b = wid / (strips * segs);
    const int rem = wid - b * strips * segs;
    const int seg = rem / strips, strip = rem - seg * strips;
    const int wpw = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    // lane = two adjacent words ja, ja + 1; strips start PASSES words early (even, so that pairs stay 8-byte aligned)
    const int ja = strip * USABLE - PASSES + 2 * lane;
    const bool ina = ja >= 0 && ja < wpw, inb = ja + 1 >= 0 && ja + 1 < wpw;
    const unsigned keepa = !ina ? 0u : (ja == wpw - 1 ? lastmask : 0xffffffffu);       // bits of the row in each word
    const unsigned keepb = !inb ? 0u : (ja + 1 == wpw - 1 ? lastmask : 0xffffffffu);
    const bool wra = ina && 2 * lane >= PASSES && 2 * lane < 64 - PASSES;
    const bool wrb = inb && 2 * lane + 1 >= PASSES && 2 * lane + 1 < 64 - PASSES;
    const bool pair = FAST || (vec && ina && inb);                                      // one 8-byte access
    const unsigned m1 = first == 0 ? 0xffffffffu : 0u;         // complement masks: erode = ~dilate(~x)
    const unsigned m2 = second == 0 ? 0xffffffffu : 0u;
    const unsigned m12 = m1 ^ m2;
    const int ys = seg * SH, ye = min(ys + SH, h);             // output rows of this warp
    const int y_first = ys - PASSES * H;                       // first input row needed
    const int n_in = (ye - ys) + 2 * PASSES * H;
    // row pointers run along with the walk (the ones outside the image are never dereferenced)
    const int jl = FAST ? min(max(ja, 0), wpw - 2) : ja;       // FAST: lanes outside the row read inside it (and are masked)
    const uint32_t *src = in + (size_t)b * in_fstride_w + (ptrdiff_t)jl + (ptrdiff_t)y_first * (ptrdiff_t)in_pitch_w;
    uint32_t *dst = out + (size_t)b * out_fstride_w + (ptrdiff_t)ja + (ptrdiff_t)(y_first - PASSES * H) * (ptrdiff_t)out_pitch_w;

    unsigned hwa[K], hwb[K], ewa[K], ewb[K];                   // windows of horizontally reduced rows
#pragma unroll
    for (int k = 0; k < K; k++) hwa[k] = hwb[k] = ewa[k] = ewb[k] = 0u;
    // every lane runs the same instruction stream (the shuffles need the whole warp); rows beyond the
    // segment in the last batch are loaded like any other and their results are not stored
    for (int i0 = 0; i0 < n_in; i0 += MORPH_S_PREFETCH) {
        unsigned ra[MORPH_S_PREFETCH], rb[MORPH_S_PREFETCH];
#pragma unroll
        for (int u = 0; u < MORPH_S_PREFETCH; u++) {
            const bool yin = (unsigned)(y_first + i0 + u) < (unsigned)h;
            ra[u] = rb[u] = m1;                                 // complemented to 0 below
            if (yin && pair) {
                const uint2 q = __ldg(reinterpret_cast<const uint2 *>(src));
                ra[u] = q.x; rb[u] = q.y;
            } else if (yin) {
                if (ina) ra[u] = __ldg(src);
                if (inb) rb[u] = __ldg(src + 1);
            }
            src += in_pitch_w;
        }
#pragma unroll
        for (int u = 0; u < MORPH_S_PREFETCH; u++) {
            const int i = i0 + u;
            // ---- pass 1 on input row y_first + i; its result is row yc = y_first + i - H
#pragma unroll
            for (int k = 0; k < K - 1; k++) { hwa[k] = hwa[k + 1]; hwb[k] = hwb[k + 1]; }
            unsigned xa = (ra[u] ^ m1) & keepa, xb = (rb[u] ^ m1) & keepb;
            morph_dilate_h2<K>(xa, xb);
            hwa[K - 1] = xa; hwb[K - 1] = xb;
            unsigned ea = hwa[0], eb = hwb[0];
#pragma unroll
            for (int k = 1; k < K; k++) { ea |= hwa[k]; eb |= hwb[k]; }
            const int yc = y_first + i - H;
            unsigned oa, ob;
            bool store;
            if (PASSES == 1) {
                oa = (ea ^ m1) & keepa; ob = (eb ^ m1) & keepb;
                store = i >= 2 * H && yc < ye;
            } else {
                // ---- pass 2 on row yc of the intermediate (rows outside the image never win)
                const bool cin = (unsigned)yc < (unsigned)h;
                unsigned ma = cin ? ((ea ^ m12) & keepa) : 0u, mb = cin ? ((eb ^ m12) & keepb) : 0u;
#pragma unroll
                for (int k = 0; k < K - 1; k++) { ewa[k] = ewa[k + 1]; ewb[k] = ewb[k + 1]; }
                morph_dilate_h2<K>(ma, mb);
                ewa[K - 1] = ma; ewb[K - 1] = mb;
                oa = ewa[0]; ob = ewb[0];
#pragma unroll
                for (int k = 1; k < K; k++) { oa |= ewa[k]; ob |= ewb[k]; }
                oa = (oa ^ m2) & keepa; ob = (ob ^ m2) & keepb;
                store = i >= 4 * H && yc - H < ye;
            }
            if (FAST) {
                if (store && wra) *reinterpret_cast<uint2 *>(dst) = make_uint2(oa, ob);
            } else if (store) {
                if (wra && wrb && pair) {
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(oa, ob);
                } else {
                    if (wra) dst[0] = oa;
                    if (wrb) dst[1] = ob;
                }
            }
            dst += out_pitch_w;
        }
    }
}

extern "C" int va_morph_bits(va_ctx *ctx, va_stream stream,
                             const uint32_t *in, size_t in_pitch_w, size_t in_fstride_w,
                             uint32_t *out, size_t out_pitch_w, size_t out_fstride_w,
                             int w, int h, int batch, int op, int shape, int kx, int ky) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && out && in != out, "va_morph_bits: null or aliased pointers");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "va_morph_bits: bad size");
    VA_REQUIRE(ctx, op >= VA_MORPH_ERODE && op <= VA_MORPH_CLOSE, "va_morph_bits: unknown morphological operation %d", op);
    const size_t wpw = (size_t)(w + 31) / 32;
    VA_REQUIRE(ctx, in_pitch_w >= wpw && out_pitch_w >= wpw, "va_morph_bits: pitch smaller than a row");
    MorphSE se;
    const int rc = morph_build_se(shape, kx, ky, &se);
    if (rc != VA_OK) VA_FAIL(ctx, rc, "va_morph_bits: structuring element shape %d size %dx%d not supported (max %d)", shape, kx, ky, MORPH_MAX_K);
    int first, second;
    switch (op) {
        case VA_MORPH_ERODE: first = 0; second = -1; break;
        case VA_MORPH_DILATE: first = 1; second = -1; break;
        case VA_MORPH_OPEN: first = 0; second = 1; break;
        default: first = 1; second = 0; break;
    }
    const int passes = second >= 0 ? 2 : 1;
    {
        bool sq = shape == VA_SE_RECT && kx == ky && (kx == 3 || kx == 5 || kx == 7);
        const char *env = getenv("VA_MORPH_STREAM");
        if (env && atoi(env) == 0) sq = false;                   // A-B checks against the band kernel
        if (sq) {
            const int usable = 64 - 2 * passes;
            const int strips = va_div_up((long long)wpw, usable);
            // rows per segment: the walk is a chain of load latencies, so many short segments (about 32
            // warps per SM -- measured 16 / 24 / 32 / 48 / 96: 26.6 / 24.6 / 22.5 / 24.1 / 23.8 us per 128 1080p frames --, at least 16 rows; each segment re-reads passes * (k - 1) rows), sized so that
            // the rows a warp reads fill whole prefetch batches
            const int halo_rows = passes * (kx - 1);
            const int wps = getenv("VA_MORPH_WPS") ? atoi(getenv("VA_MORPH_WPS")) : 32;      // tuning only
            int segs = va_div_up((long long)ctx->sm_count * wps, (long long)strips * batch);
            if (segs > h / 16) segs = h / 16;
            if (segs < 1) segs = 1;
            int SH = va_div_up(h, segs);
            SH = va_div_up(SH + halo_rows, MORPH_S_PREFETCH) * MORPH_S_PREFETCH - halo_rows;
            if (SH < 1) SH = MORPH_S_PREFETCH;
            segs = va_div_up(h, SH);
            // 8-byte accesses need even word offsets: rows start 8-byte aligned and strips start at even words
            const int vec = va_aligned(in, 8) && va_aligned(out, 8) && in_pitch_w % 2 == 0 && out_pitch_w % 2 == 0 &&
                            in_fstride_w % 2 == 0 && out_fstride_w % 2 == 0 && passes % 2 == 0;
            const bool fast = vec && wpw % 2 == 0 && wpw >= 2 && !getenv("VA_MORPH_GENERAL");
            const long long n_warps = (long long)strips * segs * batch;
            VA_REQUIRE(ctx, n_warps < (1ll << 31), "va_morph_bits: too many strips");
            const int grid = va_div_up(n_warps, MORPH_S_THREADS / 32);
#define MORPH_SGO(KK, PP)                                                                                  \
            do {                                                                                           \
                auto kfn = (PP == 2 && fast) ? morph_stream_kernel<KK, PP, PP == 2> : morph_stream_kernel<KK, PP, false>; \
                VA_LAUNCH(ctx, kfn, grid, MORPH_S_THREADS, 0, stream, in, in_pitch_w, in_fstride_w, out,   \
                          out_pitch_w, out_fstride_w, w, h, strips, segs, SH, (int)n_warps, first, second, vec); \
            } while (0)
#define MORPH_SK(KK) do { if (passes == 2) MORPH_SGO(KK, 2); else MORPH_SGO(KK, 1); } while (0)
            if (kx == 3) MORPH_SK(3); else if (kx == 5) MORPH_SK(5); else MORPH_SK(7);
#undef MORPH_SK
#undef MORPH_SGO
            return VA_OK;
        }
    }
    const int halo = ky - 1;
    const size_t smem = ((size_t)(MORPH_BAND + passes * halo) + (passes == 2 ? (size_t)(MORPH_BAND + halo) : 0)) * wpw * 4;
    VA_REQUIRE(ctx, smem <= 200 * 1024, "va_morph_bits: %d-pixel rows with a %d-row element do not fit in shared memory", w, ky);
    const int bands_per_frame = va_div_up(h, MORPH_BAND);
    const int n_bands = bands_per_frame * batch;
    bool rect = true;
    for (int i = 1; i < ky; i++) rect = rect && se.a[i] == se.a[0] && se.b[i] == se.b[0];
    const int grid = va_grid(ctx, n_bands, 8);
#define MORPH_GO(RECT, KSQ)                                                                                       \
    do {                                                                                                          \
        auto kfn = morph_bits_kernel<RECT, KSQ>;                                                                  \
        if (smem > 48 * 1024)                                                                                     \
            VA_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        VA_LAUNCH(ctx, kfn, grid, MORPH_THREADS, smem, stream, in, in_pitch_w, in_fstride_w, out, out_pitch_w,    \
                  out_fstride_w, w, h, bands_per_frame, n_bands, first, second, se);                              \
    } while (0)
    if (rect && kx == ky && kx == 3) MORPH_GO(true, 3);
    else if (rect && kx == ky && kx == 5) MORPH_GO(true, 5);
    else if (rect && kx == ky && kx == 7) MORPH_GO(true, 7);
    else if (rect) MORPH_GO(true, 0);
    else MORPH_GO(false, 0);
#undef MORPH_GO
    return VA_OK;
}
