// va_ema.cu -- K3: running-average background + |difference| > thr -> packed bits,
// and the affine partial / carry folds used when a video is frame-sharded over GPUs.
//
// The recurrence is sequential in t, parallel in pixels: a thread owns PX pixels,
// keeps their float32 background in registers across the whole batch of T frames
// and streams the T u8 frames through (8 frames of loads in flight per thread, a cp.async ring in shared memory).
// Algorithmic HBM bytes per frame: N (u8 in) + N/8 (bits out) + 8N/T (state in/out).
//   d = float(x) - bg;  bit = |d| > thr;  bg = bg + alpha * d     (mul and add rounded
// separately -- on packed float32 pairs, see ema_step_packed -- so the result is bit-identical to the NumPy
// float32 oracle).
#include <cstdlib>

#include "va_device.cuh"

#define EMA_THREADS 256

template <int PX>
__device__ __forceinline__ void ema_load(const uint8_t *rp, int x, int w, bool vec, unsigned (&v)[PX / 4]) {
    if (vec && x + PX <= w) {
        if (PX == 16) {
            const uint4 q = va_ld_stream16(rp + x);
            v[0] = q.x; v[PX / 4 > 1 ? 1 : 0] = q.y; v[PX / 4 > 2 ? 2 : 0] = q.z; v[PX / 4 > 3 ? 3 : 0] = q.w;
        } else if (PX == 8) {
            const uint2 q = __ldg(reinterpret_cast<const uint2 *>(rp + x));
            v[0] = q.x; v[PX / 4 > 1 ? 1 : 0] = q.y;
        } else {
            v[0] = __ldg(reinterpret_cast<const unsigned *>(rp + x));
        }
    } else {
#pragma unroll
        for (int k = 0; k < PX / 4; k++) v[k] = 0;
        for (int i = 0; i < PX; i++)
            if (x + i < w) v[i >> 2] |= (unsigned)rp[x + i] << (8 * (i & 3));
    }
}

// byte i of `word` as an exact float: bits 0x4B0000bb = 2^23 + b, minus 2^23
__device__ __forceinline__ float ema_byte_to_float(unsigned word, int i) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u + i)) - 8388608.0f;
}

// one frame for the PX pixels of a thread: returns the PX mask bits, updates the state.
// |d| > thr is read off the sign of thr - |d| (negative exactly when the comparison holds: a
// float subtraction never rounds across zero, and thr == |d| gives +0) and shifted into the mask
// by one funnel shift per pixel, highest pixel first.  thr is finite and not -0 (checked by the host).
template <int PX>
__device__ __forceinline__ unsigned ema_step(const unsigned (&v)[PX / 4], float (&s)[PX], float alpha, float thr) {
    unsigned m = 0;
#pragma unroll
    for (int i = PX - 1; i >= 0; i--) {
        const float xf = ema_byte_to_float(v[i >> 2], i & 3);
        const float d = __fadd_rn(xf, -s[i]);
        m = __funnelshift_l(__float_as_uint(__fadd_rn(thr, -fabsf(d))), m, 1);
        s[i] = __fadd_rn(s[i], __fmul_rn(alpha, d));
    }
    return m;
}

// The same step on packed float32 pairs (sm_100's FADD2 / FFMA2: one instruction, two IEEE-754 results, each rounded
// exactly like its scalar counterpart): byte -> float, the difference and the state update of a pixel pair are one
// instruction each, 5 instead of 7 issue slots per pixel.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2
// (a single rounding -- not what NumPy does), so the product is written as fma(alpha, d, nz) with nz = -0.0f arriving as a
// kernel argument: x + (-0) == x for every x including both zeros, and an FFMA2 followed by an FADD2 is left alone.
typedef unsigned long long ema_f2;
#ifdef VA_EMU
// CPU emulation build: the same IEEE-754 operations, one lane at a time (compiled without -ffast-math / FMA contraction)
__device__ __forceinline__ ema_f2 ema_pack(float a, float b) {
    return (ema_f2)__float_as_uint(a) | ((ema_f2)__float_as_uint(b) << 32);
}
__device__ __forceinline__ void ema_unpack(ema_f2 v, float &a, float &b) {
    a = __uint_as_float((unsigned)v);
    b = __uint_as_float((unsigned)(v >> 32));
}
__device__ __forceinline__ ema_f2 ema_add2(ema_f2 a, ema_f2 b) {
    float a0, a1, b0, b1;
    ema_unpack(a, a0, a1);
    ema_unpack(b, b0, b1);
    return ema_pack(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
}
__device__ __forceinline__ ema_f2 ema_sub2(ema_f2 a, ema_f2 b) {
    float a0, a1, b0, b1;
    ema_unpack(a, a0, a1);
    ema_unpack(b, b0, b1);
    return ema_pack(__fadd_rn(a0, -b0), __fadd_rn(a1, -b1));
}
__device__ __forceinline__ ema_f2 ema_fma2(ema_f2 a, ema_f2 b, ema_f2 c) {
    float a0, a1, b0, b1, c0, c1;
    ema_unpack(a, a0, a1);
    ema_unpack(b, b0, b1);
    ema_unpack(c, c0, c1);
    return ema_pack(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
}
#else
__device__ __forceinline__ ema_f2 ema_pack(float a, float b) {
    ema_f2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void ema_unpack(ema_f2 v, float &a, float &b) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ ema_f2 ema_add2(ema_f2 a, ema_f2 b) {
    ema_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ ema_f2 ema_sub2(ema_f2 a, ema_f2 b) {
    ema_f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ ema_f2 ema_fma2(ema_f2 a, ema_f2 b, ema_f2 c) {
    ema_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
#endif

template <int PX>
__device__ __forceinline__ unsigned ema_step_packed(const unsigned (&v)[PX / 4], float (&s)[PX], float alpha, float thr, float nz) {
    unsigned m = 0;
    const ema_f2 alpha2 = ema_pack(alpha, alpha), nz2 = ema_pack(nz, nz), two23 = ema_pack(8388608.0f, 8388608.0f);
#pragma unroll
    for (int i = PX / 2 - 1; i >= 0; i--) {
        const unsigned word = v[i >> 1];
        const int b = 2 * (i & 1);
        // (2^23 + byte as float bits by PRMT on the alu pipe; the same by IDP.4A with a one-hot byte, i.e. on the
        // fmaheavy pipe, for all or half of the pixels measured slower: 0.088 / 0.084 against 0.082 ms per 128 frames)
        const ema_f2 big = ema_pack(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u + b)),
                                    __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u + b + 1)));
        const ema_f2 s2 = ema_pack(s[2 * i], s[2 * i + 1]);
        const ema_f2 d2 = ema_sub2(ema_sub2(big, two23), s2);
        float d0, d1;
        ema_unpack(d2, d0, d1);
        m = __funnelshift_l(__float_as_uint(__fadd_rn(thr, -fabsf(d1))), m, 1);
        m = __funnelshift_l(__float_as_uint(__fadd_rn(thr, -fabsf(d0))), m, 1);
        ema_unpack(ema_add2(s2, ema_fma2(alpha2, d2, nz2)), s[2 * i], s[2 * i + 1]);
    }
    return m;
}

// Frames are streamed through a per-thread ring in shared memory filled by cp.async: RING
// frames of loads are in flight per thread at no register cost, and nobody waits on a block
// barrier because a thread only ever reads the slots it filled itself.
#define EMA_RING 8

template <int PX>
__device__ __forceinline__ void ema_issue(unsigned *slot, const uint8_t *rp, int x, int w, bool fast) {
    if (fast) {
        if (PX == 16) va_cp_async16(slot, rp + x);
        else if (PX == 8) va_cp_async8(slot, rp + x);
        else va_cp_async4(slot, rp + x);
    } else if (x < w) {          // ragged row end / unaligned input: bytewise (lanes beyond the row do nothing)
        unsigned v[PX / 4];
        ema_load<PX>(rp, x, w, false, v);
#pragma unroll
        for (int k = 0; k < PX / 4; k++) slot[k] = v[k];
    }
    va_cp_async_commit();
}

// NT = threads per block.  The 16-pixel variant runs one warp per block: a warp's work (its pixels through the whole
// batch) is long and indivisible, so the block scheduler has to be able to even the warps out over the SMs -- with
// 8-warp blocks a 1080p launch (4050 warps, 27.4 per SM) left some SMs with 32 warps and others with 24 and ran
// as long as the fullest one.
// ALLFAST: every lane of every warp owns PX pixels inside the image, reads them with one aligned cp.async and holds its
// state in aligned float4s -- the common case (any width that is a multiple of 2 PX with aligned buffers).  The frame loop
// is then unrolled over one turn of the ring: slot addresses are immediates, the next load is a predicated LDGSTS and
// the bytewise path does not exist (122 -> 104 instructions per thread and frame).
template <int PX, int NT, int MODE, bool ALLFAST>
__global__ void __launch_bounds__(NT, NT == 32 ? 32 : PX == 16 ? 4 : 5)
ema_diff_thresh_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                       float *__restrict__ bg, size_t bg_pitch_e,
                       uint32_t *__restrict__ mask, size_t mask_pitch_w, size_t mask_fstride_w,
                       int w, int h, int batch, float alpha, float thr, int first_init, int vec_in, int vec_bg, int flat,
                       float nz) {
    constexpr int NW = PX / 4;
    constexpr int RING = PX >= 8 ? 8 : 16;                             // frames in flight per thread
    __shared__ uint4 ring_raw[RING * NT * NW / 4];                     // uint4: 16-byte aligned slots
    unsigned(*ring)[NT * NW] = reinterpret_cast<unsigned(*)[NT * NW]>(ring_raw);
    const int lane = threadIdx.x & 31;
    constexpr int warps_per_block = NT >> 5;
    const int span = 32 * PX;                         // pixels per warp step
    const int chunks = (w + span - 1) / span;
    // flat: w % (2 PX) == 0, so a warp takes 32 consecutive PX-pixel groups of the image in raster order (they may
    // straddle rows; lanes that share a mask word stay in one row) and no lane idles beyond the row end
    const int groups = w / PX;                                         // flat only
    const unsigned total = flat ? (unsigned)(((long long)groups * h + 31) >> 5) : (unsigned)((long long)chunks * h);
    unsigned *myslot = &ring[0][threadIdx.x * NW];
    constexpr int SLOT_STRIDE = NT * NW;

    for (unsigned item = blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < total;
         item += gridDim.x * warps_per_block) {
        int y, x;
        if (flat) {
            const unsigned g = item * 32u + lane;
            y = (int)(g / (unsigned)groups);
            x = y < h ? (int)(g - (unsigned)y * (unsigned)groups) * PX : w;       // beyond the image: an idle lane
        } else {
            y = (int)(item / (unsigned)chunks);
            x = (int)(item - (unsigned)y * (unsigned)chunks) * span + lane * PX;
        }
        const uint8_t *rp = in + (size_t)y * in_pitch;
        float *bgp = bg + (size_t)y * bg_pitch_e + x;
        const unsigned valid = (ALLFAST || x + PX <= w) ? ((1u << PX) - 1u) : x >= w ? 0u : ((1u << (w - x)) - 1u);
        const bool fast = ALLFAST || (vec_in && x + PX <= w);
        const bool fast_bg = ALLFAST || (vec_bg && x + PX <= w);

        // prologue: RING - 1 frames in flight
        const uint8_t *rnext = rp;                     // row y of the next frame to put in flight
#pragma unroll
        for (int u = 0; u < RING - 1; u++) {
            if (u < batch) ema_issue<PX>(myslot + u * SLOT_STRIDE, rnext, x, w, fast);
            else va_cp_async_commit();
            rnext += in_fstride;
        }

        float s[PX];
        if (fast_bg) {
#pragma unroll
            for (int k = 0; k < NW; k++) {
                const float4 q = *reinterpret_cast<const float4 *>(bgp + 4 * k);
                s[4 * k] = q.x; s[4 * k + 1] = q.y; s[4 * k + 2] = q.z; s[4 * k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < PX; i++) s[i] = (x + i < w) ? bgp[i] : 0.0f;
        }

        uint32_t *mrow = mask + (size_t)y * mask_pitch_w + (x >> 5);
        const bool writer = (lane & (32 / PX - 1)) == 0 && x < w;
        if constexpr (ALLFAST) {
            // one frame: `uslot` = ring slot of frame t (static), frame t + RING - 1 goes into the slot before it
            auto frame = [&](int t, int uslot, bool may_init) {
                if (t + RING - 1 < batch) {
                    if (PX == 16) va_cp_async16(myslot + ((uslot + RING - 1) % RING) * SLOT_STRIDE, rnext + x);
                    else if (PX == 8) va_cp_async8(myslot + ((uslot + RING - 1) % RING) * SLOT_STRIDE, rnext + x);
                    else va_cp_async4(myslot + ((uslot + RING - 1) % RING) * SLOT_STRIDE, rnext + x);
                }
                va_cp_async_commit();
                rnext += in_fstride;
                va_cp_async_wait_group<RING - 1>();          // frame t has landed
                unsigned v[NW];
                const unsigned *slot = myslot + uslot * SLOT_STRIDE;
                if (PX == 16) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(slot);
                    v[0] = q.x; v[NW > 1 ? 1 : 0] = q.y; v[NW > 2 ? 2 : 0] = q.z; v[NW > 3 ? 3 : 0] = q.w;
                } else if (PX == 8) {
                    const uint2 q = *reinterpret_cast<const uint2 *>(slot);
                    v[0] = q.x; v[NW > 1 ? 1 : 0] = q.y;
                } else {
                    v[0] = slot[0];
                }
                unsigned m;
                if (may_init && t == 0 && first_init) {
#pragma unroll
                    for (int i = 0; i < PX; i++) s[i] = ema_byte_to_float(v[i >> 2], i & 3);
                    m = 0;
                } else {
                    m = MODE ? ema_step_packed<PX>(v, s, alpha, thr, nz) : ema_step<PX>(v, s, alpha, thr);
                }
#pragma unroll
                for (int sh = PX, d = 1; sh < 32; sh <<= 1, d <<= 1) m |= __shfl_down_sync(0xffffffffu, m, d) << sh;
                if (writer) *mrow = m;
                mrow += mask_fstride_w;
            };
            int t0 = 0;
            for (; t0 + RING <= batch; t0 += RING) {
#pragma unroll
                for (int u = 0; u < RING; u++) frame(t0 + u, u, u == 0);
            }
#pragma unroll
            for (int u = 0; u < RING - 1; u++)
                if (t0 + u < batch) frame(t0 + u, u, u == 0);
        } else
        for (int t = 0; t < batch; t++) {
            const int tn = t + RING - 1;                 // frame to put in flight now
            if (tn < batch) ema_issue<PX>(myslot + (tn % RING) * SLOT_STRIDE, rnext, x, w, fast);
            else va_cp_async_commit();
            rnext += in_fstride;
            va_cp_async_wait_group<RING - 1>();          // frame t has landed
            unsigned v[NW];
            const unsigned *slot = myslot + (t % RING) * SLOT_STRIDE;
            if (PX == 16) {
                const uint4 q = *reinterpret_cast<const uint4 *>(slot);
                v[0] = q.x; v[NW > 1 ? 1 : 0] = q.y; v[NW > 2 ? 2 : 0] = q.z; v[NW > 3 ? 3 : 0] = q.w;
            } else if (PX == 8) {
                const uint2 q = *reinterpret_cast<const uint2 *>(slot);
                v[0] = q.x; v[NW > 1 ? 1 : 0] = q.y;
            } else {
                v[0] = slot[0];
            }
            unsigned m;
            if (t == 0 && first_init) {          // frame 0 initialises the model and gets an empty mask
#pragma unroll
                for (int i = 0; i < PX; i++) s[i] = ema_byte_to_float(v[i >> 2], i & 3);
                m = 0;
            } else {
                m = (MODE ? ema_step_packed<PX>(v, s, alpha, thr, nz) : ema_step<PX>(v, s, alpha, thr)) & valid;
            }
#pragma unroll
            for (int sh = PX, d = 1; sh < 32; sh <<= 1, d <<= 1) m |= __shfl_down_sync(0xffffffffu, m, d) << sh;
            if (writer) *mrow = m;
            mrow += mask_fstride_w;
        }
        va_cp_async_wait_group<0>();

        if (fast_bg) {
#pragma unroll
            for (int k = 0; k < NW; k++)
                *reinterpret_cast<float4 *>(bgp + 4 * k) = make_float4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < PX; i++)
                if (x + i < w) bgp[i] = s[i];
        }
    }
}

extern "C" int va_ema_diff_thresh(va_ctx *ctx, va_stream stream,
                                  const uint8_t *in, size_t in_pitch, size_t in_fstride,
                                  float *bg, size_t bg_pitch_e,
                                  uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                                  int w, int h, int batch, float alpha, float thr, int first_frame_inits) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && bg && mask, "va_ema_diff_thresh: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "va_ema_diff_thresh: bad size");
    VA_REQUIRE(ctx, mask_pitch_w >= (size_t)((w + 31) / 32) && bg_pitch_e >= (size_t)w, "va_ema_diff_thresh: pitch smaller than a row");
    VA_REQUIRE(ctx, thr - thr == 0.0f, "va_ema_diff_thresh: threshold must be finite");
    thr += 0.0f;                                        // -0 -> +0 (the kernel tests the sign of thr - |d|)
    const int vec_bg = va_aligned(bg, 16) && bg_pitch_e % 4 == 0;
    // pixels per thread: 16 while that gives every SM sub-partition two warps or more, 8 for smaller frames (VGA: 600 ->
    // 1200 one-warp blocks), 4 for tiny ones
    const int packed = getenv("VA_EMA_PACKED") ? atoi(getenv("VA_EMA_PACKED")) : 1;      // 0: the scalar arithmetic of round 1
    const long long warps16 = (long long)((w + 511) / 512) * h;
    int px = warps16 >= (long long)ctx->sm_count * 8 ? 16 : warps16 >= (long long)ctx->sm_count ? 8 : 4;
    if (getenv("VA_EMA_PX")) px = atoi(getenv("VA_EMA_PX")) == 16 ? 16 : atoi(getenv("VA_EMA_PX")) == 8 ? 8 : 4;     // tuning only
    if (px >= 8) {
        const int span = 32 * px;
        const int vec_in = va_aligned(in, px) && in_pitch % px == 0 && in_fstride % px == 0;
        const int flat = w % 32 == 0 && w % span != 0 && !getenv("VA_EMA_NOFLAT");
        const long long warps = flat ? ((long long)(w / px) * h + 31) / 32 : (long long)((w + span - 1) / span) * h;
        VA_REQUIRE(ctx, warps < (1ll << 31), "va_ema_diff_thresh: frame too large");
        const int nt = getenv("VA_EMA_NT") ? atoi(getenv("VA_EMA_NT")) : 32;                // tuning only
        if (nt == 256 && px == 16) {
            const int grid = va_grid(ctx, (warps + 7) / 8, 8);
            auto kfn = ema_diff_thresh_kernel<16, 256, 0, false>;
            VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, bg, bg_pitch_e, mask, mask_pitch_w,
                      mask_fstride_w, w, h, batch, alpha, thr, first_frame_inits, vec_in, vec_bg, flat, -0.0f);
            return VA_OK;
        }
        const int grid = (int)warps;                      // one warp per block, no grid-stride rounds
        const bool allfast = vec_in && vec_bg && (flat ? ((long long)(w / px) * h) % 32 == 0 : w % span == 0) && !getenv("VA_EMA_GENERAL");
        auto kfn = px == 16 ? (!packed ? ema_diff_thresh_kernel<16, 32, 0, false>
                               : allfast ? ema_diff_thresh_kernel<16, 32, 1, true> : ema_diff_thresh_kernel<16, 32, 1, false>)
                            : (allfast ? ema_diff_thresh_kernel<8, 32, 1, true> : ema_diff_thresh_kernel<8, 32, 1, false>);
        VA_LAUNCH(ctx, kfn, grid, 32, 0, stream, in, in_pitch, in_fstride, bg, bg_pitch_e, mask, mask_pitch_w,
                  mask_fstride_w, w, h, batch, alpha, thr, first_frame_inits, vec_in, vec_bg, flat, -0.0f);
    } else {
        const int vec_in = va_aligned(in, 4) && in_pitch % 4 == 0 && in_fstride % 4 == 0;
        const int flat = w % 32 == 0 && w % 128 != 0 && !getenv("VA_EMA_NOFLAT");
        const long long warps = flat ? ((long long)(w / 4) * h + 31) / 32 : (long long)((w + 127) / 128) * h;
        const int grid = va_grid(ctx, (warps + 7) / 8, 8);
        const bool allfast = vec_in && vec_bg && (flat ? ((long long)(w / 4) * h) % 32 == 0 : w % 128 == 0) && !getenv("VA_EMA_GENERAL");
        auto kfn = !packed ? ema_diff_thresh_kernel<4, EMA_THREADS, 0, false>
                 : allfast ? ema_diff_thresh_kernel<4, EMA_THREADS, 1, true> : ema_diff_thresh_kernel<4, EMA_THREADS, 1, false>;
        VA_LAUNCH(ctx, kfn, grid, EMA_THREADS, 0, stream, in, in_pitch, in_fstride, bg, bg_pitch_e, mask, mask_pitch_w,
                  mask_fstride_w, w, h, batch, alpha, thr, first_frame_inits, vec_in, vec_bg, flat, -0.0f);
    }
    return VA_OK;
}

// ---------------------------------------------------------------------------------
// frame-sharded EMA: partial fold from a zero state and carry combination
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ema_partial_kernel(const uint8_t *__restrict__ in, size_t in_pitch, size_t in_fstride,
                   float *__restrict__ S, size_t s_pitch_e, int w, int h, int batch,
                   float alpha, float a, int accumulate, int vec_in) {
    const int groups = (w + 3) >> 2;
    const unsigned total = (unsigned)((long long)groups * h);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += gridDim.x * blockDim.x) {
        const int y = (int)(i / (unsigned)groups);
        const int x = 4 * (int)(i - (unsigned)y * (unsigned)groups);
        const uint8_t *rp = in + (size_t)y * in_pitch;
        float *sp = S + (size_t)y * s_pitch_e + x;
        float s[4];
#pragma unroll
        for (int k = 0; k < 4; k++) s[k] = (accumulate && x + k < w) ? sp[k] : 0.0f;
        // eight frames of loads in flight per thread (the recurrence itself is sequential in t)
        for (int t0 = 0; t0 < batch; t0 += 8) {
            unsigned v[8][1];
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (t0 + u < batch) ema_load<4>(rp + (size_t)(t0 + u) * in_fstride, x, w, vec_in != 0, v[u]);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (t0 + u >= batch) break;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    s[k] = __fadd_rn(__fmul_rn(a, s[k]), __fmul_rn(alpha, ema_byte_to_float(v[u][0], k)));
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (x + k < w) sp[k] = s[k];
    }
}

extern "C" int va_ema_partial(va_ctx *ctx, va_stream stream,
                              const uint8_t *in, size_t in_pitch, size_t in_fstride,
                              float *S, size_t s_pitch_e, int w, int h, int batch,
                              float alpha, int accumulate) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, in && S, "va_ema_partial: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && s_pitch_e >= (size_t)w, "va_ema_partial: bad size");
    const int vec_in = va_aligned(in, 4) && in_pitch % 4 == 0 && in_fstride % 4 == 0;
    const long long items = (long long)((w + 3) / 4) * h;
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = ema_partial_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, in, in_pitch, in_fstride, S, s_pitch_e, w, h, batch,
              alpha, 1.0f - alpha, accumulate, vec_in);
    return VA_OK;
}

__global__ void __launch_bounds__(256)
ema_fold_kernel(float *__restrict__ carry, const float *__restrict__ S, size_t pitch_e, int w, int h, float scale) {
    const unsigned total = (unsigned)((long long)w * h);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += gridDim.x * blockDim.x) {
        const int y = (int)(i / (unsigned)w);
        const int x = (int)(i - (unsigned)y * (unsigned)w);
        const size_t o = (size_t)y * pitch_e + x;
        carry[o] = __fadd_rn(__fmul_rn(scale, carry[o]), S[o]);
    }
}

extern "C" int va_ema_fold(va_ctx *ctx, va_stream stream, float *carry, const float *S,
                           size_t pitch_e, int w, int h, float scale) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, carry && S, "va_ema_fold: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && pitch_e >= (size_t)w, "va_ema_fold: bad size");
    const long long items = (long long)w * h;
    const int grid = va_grid(ctx, (items + 255) / 256, 8);
    auto kfn = ema_fold_kernel;
    VA_LAUNCH(ctx, kfn, grid, 256, 0, stream, carry, S, pitch_e, w, h, scale);
    return VA_OK;
}
