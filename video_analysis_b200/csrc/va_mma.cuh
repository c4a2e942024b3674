// va_mma.cuh -- the Blackwell-side building blocks of the MMA stencil kernels (va_gauss_mma.cu):
//   * exact integer tensor-core products   mma.sync.aligned.m16n8k32 / m16n8k16 .s32.u8.u8.s32   (SASS: IMMA.16832.U8.U8)
//   * TMA tile loads   cp.async.bulk.tensor.3d ... mbarrier::complete_tx::bytes   and the mbarrier calls around them
// plus thread-emulation versions (VA_EMU: tests/emu, index logic only) with the same fragment / box semantics.
//
// Fragment layouts (PTX ISA, "Matrix Fragments for mma.m16n8k32", 8-bit types), lane = 4 g + t:
//   A (16 x 32, row)  a0: row g,     k 4t..4t+3     a1: row g + 8, k 4t..4t+3     a2 / a3: the same rows, k 16 + 4t..
//   B (32 x 8,  col)  b0: k 4t..4t+3, col g         b1: k 16 + 4t.., col g
//   C (16 x 8,  s32)  c0: (g, 2t)  c1: (g, 2t + 1)  c2: (g + 8, 2t)  c3: (g + 8, 2t + 1)
// m16n8k16: A = a0, a1 (k 4t..4t+3), B = b0.  Bytes are packed little-endian: k = 4t + j is byte j.
#pragma once
#include "va_device.cuh"

#ifndef VA_EMU
#include <cuda.h>      // CUtensorMap and the enums of cuTensorMapEncodeTiled (types only; the entry point is looked up at run time)
#endif

// ---------------------------------------------------------------------------------------------------------
// integer MMA
// ---------------------------------------------------------------------------------------------------------
#ifdef VA_EMU
namespace emu {
// every lane publishes one word, every lane reads all 32 (two barriers)
inline void allgather(unsigned v, unsigned (&all)[32]) {
    Warp &w = *cur_block->warps[warp];
    w.slot[lane] = v;
    w.bar.arrive_and_wait();
    for (int i = 0; i < 32; i++) all[i] = (unsigned)w.slot[i];
    w.bar.arrive_and_wait();
}
inline void imma(int (&c)[4], const unsigned *a, int na, const unsigned *b, int nb) {
    // A[m][k]: k-halves of 16; B[k][n]
    unsigned A[4][32], B[2][32];
    for (int i = 0; i < na; i++) allgather(a[i], A[i]);
    for (int i = 0; i < nb; i++) allgather(b[i], B[i]);
    const int g = lane >> 2, t = lane & 3;
    for (int ci = 0; ci < 4; ci++) {
        const int m = g + ((ci & 2) ? 8 : 0), n = 2 * t + (ci & 1);
        int acc = c[ci];
        for (int half = 0; half < nb; half++)
            for (int tt = 0; tt < 4; tt++)
                for (int j = 0; j < 4; j++) {
                    // A row m: register index = (m >= 8) + 2 * half, lane = 4 * (m % 8) + tt
                    const unsigned aw = A[(m >= 8 ? 1 : 0) + 2 * half][4 * (m & 7) + tt];
                    const unsigned bw = B[half][4 * n + tt];
                    acc += (int)((aw >> (8 * j)) & 0xFF) * (int)((bw >> (8 * j)) & 0xFF);
                }
        c[ci] = acc;
    }
}
}  // namespace emu
#endif

__device__ __forceinline__ void va_imma_16832(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
#ifdef VA_EMU
    const unsigned b[2] = {b0, b1};
    emu::imma(c, a, 4, b, 2);
#else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
}
__device__ __forceinline__ void va_imma_16816(int (&c)[4], unsigned a0, unsigned a1, unsigned b0) {
#ifdef VA_EMU
    const unsigned a[2] = {a0, a1}, b[1] = {b0};
    emu::imma(c, a, 2, b, 1);
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
#endif
}

// ---------------------------------------------------------------------------------------------------------
// TMA: a 3-D tensor (bytes of a row as 32-bit elements, rows, frames) and boxes of (box_w32 elements, box_rows, 1)
// ---------------------------------------------------------------------------------------------------------
#ifdef VA_EMU
struct va_tmap {
    const uint8_t *base;
    unsigned w32, rows, frames;          // tensor extent
    size_t pitch, fstride;               // bytes
    unsigned box_w32, box_rows;
};
#else
typedef CUtensorMap va_tmap;
#endif

// host: describe `frames` images of `rows` rows of `row_bytes` bytes (multiple of 4; base, pitch, fstride multiples of 16)
int va_tmap_encode(va_tmap *map, const void *base, size_t row_bytes, size_t rows, size_t frames, size_t pitch, size_t fstride,
                   unsigned box_w32, unsigned box_rows);

__device__ __forceinline__ void va_mbar_init(uint64_t *bar, unsigned count) {
#ifdef VA_EMU
    *bar = 0;
    (void)count;
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
#endif
}
// make freshly initialised barriers visible to the async proxy (the TMA unit)
__device__ __forceinline__ void va_mbar_fence_init() {
#ifndef VA_EMU
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void va_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
#ifdef VA_EMU
    (void)bar; (void)bytes;
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void va_mbar_wait(uint64_t *bar, unsigned parity) {
#ifdef VA_EMU
    (void)bar; (void)parity;             // emulated TMA loads complete inside va_tma_load_3d, which the issuing lane
    __syncwarp();                        // has left by the time it arrives here (every lane of the warp waits)
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VA_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra VA_MBAR_DONE;\n"
        "bra VA_MBAR_WAIT;\n"
        "VA_MBAR_DONE:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
#endif
}
// box at element (x32, y, z) of the tensor -> dense rows at smem_dst; out-of-range elements arrive as zeros
__device__ __forceinline__ void va_tma_load_3d(void *smem_dst, const va_tmap *map, uint64_t *bar, int x32, int y, int z) {
#ifdef VA_EMU
    (void)bar;
    uint8_t *dst = reinterpret_cast<uint8_t *>(smem_dst);
    for (unsigned r = 0; r < map->box_rows; r++)
        for (unsigned e = 0; e < map->box_w32; e++) {
            const long long xx = (long long)x32 + e, yy = (long long)y + r;
            unsigned v = 0;
            if (xx >= 0 && xx < (long long)map->w32 && yy >= 0 && yy < (long long)map->rows && z >= 0 && (unsigned)z < map->frames)
                std::memcpy(&v, map->base + (size_t)z * map->fstride + (size_t)yy * map->pitch + 4 * (size_t)xx, 4);
            std::memcpy(dst + ((size_t)r * map->box_w32 + e) * 4, &v, 4);
        }
#else
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(x32), "r"(y), "r"(z)
                 : "memory");
#endif
}
