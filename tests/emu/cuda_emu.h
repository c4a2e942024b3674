// cuda_emu.h -- TEST INFRASTRUCTURE ONLY (development aid, never shipped).
//
// A tiny CUDA-on-CPU shim: it lets the *same* csrc/*.cu kernel sources be
// compiled by g++ (-DVA_EMU) into tests/emu/libva_b200_emu.so so that index
// arithmetic, halo handling, bit tricks and union-find logic can be checked
// on the GPU-less development container before GPU minutes are spent.  Each
// CUDA thread is a real OS thread, __syncthreads() and the warp collectives are
// std::barriers.  It says nothing about performance, memory-model races or
// real-hardware behaviour; the `-m gpu` parity tests are the gate.
//
// The product package (video_analysis_b200) never loads the emulated library:
// video_analysis_b200/_lib.py only ever opens csrc/libva_b200.so and raises if
// it is missing.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) alignas(n)
#define __grid_constant__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort2 { unsigned short x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

namespace emu {
struct Warp {
    std::barrier<> bar;
    uint64_t slot[32];
    explicit Warp(int lanes) : bar(lanes) {}
};
struct Block {
    std::barrier<> bar;
    std::atomic<int> vote{0};
    std::vector<std::unique_ptr<Warp>> warps;
    explicit Block(int n) : bar(n) {
        for (int i = 0; i < n; i += 32) warps.emplace_back(new Warp(std::min(32, n - i)));
    }
};
inline thread_local Block *cur_block = nullptr;
inline thread_local int lane = 0, warp = 0;
inline unsigned char *dyn_smem_ptr = nullptr;
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

static inline void __syncthreads() { emu::cur_block->bar.arrive_and_wait(); }
static inline int __syncthreads_or(int pred) {
    emu::Block *b = emu::cur_block;
    if (pred) b->vote.store(1);
    b->bar.arrive_and_wait();
    const int r = b->vote.load();
    b->bar.arrive_and_wait();
    b->vote.store(0);
    b->bar.arrive_and_wait();
    return r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::cur_block->warps[emu::warp]->bar.arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace emu {
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
template <class T> inline T exchange(T v, int src_lane) {
    Warp &w = *cur_block->warps[warp];
    w.slot[lane] = to_bits(v);
    w.bar.arrive_and_wait();
    T r = from_bits<T>(w.slot[src_lane & 31]);
    w.bar.arrive_and_wait();
    return r;
}
}  // namespace emu

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int base = emu::lane & ~(width - 1);
    return emu::exchange(v, base + (src & (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    int src = emu::lane ^ m;
    if ((src & ~(width - 1)) != (emu::lane & ~(width - 1))) src = emu::lane;
    return emu::exchange(v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int base = emu::lane & ~(width - 1);
    int src = emu::lane - (int)d;
    if (src < base) src = emu::lane;
    return emu::exchange(v, src);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int base = emu::lane & ~(width - 1);
    int src = emu::lane + (int)d;
    if (src >= base + width) src = emu::lane;
    return emu::exchange(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Warp &w = *emu::cur_block->warps[emu::warp];
    w.slot[emu::lane] = pred ? 1 : 0;
    w.bar.arrive_and_wait();
    unsigned r = 0;
    int lanes = std::min(32u, blockDim.x * blockDim.y * blockDim.z - 32u * emu::warp);
    for (int i = 0; i < lanes; i++) r |= (unsigned)(w.slot[i] & 1) << i;
    w.bar.arrive_and_wait();
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
    int lanes = std::min(32u, blockDim.x * blockDim.y * blockDim.z - 32u * emu::warp);
    unsigned full = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
    return __ballot_sync(m, pred) == full;
}

// ---- integer / bit intrinsics ------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (sh & 31)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
    return c;
}
static inline unsigned __dp2a_lo(unsigned a, unsigned b, unsigned c) {
    return c + (a & 0xFFFF) * (b & 0xFF) + (a >> 16) * ((b >> 8) & 0xFF);
}
static inline unsigned __dp2a_hi(unsigned a, unsigned b, unsigned c) {
    return c + (a & 0xFFFF) * ((b >> 16) & 0xFF) + (a >> 16) * ((b >> 24) & 0xFF);
}
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline int __float2int_rn(float a) { return (int)std::nearbyintf(a); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(u); }
static inline unsigned __float_as_uint(float f) { return (unsigned)emu::to_bits(f); }
static inline float __int_as_float(int u) { return emu::from_bits<float>((unsigned)u); }
static inline int __float_as_int(float f) { return (int)emu::to_bits(f); }
template <class T> static inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

// ---- atomics -------------------------------------------------------------------
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float *p, float v) {
    unsigned *u = reinterpret_cast<unsigned *>(p);
    unsigned old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nv;
    do { nv = __float_as_uint(__uint_as_float(old) + v); }
    while (!__atomic_compare_exchange_n(u, &old, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return __uint_as_float(old);
}
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
template <class T> static inline T atomicMin(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <class T> static inline T atomicMax(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

// ---- runtime shims -------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 2; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memcpy(d, s, n);
    return cudaSuccess;
}

namespace emu {
template <class K, class... A>
inline void launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
    int nthreads = block.x * block.y * block.z;
    std::vector<unsigned char> dyn(smem + 64);
    dyn_smem_ptr = reinterpret_cast<unsigned char *>(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block blk(nthreads);
                std::vector<std::thread> th;
                th.reserve(nthreads);
                for (int t = 0; t < nthreads; t++) {
                    th.emplace_back([&, t]() {
                        cur_block = &blk;
                        lane = t & 31;
                        warp = t >> 5;
                        ::blockIdx = dim3(bx, by, bz);
                        ::blockDim = block;
                        ::gridDim = grid;
                        ::threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        kernel(args...);
                    });
                }
                for (auto &x : th) x.join();
            }
}
}  // namespace emu
