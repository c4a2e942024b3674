// ubench_mma.cu -- issue rate of the exact integer MMA  mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32  (and, for
// scale, of IDP.4A and the f16 HMMA m16n8k16) on B200, in warp-instructions per clock per SM.  One IMMA is 4096
// multiply-accumulates; 32 lanes of IDP.4A are 128.  Also times IMMA mixed with the ALU / LSU work a separable
// convolution needs around it (PRMT packing, LDS operand loads).  Not part of the library:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_mma.cu -o tools/ubench_mma_bin && tools/ubench_mma_bin
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

__device__ __forceinline__ void imma(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void imma16(int (&c)[4], const unsigned (&a)[2], unsigned b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b));
}
__device__ __forceinline__ void hmma(unsigned (&c)[2], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};\n"
                 : "+r"(c[0]), "+r"(c[1])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int OP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(unsigned *out, unsigned seed, long long *clk) {
    __shared__ unsigned sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * (i + 1);
    int c[ILP][4];
    unsigned h[ILP][2];
    unsigned a[4] = {seed | 1u, seed * 3u, seed * 5u, seed * 7u};
    unsigned b[ILP][2];
    unsigned x = seed + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        b[i][0] = seed + i; b[i][1] = seed * 9u + i;
        h[i][0] = h[i][1] = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) c[i][j] = i + j;
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) imma(c[i], a, b[i]);
            if (OP == 1) { c[i][0] = __dp4a((unsigned)c[i][0], a[0], (unsigned)c[i][1]); }
            if (OP == 2) hmma(h[i], a, b[i]);
            if (OP == 3) { const unsigned a2[2] = {a[0], a[1]}; imma16(c[i], a2, b[i][0]); }
            if (OP == 4) {             // IMMA + 2 PRMT (packing) + 1 LDS (operand) per MMA
                imma(c[i], a, b[i]);
                b[i][0] = __byte_perm(b[i][0], x, 0x6420);
                b[i][1] = __byte_perm(b[i][1], x, 0x7531);
                x += sm[(x + i) & 4095];
            }
            if (OP == 5) {             // IMMA + 4 IDP.4A: can the dot-product pipe run beside the tensor pipe?
                imma(c[i], a, b[i]);
#pragma unroll
                for (int j = 0; j < 4; j++) h[i][j & 1] = __dp4a(h[i][j & 1], a[j], (unsigned)x);
            }
        }
    }
    long long t1 = clock64();
    unsigned r = x;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= c[i][0] ^ c[i][1] ^ c[i][2] ^ c[i][3] ^ h[i][0] ^ h[i][1] ^ b[i][0] ^ b[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP, int WARPS>
void run(const char *name, double macs_per_instr) {
    unsigned *out;
    long long *clk;
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = WARPS * 32, blocks = sms;       // one CTA per SM
    cudaMalloc(&out, blocks * threads * 4);
    cudaMalloc(&clk, blocks * 8);
    k<OP, WARPS><<<blocks, threads>>>(out, 12345u, clk);
    k<OP, WARPS><<<blocks, threads>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    long long h[2048];
    cudaMemcpy(h, clk, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h[i];
    avg /= blocks;
    const double wi = (double)WARPS * ITERS * ILP / avg;
    printf("%-34s %2d warps/SM  %7.3f instr/clk/SM  (%6.2f clk per instr per SMSP)  %9.0f MAC/clk/SM\n", name, WARPS, wi,
           4.0 / wi, wi * macs_per_instr);
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0, 4>("IMMA m16n8k32 u8", 4096);
    run<0, 8>("IMMA m16n8k32 u8", 4096);
    run<0, 16>("IMMA m16n8k32 u8", 4096);
    run<0, 32>("IMMA m16n8k32 u8", 4096);
    run<3, 16>("IMMA m16n8k16 u8", 2048);
    run<2, 16>("HMMA m16n8k16 f16", 2048);
    run<1, 16>("IDP.4A", 128);
    run<1, 32>("IDP.4A", 128);
    run<4, 16>("IMMA + 2 PRMT + LDS", 4096);
    run<5, 16>("IMMA + 4 IDP.4A", 4096 + 4 * 128);
    cudaError_t e = cudaGetLastError();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
