// ubench.cu -- instruction throughput micro-benchmark for the arithmetic the filter kernels
// lean on (IDP.4A / IDP.2A / IMAD / FFMA / PRMT / LOP3 / SHF ...), in warp-instructions per
// clock per SM.  Not part of the library; built and run by hand:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench.cu -o gpurun_out/ubench && gpurun_out/ubench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

// packed float32 pairs (sm_100: FADD2 / FFMA2)
__device__ __forceinline__ void pk_add(float &x, float &y, float g, float h) {
    unsigned long long a, b;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(x), "f"(y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(g), "f"(h));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
}
__device__ __forceinline__ void pk_fma(float &x, float &y, float g, float h) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(x), "f"(y));
    asm("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(g));
    asm("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(h));
    asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
}

template <int OP>
__global__ void __launch_bounds__(1024) k(unsigned *out, unsigned seed, long long *clk) {
    unsigned a[ILP], b = seed | 1u, c = seed * 3u + 7u;
    float f[ILP], g = __uint_as_float(0x3f800001u + (seed & 1u)), hh = __uint_as_float(0x3e800000u + (seed & 1u));
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed + i * 17u + threadIdx.x; f[i] = (float)(i + 1) + g; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) a[i] = __dp4a(a[i], b, c);
            if (OP == 1) a[i] = __dp2a_lo(a[i], b, c);
            if (OP == 2) a[i] = a[i] * b + c;                                  // IMAD
            if (OP == 3) f[i] = fmaf(f[i], g, hh);                             // FFMA
            if (OP == 4) a[i] = __byte_perm(a[i], b, 0x5410 + (c & 0x1111));   // PRMT
            if (OP == 5) a[i] = (a[i] & b) ^ c;                                // LOP3
            if (OP == 6) a[i] = __funnelshift_r(a[i], b, c);                   // SHF
            if (OP == 7) a[i] = a[i] + b + c;                                  // IADD3
            if (OP == 8) f[i] = __fadd_rn(f[i], g);                            // FADD
            if (OP == 9) f[i] = __fmul_rn(f[i], g);                            // FMUL
            if (OP == 10) { a[i] = __dp4a(a[i], b, c); f[i] = fmaf(f[i], g, hh); }      // IDP + FFMA
            if (OP == 11) { a[i] = __dp4a(a[i], b, c); a[(i + 1) % ILP] ^= b; }           // IDP + LOP3 (alu)
            if (OP == 12) { a[i] = a[i] * b + c; f[i] = fmaf(f[i], g, hh); }              // IMAD + FFMA
            if (OP == 13) a[i] = __umulhi(a[i], b);                            // IMAD.HI
            if (OP == 14) f[i] = __uint_as_float(a[i] & 0xff) + f[i];          // (cheap int->float path)
            if (OP == 15) { a[i] = __dp2a_lo(a[i], b, c); f[i] = fmaf(f[i], g, hh); }     // IDP.2A + FFMA
            if (OP == 16) a[i] = __popc(a[i]) + c;                             // POPC
            if (OP == 17) f[i] = (float)a[i];                                  // I2F
            if (OP == 18) { a[i] = __dp4a(a[i], b, c); a[(i + 1) % ILP] = a[(i + 1) % ILP] * b + c; }   // IDP + IMAD
            if (OP == 19 && (i & 1) == 0) pk_add(f[i], f[i + 1], g, hh);                                 // FADD2
            if (OP == 20 && (i & 1) == 0) pk_fma(f[i], f[i + 1], g, hh);                                 // FFMA2
            if (OP == 21) { if ((i & 1) == 0) pk_add(f[i], f[i + 1], g, hh); a[i] = __funnelshift_r(a[i], b, c); }   // FADD2 per pair + SHF each
            if (OP == 22) { f[i] = __fadd_rn(f[i], g); a[i] = __funnelshift_r(a[i], b, c); }             // FADD + SHF
        }
    }
    long long t1 = clock64();
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= a[i] ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter) {
    unsigned *out;
    long long *clk;
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, blocks = sms * 2;       // 64 warps per SM
    cudaMalloc(&out, blocks * threads * 4);
    cudaMalloc(&clk, blocks * 8);
    k<OP><<<blocks, threads>>>(out, 12345u, clk);
    k<OP><<<blocks, threads>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    long long h[2048];
    cudaMemcpy(h, clk, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h[i];
    avg /= blocks;
    // warp-instructions per clock per SM: 64 warps/SM (2 CTAs x 32 warps) each issuing ITERS*ILP*per_iter
    const double wi = 64.0 * ITERS * ILP * per_iter / avg;
    printf("%-22s %8.2f warp-instr/clk/SM  = %7.1f lanes/clk/SM   (%.0f clk)\n", name, wi, wi * 32, avg);
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0>("IDP.4A", 1);
    run<1>("IDP.2A", 1);
    run<2>("IMAD", 1);
    run<3>("FFMA", 1);
    run<4>("PRMT", 1);
    run<5>("LOP3", 1);
    run<6>("SHF", 1);
    run<7>("IADD3", 1);
    run<8>("FADD", 1);
    run<9>("FMUL", 1);
    run<13>("IMAD.HI", 1);
    run<16>("POPC+IADD", 2);
    run<17>("I2F", 1);
    run<10>("IDP.4A + FFMA", 2);
    run<15>("IDP.2A + FFMA", 2);
    run<11>("IDP.4A + LOP3", 2);
    run<12>("IMAD + FFMA", 2);
    run<18>("IDP.4A + IMAD", 2);
    printf("packed float32 pairs: rates are PACKED instructions (2 lanes-worth each) per clock\n");
    run<19>("FADD2 (per pair)", 1);      // ILP/2 packed instructions per ILP slots: printed rate is x2 the instruction rate
    run<20>("FFMA2 (per pair)", 1);
    run<22>("FADD + SHF", 2);
    run<21>("FADD2/2 + SHF", 2);
    cudaError_t e = cudaGetLastError();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
