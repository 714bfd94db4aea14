// Integer-pipe micro-benchmarks that size the NTT butterfly budget on B200:
// issue rates of IMAD.WIDE / IMAD / IADD3 / mul.hi.u64 / DFMA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int OP> __global__ void __launch_bounds__(256) k(u64* out, u64 seed, int iters)
{
    u64 a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    u64 b = seed | 1;
    u32 x0 = (u32) a0, x1 = (u32) a1, x2 = (u32) a2, x3 = (u32) a3, y = (u32) b | 1;
    double d0 = (double) a0, d1 = (double) a1, d2 = (double) a2, d3 = (double) a3, e = 1.0000001;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int j = 0; j < 16; ++j)
        {
            if (OP == 0) { // IMAD.WIDE.U32 : 64-bit accumulate
                a0 = (u64) (u32) a0 * y + a0; a1 = (u64) (u32) a1 * y + a1;
                a2 = (u64) (u32) a2 * y + a2; a3 = (u64) (u32) a3 * y + a3;
            } else if (OP == 1) { // IMAD lo
                x0 = x0 * y + x0; x1 = x1 * y + x1; x2 = x2 * y + x2; x3 = x3 * y + x3;
            } else if (OP == 2) { // IADD3 / LOP
                x0 = (x0 + y) ^ x1; x1 = (x1 + y) ^ x2; x2 = (x2 + y) ^ x3; x3 = (x3 + y) ^ x0;
            } else if (OP == 3) { // mul.hi.u64
                a0 = __umul64hi(a0, b) + a0; a1 = __umul64hi(a1, b) + a1;
                a2 = __umul64hi(a2, b) + a2; a3 = __umul64hi(a3, b) + a3;
            } else if (OP == 4) { // DFMA
                d0 = d0 * e + d1; d1 = d1 * e + d2; d2 = d2 * e + d3; d3 = d3 * e + d0;
            } else if (OP == 5) { // Shoup modmul (exact mulhi)
                u64 p = b >> 3, w = seed >> 4, ws = ~seed;
                a0 = a0 * w - __umul64hi(a0, ws) * p; a1 = a1 * w - __umul64hi(a1, ws) * p;
                a2 = a2 * w - __umul64hi(a2, ws) * p; a3 = a3 * w - __umul64hi(a3, ws) * p;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + x0 + x1 + x2 + x3 + (u64) (d0 + d1 + d2 + d3);
}

template <int OP> void run(const char* name, double ops_per_iter)
{
    u64* out;
    int blocks = 148 * 8, iters = 4096;
    cudaMalloc(&out, blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, 0x9E3779B97F4A7C15ull, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 0x9E3779B97F4A7C15ull, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double total = (double) blocks * 256 * iters * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Gop/s  %6.1f ops/clk/SM @1.9GHz\n", name, ms, total / ms / 1e6,
           total / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main()
{
    run<0>("IMAD.WIDE.U32 (64b acc)", 64);
    run<1>("IMAD lo", 64);
    run<2>("IADD3+LOP3 pairs", 64);
    run<3>("mul.hi.u64 (+add)", 64);
    run<4>("DFMA", 64);
    run<5>("Shoup modmul 64b", 64);
    return 0;
}
