// Pure-ALU cost of the butterfly variants (register resident, no memory):
// cycles per warp-butterfly per SM sub-partition at full occupancy.
#include <cstdio>
#include "../heongpu_b200/csrc/ntt_core.cuh"
using namespace heon;

template <int OP> __global__ void __launch_bounds__(256) k(u64* out, u64 seed, int iters)
{
    const u64 p = (seed >> 8) | 1, w = seed >> 9, ws = ~seed;
    const BflyConst bc{p, 2 * p, 4 * p, 0 - p};
    const TwPair tw{w, ws};
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + i + 1);
    unsigned a0 = (unsigned) v[0], a1 = (unsigned) v[1], a2 = (unsigned) v[2], a3 = (unsigned) v[3], y = (unsigned) seed | 1;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            if (OP == 0) { ct_bfly<0>(v[0], v[1], tw, bc); ct_bfly<0>(v[2], v[3], tw, bc); ct_bfly<0>(v[4], v[5], tw, bc); ct_bfly<0>(v[6], v[7], tw, bc); }
            if (OP == 1) { ct_bfly<1>(v[0], v[1], tw, bc); ct_bfly<1>(v[2], v[3], tw, bc); ct_bfly<1>(v[4], v[5], tw, bc); ct_bfly<1>(v[6], v[7], tw, bc); }
            if (OP == 2) { ct_bfly<2>(v[0], v[1], tw, bc); ct_bfly<2>(v[2], v[3], tw, bc); ct_bfly<2>(v[4], v[5], tw, bc); ct_bfly<2>(v[6], v[7], tw, bc); }
            if (OP == 3) { gs_bfly<1>(v[0], v[1], tw, bc); gs_bfly<1>(v[2], v[3], tw, bc); gs_bfly<1>(v[4], v[5], tw, bc); gs_bfly<1>(v[6], v[7], tw, bc); }
            if (OP == 6) { ct_bfly<3>(v[0], v[1], tw, bc); ct_bfly<3>(v[2], v[3], tw, bc); ct_bfly<3>(v[4], v[5], tw, bc); ct_bfly<3>(v[6], v[7], tw, bc); }
            if (OP == 7) { ct_bfly<4>(v[0], v[1], tw, bc); ct_bfly<4>(v[2], v[3], tw, bc); ct_bfly<4>(v[4], v[5], tw, bc); ct_bfly<4>(v[6], v[7], tw, bc); }
            if (OP == 4) { a0 = __umulhi(a0, y) + a1; a1 = __umulhi(a1, y) + a2; a2 = __umulhi(a2, y) + a3; a3 = __umulhi(a3, y) + a0; }
            if (OP == 5) { v[0] = shoup_lazy_ptx(v[0], w, ws, bc.np); v[1] = shoup_lazy_ptx(v[1], w, ws, bc.np); v[2] = shoup_lazy_ptx(v[2], w, ws, bc.np); v[3] = shoup_lazy_ptx(v[3], w, ws, bc.np); }
        }
    }
    u64 s = a0 + a1 + a2 + a3;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name)
{
    u64* out;
    int blocks = 148 * 8, iters = 2048;
    cudaMalloc(&out, blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, 0x9E3779B97F4A7C15ull, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 0x9E3779B97F4A7C15ull, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_ops = (double) blocks * 8 * iters * 16; // 16 ops per thread-iteration, per warp
    double cyc = ms * 1e-3 * 1.965e9 * 148 * 4 / warp_ops;
    printf("%-34s %8.3f ms  %6.2f SMSP-cycles per warp-op @1.965GHz\n", name, ms, cyc);
    cudaFree(out);
}

int main()
{
    run<0>("ct_bfly VAR0 (Harvey exact)");
    run<1>("ct_bfly VAR1 (csub 4p, ptx)");
    run<2>("ct_bfly VAR2 (no csub, ptx)");
    run<3>("gs_bfly GVAR1 (ptx)");
    run<4>("mul.hi.u32 + add");
    run<5>("shoup_lazy_ptx only");
    run<6>("ct_bfly VAR3 (fp64 quotient, nc)");
    run<7>("ct_bfly VAR4 (fp64 quotient, csub)");
    return 0;
}
