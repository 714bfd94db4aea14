// All-FP64 butterfly cost (register resident): cycles per warp-butterfly per SM sub-partition.
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
#define MAGIC 6755399441055744.0 /* 1.5 * 2^52 */

__device__ __forceinline__ void fbfly(double& X, double& Y, double w, double winv, double np)
{
    const double q = __fma_rn(Y, winv, MAGIC) - MAGIC;
    const double h = __dmul_rn(Y, w);
    const double l = __fma_rn(Y, w, -h);
    const double r = __fma_rn(q, np, h);
    const double T = __dadd_rn(r, l);
    const double x = X;
    X = __dadd_rn(x, T);
    Y = __dsub_rn(x, T);
}
__device__ __forceinline__ void fbfly_red(double& X, double& Y, double w, double winv, double np, double pinv)
{
    const double qx = __fma_rn(X, pinv, MAGIC) - MAGIC;
    const double x = __fma_rn(qx, np, X);
    const double q = __fma_rn(Y, winv, MAGIC) - MAGIC;
    const double h = __dmul_rn(Y, w);
    const double l = __fma_rn(Y, w, -h);
    const double r = __fma_rn(q, np, h);
    const double T = __dadd_rn(r, l);
    X = __dadd_rn(x, T);
    Y = __dsub_rn(x, T);
}
__device__ __forceinline__ void gsbfly(double& X, double& Y, double w, double winv, double np)
{
    const double s = __dadd_rn(X, Y), d = __dsub_rn(X, Y);
    const double q = __fma_rn(d, winv, MAGIC) - MAGIC;
    const double h = __dmul_rn(d, w);
    const double l = __fma_rn(d, w, -h);
    const double r = __fma_rn(q, np, h);
    X = s;
    Y = __dadd_rn(r, l);
}

template <int OP> __global__ void __launch_bounds__(256) k(double* out, u64 seed, int iters)
{
    const double p = (double) ((seed >> 15) | 1), np = -p, pinv = 1.0 / p;
    const double w = (double) (seed >> 16), winv = w / p;
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (double) ((seed * (threadIdx.x + i + 1)) >> 15);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            if (OP == 0) { fbfly(v[0], v[1], w, winv, np); fbfly(v[2], v[3], w, winv, np); fbfly(v[4], v[5], w, winv, np); fbfly(v[6], v[7], w, winv, np); }
            if (OP == 1) { fbfly_red(v[0], v[1], w, winv, np, pinv); fbfly_red(v[2], v[3], w, winv, np, pinv); fbfly_red(v[4], v[5], w, winv, np, pinv); fbfly_red(v[6], v[7], w, winv, np, pinv); }
            if (OP == 2) { gsbfly(v[0], v[1], w, winv, np); gsbfly(v[2], v[3], w, winv, np); gsbfly(v[4], v[5], w, winv, np); gsbfly(v[6], v[7], w, winv, np); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name)
{
    double* out;
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
    double warp_ops = (double) blocks * 8 * iters * 16;
    double cyc = ms * 1e-3 * 1.965e9 * 148 * 4 / warp_ops;
    printf("%-40s %8.3f ms  %6.2f SMSP-cycles per warp-butterfly @1.965GHz\n", name, ms, cyc);
    cudaFree(out);
}
int main()
{
    run<0>("fp64 CT butterfly (8 ops)");
    run<1>("fp64 CT butterfly + X reduce (11 ops)");
    run<2>("fp64 GS butterfly (8 ops)");
    return 0;
}
