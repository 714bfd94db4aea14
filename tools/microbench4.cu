// Round-2 microbenchmarks (register resident, no memory traffic): SM sub-partition cycles per
// warp-operation for the alternatives weighed in DESIGN.md section 4:
//   * the FP64 butterfly with the quotient rounded by FRND (cvt.rni.f64.f64) instead of the
//     magic-constant add/sub pair (does the rounding leave the FP64 pipe?),
//   * canonicalisation of a lazy FP64 word: all-FP64 (fp_canon) vs magic-convert + integer fix-up,
//   * the key-switch inner product per (digit, coefficient): 128-bit integer MAC vs FP64 mulmod MAC,
//   * the same with an FP64 butterfly stream running in the same warp (do the pipes overlap?).
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
#define MAGIC 6755399441055744.0 /* 1.5 * 2^52 */

__device__ __forceinline__ double mulmod_magic(double y, double w, double winv, double np)
{
    const double q = __fma_rn(y, winv, MAGIC) - MAGIC;
    const double h = __dmul_rn(y, w);
    const double l = __fma_rn(y, w, -h);
    const double r = __fma_rn(q, np, h);
    return __dadd_rn(r, l);
}
__device__ __forceinline__ double mulmod_frnd(double y, double w, double winv, double np)
{
    double q;
    const double t = __dmul_rn(y, winv);
    asm("cvt.rni.f64.f64 %0, %1;" : "=d"(q) : "d"(t));
    const double h = __dmul_rn(y, w);
    const double l = __fma_rn(y, w, -h);
    const double r = __fma_rn(q, np, h);
    return __dadd_rn(r, l);
}
template <int FR> __device__ __forceinline__ void bfly(double& X, double& Y, double w, double winv, double np)
{
    const double T = FR ? mulmod_frnd(Y, w, winv, np) : mulmod_magic(Y, w, winv, np);
    const double x = X;
    X = __dadd_rn(x, T);
    Y = __dsub_rn(x, T);
}
__device__ __forceinline__ u64 canon_fp(double v, double pinv, double np, double dp)
{
    const double q = __fma_rn(v, pinv, MAGIC) - MAGIC;
    double r = __fma_rn(q, np, v);
    if (r < 0.0)
        r = __dadd_rn(r, dp);
    return (u64) __double_as_longlong(__dadd_rn(r, 4503599627370496.0)) & 0x000FFFFFFFFFFFFFull;
}
// reduce on the FP64 pipe (3), then convert with one magic add and fix the sign with integer ops
__device__ __forceinline__ u64 canon_mixed(double v, double pinv, double np, u64 p)
{
    const double q = __fma_rn(v, pinv, MAGIC) - MAGIC;
    const double r = __fma_rn(q, np, v); // [-p/2, p/2]
    const long long i = (__double_as_longlong(__dadd_rn(r, MAGIC)) << 12) >> 12; // two's complement 52-bit
    return (u64) (i < 0 ? i + (long long) p : i);
}
__device__ __forceinline__ void mac128(u64& lo, u64& hi, u64 a, u64 b)
{
    const u64 pl = a * b, ph = __umul64hi(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

template <int OP> __global__ void __launch_bounds__(256) k(double* out, u64 seed, int iters)
{
    const u64 pi = (seed >> 15) | 1;
    const double p = (double) pi, np = -p, pinv = 1.0 / p;
    const double w = (double) (seed >> 16), winv = w / p;
    double v[8];
    u64 a[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        v[i] = (double) ((seed * (threadIdx.x + i + 1)) >> 15);
        a[i] = seed * (threadIdx.x + 3 * i + 7);
    }
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            if (OP == 0 || OP == 1 || OP == 6 || OP == 7)
            {
                bfly<OP == 1>(v[0], v[1], w, winv, np);
                bfly<OP == 1>(v[2], v[3], w, winv, np);
                bfly<OP == 1>(v[4], v[5], w, winv, np);
                bfly<OP == 1>(v[6], v[7], w, winv, np);
            }
            if (OP == 2)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[i] = (double) (long long) canon_fp(v[i] + v[i + 4], pinv, np, p);
            if (OP == 3)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[i] = (double) (long long) canon_mixed(v[i] + v[i + 4], pinv, np, pi);
            if (OP == 4 || OP == 6) // integer MAC: 4 coefficient-digit pairs, two key components each
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    mac128(acc[2 * i], acc[2 * i + 1], a[i], a[i + 4]);
                    a[i] += acc[2 * i + 1];
                }
            if (OP == 5 || OP == 7) // FP64 MAC: x reduced, xinv, then k converted + mulmod + add
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    const double kd = __longlong_as_double(0x4330000000000000ll | (long long) (a[i] & 0xFFFFFFFFFFFFFull)) - 4503599627370496.0;
                    v[i] = __dadd_rn(v[i], mulmod_magic(kd, w, winv, np));
                    a[i] += 0x9E3779B9ull;
                }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += v[i] + (double) acc[i] + (double) a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name, double units)
{
    double* out;
    int blocks = 148 * 8, iters = 2048;
    cudaMalloc(&out, blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, 0x9E3779B97F4A7C15ull, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 0x9E3779B97F4A7C15ull, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_ops = (double) blocks * 8 * iters * 4 * units;
    double cyc = ms * 1e-3 * 1.965e9 * 148 * 4 / warp_ops;
    printf("%-58s %8.3f ms  %6.2f SMSP-cycles per warp-op @1.965GHz\n", name, ms, cyc);
    cudaFree(out);
}
int main()
{
    run<0>("fp64 CT butterfly, magic rounding (8 FP64)", 4);
    run<1>("fp64 CT butterfly, FRND rounding (7 FP64 + FRND)", 4);
    run<2>("canonicalise, all FP64 (fp_canon) [+1 add]", 4);
    run<3>("canonicalise, FP64 reduce + integer fix-up [+1 add]", 4);
    run<4>("128-bit integer MAC (one product)", 4);
    run<5>("FP64 MAC (convert + mulmod + add, one product)", 4);
    run<6>("butterfly + integer MAC in the same warp (per pair)", 4);
    run<7>("butterfly + FP64 MAC in the same warp (per pair)", 4);
    return 0;
}
