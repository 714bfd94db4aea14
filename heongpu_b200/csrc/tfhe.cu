// TFHE gate bootstrapping (BASELINE config 5): NAND / AND / OR / NOR / XOR / XNOR /
// NOT / MUX on batches of LWE samples, with the reference's parameter set (tfhe/context.cu:23-56: n = 512,
// N = 1024, k = 1, l = 2, Bg = 2^10, key-switch base 4 x 8 digits, the 60-bit NTT prime 1152921504606877697).
//
// reference: src/lib/host/tfhe/operator.cu:24-290 (gate pre-computation, bootstrapping, key switching),
// src/lib/kernel/bootstrapping.cu:378-660 (gate kernels), :662-674 (modulus switch), :875-1312 (blind rotation
// as 1 + 2*511 launches through two global buffers), :1314-1349 (sample extraction), :1351-1437 (key switch),
// src/lib/kernel/small_ntt.cu (1024-point NTT in shared memory, two coefficients per thread),
// src/lib/kernel/keygeneration.cu:1032-1440 + src/lib/host/tfhe/keygenerator.cu (keys),
// src/lib/kernel/encryption.cu:280-327, decryption.cu:441-478 (LWE encrypt / decrypt).
//
// B200 design: the whole blind rotation of one LWE sample is ONE CTA of one launch.  The accumulator (2 x 1024
// int32) lives in shared memory for all 512 steps; per step the four digit polynomials are transformed by 64
// threads each with sixteen coefficients in registers (three register rounds 4 + 4 + 2 stages, two exchanges
// through a swizzled shared buffer, hand-written PTX Shoup butterflies: the 60-bit prime has no FP64 form),
// multiplied with the bootstrapping key (33.5 MB, resident in L2) into 128-bit lazy sums, and the two inverse
// transforms add straight into the accumulator.  Nothing but the key is read from global memory inside the
// loop, and there is one launch instead of 1023.  Steps whose rotation amount is 0 are skipped (they add 0).
// Every integer the reference computes is computed here (same prime, same transform order, canonical sums), so
// outputs are bit-identical to the reference kernels' (tests/test_gpu_tfhe.py).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "modarith.cuh"
#include "ntt_core.cuh"
#include "tma.cuh"
#include "ops.hpp"

namespace heon {

namespace {

constexpr int TN = 1024; // ring degree N
constexpr int TLOGN = 10;

__host__ __device__ __forceinline__ u64 t_mix64(u64 z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 t_rnd64(u64 seed, u64 stream, u64 ctr)
{
    return t_mix64(t_mix64(seed ^ (stream * 0xD1342543DE82EF95ull)) + ctr * 0x9E3779B97F4A7C15ull);
}
// standard normal by Box-Muller from two counter-based draws
__device__ __forceinline__ double t_normal(u64 seed, u64 stream, u64 ctr)
{
    const double u1 = ((double) (t_rnd64(seed, stream, 2 * ctr) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = (double) (t_rnd64(seed, stream, 2 * ctr + 1) >> 11) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
// reference: double_to_torus32 (tfhe/encryptor.cu:100-107, keygeneration.cu:1134-1137)
__host__ __device__ __forceinline__ int t_double_to_torus32(double x)
{
    const double frac = x - trunc(x);
    return (int) (unsigned) (long long) floor(frac * 4294967296.0 + 0.5);
}

// torus_modulus_switch_log (bootstrapping.cu:662-674): round a 32-bit torus element to Z_{2N}
__device__ __forceinline__ int t_mod_switch(int x)
{
    const unsigned long long r = ((unsigned long long) (unsigned) x << 32) + (1ull << (62 - TLOGN));
    return (int) (r >> (63 - TLOGN));
}

// shared-buffer index of coefficient c: the 16-byte chunk inside a 128-byte line is XORed with a 3-bit function of
// the line index that separates (i) the eight consecutive lines a quarter-warp touches with 128-bit accesses
// (c = 16j + 2m) and (ii) the four lines 64 words apart a half-warp touches in the middle round (c = 64B + r + 4k);
// the linear patterns (c = j + 64k, c = tid + 128m) stay inside one line per half-warp.  Measured before this
// function had the second term: 16 wavefronts per 128-bit access instead of 4.
__device__ __forceinline__ int t_sw(int c) { return c ^ ((((c >> 4) ^ (c >> 6)) & 7) << 1); }

// coefficient c of X^a * acc (a in [0, 2N)), acc a negacyclic polynomial of int32
__device__ __forceinline__ int t_rot(const int* acc, int c, int a)
{
    if (a < TN)
        return (c < a) ? -acc[TN - a + c] : acc[c - a];
    const int am = a - TN;
    return (c < am) ? acc[TN - am + c] : -acc[c - am];
}

// forward transform of one polynomial by 64 threads (j = 0..63), in: v[k] = coefficient j + 64k (canonical),
// out: v[k] = NTT word 16j + k (lazy, below 8p).  buf: this polynomial's 1024-word exchange buffer.
__device__ __forceinline__ void t_ntt_fwd(u64 (&v)[16], u64* buf, int j, int bar_id, const TwPair* __restrict__ tw,
                                          const BflyConst& bc)
{
    ct_round_a<1>(v, tw, 0, 0, bc); // stages 0..3: distances 512..64
#pragma unroll
    for (int k = 0; k < 16; ++k)
        buf[t_sw(j + 64 * k)] = v[k];
    named_bar_sync(bar_id, 64);
    const int B = j >> 2, r = j & 3;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        v[k] = buf[t_sw(64 * B + r + 4 * k)];
    ct_round_a<1>(v, tw, 4, B, bc); // stages 4..7: distances 32..4 inside 64-word block B
#pragma unroll
    for (int k = 0; k < 16; ++k)
        buf[t_sw(64 * B + r + 4 * k)] = v[k];
    named_bar_sync(bar_id, 64);
#pragma unroll
    for (int m = 0; m < 8; ++m)
    {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(buf + t_sw(16 * j + 2 * m));
        v[2 * m] = t.x;
        v[2 * m + 1] = t.y;
    }
    stage16<1, false, 1, 1>(v, tw + 256 + 4 * j, bc); // stage 8: distance 2
    stage16<0, false, 1, 1>(v, tw + 512 + 8 * j, bc); // stage 9: distance 1
}

// inverse transform by 64 threads, in: v[k] = NTT word 16j + k (below 4p), out: v[k] = coefficient j + 64k,
// canonical, multiplied by N^-1 (folded into the last stage: ninv = N^-1, wninv = w_1 * N^-1)
__device__ __forceinline__ void t_ntt_inv(u64 (&v)[16], u64* buf, int j, int bar_id, const TwPair* __restrict__ tw,
                                          const TwPair& ninv, const TwPair& wninv, const BflyConst& bc)
{
    stage16<0, true, 1, 1>(v, tw + 512 + 8 * j, bc);
    stage16<1, true, 1, 1>(v, tw + 256 + 4 * j, bc);
#pragma unroll
    for (int m = 0; m < 8; ++m)
    {
        ulonglong2 t;
        t.x = v[2 * m];
        t.y = v[2 * m + 1];
        *reinterpret_cast<ulonglong2*>(buf + t_sw(16 * j + 2 * m)) = t;
    }
    named_bar_sync(bar_id, 64);
    const int B = j >> 2, r = j & 3;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        v[k] = buf[t_sw(64 * B + r + 4 * k)];
    stage16<0, true, 1, 1>(v, tw + 128 + 8 * B, bc);
    stage16<1, true, 1, 1>(v, tw + 64 + 4 * B, bc);
    stage16<2, true, 1, 1>(v, tw + 32 + 2 * B, bc);
    stage16<3, true, 1, 1>(v, tw + 16 + B, bc);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        buf[t_sw(64 * B + r + 4 * k)] = v[k];
    named_bar_sync(bar_id, 64);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        v[k] = buf[t_sw(j + 64 * k)];
    stage16<0, true, 1, 1>(v, tw + 8, bc);
    stage16<1, true, 1, 1>(v, tw + 4, bc);
    stage16<2, true, 1, 1>(v, tw + 2, bc);
    // last stage with N^-1: X' = (X + Y) * N^-1, Y' = (X - Y) * w_1 * N^-1, canonical
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        u64 s = v[k], d = v[k + 8];
        addsub_ptx(s, d, v[k + 8], bc.p4); // s = X + Y, d = X + 4p - Y
        v[k] = csub(csub(shoup_mul_lazy3(s, ninv.w, ninv.ws, bc.p), bc.p2), bc.p);
        v[k + 8] = csub(csub(shoup_mul_lazy3(d, wninv.w, wninv.ws, bc.p), bc.p2), bc.p);
    }
}

// ---------------------------------------------------------------------------
// gates: out = enc + s1*in1 + s2*in2 on (a, b)   (bootstrapping.cu:378-660)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_tfhe_linear(const int* __restrict__ a1, const int* __restrict__ b1, const int* __restrict__ a2,
                  const int* __restrict__ b2, int* __restrict__ oa, int* __restrict__ ob, int enc, int s1, int s2, int n,
                  int shape)
{
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    const long long total = (long long) shape * n;
    if (idx < total)
    {
        unsigned v = (unsigned) s1 * (unsigned) a1[idx];
        if (a2)
            v += (unsigned) s2 * (unsigned) a2[idx];
        oa[idx] = (int) v;
    }
    if (idx < shape)
    {
        unsigned v = (unsigned) enc + (unsigned) s1 * (unsigned) b1[idx];
        if (b2)
            v += (unsigned) s2 * (unsigned) b2[idx];
        ob[idx] = (int) v;
    }
}

// ---------------------------------------------------------------------------
// blind rotation + sample extraction: one CTA per LWE sample
// ---------------------------------------------------------------------------
struct TfheDev {
    PrimeConst pc;
    TwPair ninv, wninv;
    int mu; // encode_to_torus32(1, 8)
    int bk_offset, bk_half, bk_mask, bk_bit;
    int n;
};

constexpr int kBrThreads = 128;
constexpr int kBrSmem = 2 * TN * 4 + 4 * TN * 8;

// 128 threads = two groups of 64.  Per step, group g transforms the two digit polynomials of accumulator
// component g (one after the other: both come from the same rotated difference), all threads form the inner
// products, and group g runs the inverse transform of output component g: six transforms of equal cost on two
// groups in three rounds, nobody idles.  The inner-product sums overwrite the first two digit buffers (every
// thread owns its coefficient columns in that phase), so a CTA needs 40 KiB and five fit on an SM.
__global__ void __launch_bounds__(kBrThreads, 4)
    k_tfhe_blind_rotate(const int* __restrict__ in_a, const int* __restrict__ in_b, int* __restrict__ out_a,
                        int* __restrict__ out_b, const u64* __restrict__ bk, const TwPair* __restrict__ fwd,
                        const TwPair* __restrict__ inv, const TfheDev P)
{
    extern __shared__ __align__(16) unsigned char t_smem[];
    int* acc = reinterpret_cast<int*>(t_smem); // [2][1024]
    u64* work = reinterpret_cast<u64*>(t_smem + 2 * TN * 4); // [4][1024] transformed digits; [0..1] reused for the sums
    const int tid = threadIdx.x;
    const long long s = blockIdx.x;
    const BflyConst bc = make_bc(P.pc);
    const int g = tid >> 6, j = tid & 63;

    // accumulator = (0, X^{-b~} * testvector), testvector = mu at every coefficient (bootstrapping.cu:912-937)
    {
        const int bN = 2 * TN - t_mod_switch(in_b[s]);
        for (int c = tid; c < TN; c += kBrThreads)
        {
            int t;
            if (bN < TN)
                t = (c < bN) ? -P.mu : P.mu;
            else
                t = (c < bN - TN) ? P.mu : -P.mu;
            acc[c] = 0;
            acc[TN + c] = t;
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < P.n; ++i)
    {
        const int a = t_mod_switch(in_a[s * P.n + i]);
        if (a == 0)
            continue; // (X^0 - 1) * acc = 0: every digit is 0
        u64 v[16];
        {
            // digit z of (X^a - 1) * acc_g at the coefficients j + 64k, as residues
            const int* ay = acc + g * TN;
#pragma unroll 1
            for (int z = 0; z < 2; ++z)
            {
                const int shift = 32 - P.bk_bit * (z + 1);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                {
                    const int c = j + 64 * k;
                    const unsigned diff = (unsigned) t_rot(ay, c, a) - (unsigned) ay[c];
                    const int dg = (int) (((diff + (unsigned) P.bk_offset) >> shift) & (unsigned) P.bk_mask) - P.bk_half;
                    v[k] = dg < 0 ? P.pc.p + (u64) (long long) dg : (u64) dg;
                }
                u64* wq = work + (g * 2 + z) * TN;
                t_ntt_fwd(v, wq, j, 1 + g, fwd, bc);
#pragma unroll
                for (int m = 0; m < 8; ++m)
                {
                    ulonglong2 t;
                    t.x = csub(v[2 * m], bc.p4);
                    t.y = csub(v[2 * m + 1], bc.p4);
                    *reinterpret_cast<ulonglong2*>(wq + t_sw(16 * j + 2 * m)) = t;
                }
            }
        }
        __syncthreads();
        // out_jj[c] = sum over (y, z) of digit word * bk[i][y][z][jj][c]: canonical, as the reference's sum of
        // canonical products (bootstrapping.cu:1127-1139, 1263-1283)
        {
            const u64* bki = bk + (size_t) i * 8 * TN;
#pragma unroll 1
            for (int m = 0; m < TN / kBrThreads; ++m)
            {
                const int c = tid + kBrThreads * m, cs = t_sw(c);
                u64 l0 = 0, h0 = 0, l1 = 0, h1 = 0;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
                {
                    const u64 x = work[qq * TN + cs];
                    mac128(l0, h0, x, __ldg(bki + (qq * 2 + 0) * TN + c));
                    mac128(l1, h1, x, __ldg(bki + (qq * 2 + 1) * TN + c));
                }
                work[cs] = reduce_u128(l0, h0, P.pc); // column cs of every buffer belongs to this thread here
                work[TN + cs] = reduce_u128(l1, h1, P.pc);
            }
        }
        __syncthreads();
        {
            u64* ob = work + g * TN;
#pragma unroll
            for (int m = 0; m < 8; ++m)
            {
                const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(ob + t_sw(16 * j + 2 * m));
                v[2 * m] = t.x;
                v[2 * m + 1] = t.y;
            }
            t_ntt_inv(v, ob, j, 1 + g, inv, P.ninv, P.wninv, bc);
            const u64 thr = P.pc.p >> 1;
            int* aj = acc + g * TN;
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                const int c = j + 64 * k;
                const int add = (v[k] >= thr) ? (int) (long long) (v[k] - P.pc.p) : (int) (long long) v[k];
                aj[c] = (int) ((unsigned) aj[c] + (unsigned) add);
            }
        }
        __syncthreads();
    }
    // sample extraction at index 0 (bootstrapping.cu:1314-1349)
    for (int c = tid; c < TN; c += kBrThreads)
        out_a[s * TN + c] = (c < 1) ? acc[c] : -acc[TN - c];
    if (tid == 0)
        out_b[s] = acc[TN];
}

// ---------------------------------------------------------------------------
// key switching N -> n (bootstrapping.cu:1351-1437): out = (0, b) - sum over (i, digit i2) of row(i, i2, value).
// One CTA takes KS_S samples and one thread one output coefficient: the three candidate rows of a (i, i2) are
// loaded once for all KS_S samples (the reference and a one-sample CTA stream 12 MB of key per sample from L2).
// The 2-bit digits of every input coefficient are packed into 16 bits in shared memory first.
// ---------------------------------------------------------------------------
constexpr int KS_S = 8;
__global__ void __launch_bounds__(512)
    k_tfhe_keyswitch(const int* __restrict__ in_a, const int* __restrict__ in_b, int* __restrict__ out_a,
                     int* __restrict__ out_b, const int* __restrict__ ks_a, const int* __restrict__ ks_b, int base_bit,
                     int length, int n, int Nk, int shape)
{
    __shared__ unsigned short dgs[KS_S][TN];
    const int tid = threadIdx.x;
    const long long s0 = (long long) blockIdx.x * KS_S;
    const int ns = (int) min((long long) KS_S, shape - s0);
    const int mask = (1 << base_bit) - 1;
    const unsigned prec = 1u << (32 - (1 + base_bit * length));
    for (int e = tid; e < KS_S * Nk; e += blockDim.x)
    {
        const int sidx = e / Nk, c = e - sidx * Nk;
        unsigned short pk = 0;
        if (sidx < ns)
        {
            const unsigned av = (unsigned) in_a[(s0 + sidx) * Nk + c] + prec;
            for (int i2 = 0; i2 < length; ++i2)
                pk |= (unsigned short) (((av >> (32 - (i2 + 1) * base_bit)) & (unsigned) mask) << (2 * i2));
        }
        dgs[sidx][c] = pk;
    }
    __syncthreads();
    unsigned acc[KS_S];
#pragma unroll
    for (int q = 0; q < KS_S; ++q)
        acc[q] = 0;
    unsigned accb = (tid < ns) ? (unsigned) in_b[s0 + tid] : 0u;
    if (base_bit == 2 && length == 8 && tid < n)
    {
        // the 24 key words of coefficient i + 1 are requested before coefficient i is consumed: the kernel is bound by
        // the latency of these L2 reads (ncu: long_scoreboard 8.6 per issue without the prefetch), not by bandwidth
        unsigned cur[24], nxt[24];
        const int* rowp = ks_a + tid;
#pragma unroll
        for (int e = 0; e < 24; ++e)
            cur[e] = (unsigned) __ldg(rowp + (size_t) e * n);
#pragma unroll 1
        for (int i = 0; i < Nk; ++i)
        {
            if (i + 1 < Nk)
            {
                const int* nrow = ks_a + ((size_t) (i + 1) * 24) * n + tid;
#pragma unroll
                for (int e = 0; e < 24; ++e)
                    nxt[e] = (unsigned) __ldg(nrow + (size_t) e * n);
            }
            unsigned short d[KS_S];
#pragma unroll
            for (int q = 0; q < KS_S; ++q)
                d[q] = dgs[q][i];
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2)
            {
                const unsigned r1 = cur[i2 * 3 + 0], r2 = cur[i2 * 3 + 1], r3 = cur[i2 * 3 + 2];
#pragma unroll
                for (int q = 0; q < KS_S; ++q)
                {
                    const unsigned dg = (d[q] >> (2 * i2)) & 3u;
                    const unsigned x = dg == 1 ? r1 : (dg == 2 ? r2 : r3);
                    acc[q] -= dg ? x : 0u;
                }
            }
            if (tid < ns)
            {
                const unsigned short dd = dgs[tid][i];
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2)
                {
                    const unsigned dg = (dd >> (2 * i2)) & 3u;
                    if (dg)
                        accb -= (unsigned) __ldg(ks_b + (size_t) i * 24 + i2 * 3 + (dg - 1));
                }
            }
#pragma unroll
            for (int e = 0; e < 24; ++e)
                cur[e] = nxt[e];
        }
    }
    else if (tid < n)
    {
        // general (base, length): one sample at a time
        for (int q = 0; q < ns; ++q)
            for (int i = 0; i < Nk; ++i)
            {
                const unsigned av = (unsigned) in_a[(s0 + q) * Nk + i] + prec;
                for (int i2 = 0; i2 < length; ++i2)
                {
                    const int dg = (int) ((av >> (32 - (i2 + 1) * base_bit)) & (unsigned) mask);
                    if (dg)
                    {
                        const size_t row = ((size_t) i * length + i2) * mask + (dg - 1);
                        acc[q] -= (unsigned) __ldg(ks_a + row * n + tid);
                        if (tid == q)
                            accb -= (unsigned) __ldg(ks_b + row);
                    }
                }
            }
    }
    if (tid < n)
    {
#pragma unroll
        for (int q = 0; q < KS_S; ++q)
            if (q < ns)
                out_a[(s0 + q) * n + tid] = (int) acc[q];
    }
    if (tid < ns)
        out_b[s0 + tid] = (int) accb;
}

// ---------------------------------------------------------------------------
// client side: keys, encryption, decryption
// ---------------------------------------------------------------------------
__global__ void k_tfhe_bits(int* out, int count, u64 seed, u64 stream)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count)
        out[idx] = (int) (t_rnd64(seed, stream, (u64) idx) & 1);
}

// one TGSW row (LWE bit i, component y, digit z): a uniform, b = a * s + e (negacyclic, exact through the NTT:
// |a * s| < 2^41), message s_i * 2^(32 - (z+1)*bit) added to the constant coefficient of component y, both
// polynomials stored in the NTT domain (keygeneration.cu:1153-1440).  64 threads per polynomial.
__global__ void __launch_bounds__(64)
    k_tfhe_bootkey_row(u64* __restrict__ bk, const int* __restrict__ lwe_key, const int* __restrict__ tlwe_key,
                       const TwPair* __restrict__ fwd, const TwPair* __restrict__ inv, const TfheDev P, double stdev,
                       u64 seed)
{
    __shared__ __align__(16) u64 buf[TN];
    const int j = threadIdx.x;
    const int row = blockIdx.x; // (i * 2 + y) * 2 + z
    const int z = row & 1, y = (row >> 1) & 1, i = row >> 2;
    const BflyConst bc = make_bc(P.pc);
    const u64 p = P.pc.p;
    const unsigned msg = (unsigned) lwe_key[i] << (32 - (z + 1) * P.bk_bit);
    u64 v[16], sk[16];
    int a[16];
    // NTT(s)
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sk[k] = (u64) tlwe_key[j + 64 * k];
    t_ntt_fwd(sk, buf, j, 1, fwd, bc);
    // NTT(a)
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
        a[k] = (int) (unsigned) t_rnd64(seed, 0x7F0 + row, (u64) (j + 64 * k));
        v[k] = a[k] < 0 ? p + (u64) (long long) a[k] : (u64) a[k];
    }
    named_bar_sync(1, 64);
    t_ntt_fwd(v, buf, j, 1, fwd, bc);
    // a * s: word 16j + k of both
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
        u64 lo = 0, hi = 0;
        mac128(lo, hi, csub(v[k], bc.p4), csub(sk[k], bc.p4));
        v[k] = reduce_u128(lo, hi, P.pc);
    }
    named_bar_sync(1, 64);
    t_ntt_inv(v, buf, j, 1, inv, P.ninv, P.wninv, bc);
    // rows: a' = a (+ msg at X^0 if y == 0), b = a*s + e (+ msg at X^0 if y == 1)
    u64 bw[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
        const int c = j + 64 * k;
        const int prod = (v[k] >= (p >> 1)) ? (int) (long long) (v[k] - p) : (int) (long long) v[k];
        const int e = t_double_to_torus32(stdev * t_normal(seed, 0x9E0 + row, (u64) c));
        unsigned b = (unsigned) prod + (unsigned) e;
        unsigned aa = (unsigned) a[k];
        if (c == 0)
        {
            if (y == 0)
                aa += msg;
            else
                b += msg;
        }
        v[k] = (int) aa < 0 ? p + (u64) (long long) (int) aa : (u64) (int) aa;
        bw[k] = (int) b < 0 ? p + (u64) (long long) (int) b : (u64) (int) b;
    }
    named_bar_sync(1, 64);
    t_ntt_fwd(v, buf, j, 1, fwd, bc);
    u64* o = bk + (size_t) row * 2 * TN;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        o[16 * j + k] = csub(csub(csub(v[k], bc.p4), bc.p2), bc.p);
    named_bar_sync(1, 64);
    t_ntt_fwd(bw, buf, j, 1, fwd, bc);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        o[TN + 16 * j + k] = csub(csub(csub(bw[k], bc.p4), bc.p2), bc.p);
}

// key-switch key: row (i, i2, v): a uniform in Z_{2^32}^n, b = <a, s_lwe> + s_tlwe[i] * (v+1) * 2^(32 - (i2+1)*bit) + e
__global__ void __launch_bounds__(256)
    k_tfhe_switchkey(int* __restrict__ ks_a, int* __restrict__ ks_b, const int* __restrict__ lwe_key,
                     const int* __restrict__ tlwe_key, int n, int base_bit, int length, double stdev, u64 seed)
{
    __shared__ unsigned red[8];
    const long long row = blockIdx.x;
    const int mask = (1 << base_bit) - 1;
    const int vdx = (int) (row % mask), i2 = (int) ((row / mask) % length), i = (int) (row / ((long long) mask * length));
    unsigned sum = 0;
    for (int t = threadIdx.x; t < n; t += 256)
    {
        const unsigned r = (unsigned) t_rnd64(seed, 0xA50, (u64) row * n + t);
        ks_a[row * n + t] = (int) r;
        sum += r * (unsigned) lwe_key[t];
    }
    for (int o = 16; o; o >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned tot = 0;
        for (int w = 0; w < 8; ++w)
            tot += red[w];
        const unsigned msg = (unsigned) tlwe_key[i] * ((unsigned) (vdx + 1) << (32 - (i2 + 1) * base_bit));
        const int e = t_double_to_torus32(stdev * t_normal(seed, 0xA51, (u64) row));
        ks_b[row] = (int) (tot + msg + (unsigned) e);
    }
}

// LWE encryption of torus32 messages: b = <a, s> + m + e (encryption.cu:280-327), one CTA per sample
__global__ void __launch_bounds__(256)
    k_tfhe_encrypt(int* __restrict__ out_a, int* __restrict__ out_b, const int* __restrict__ msg,
                   const int* __restrict__ lwe_key, int n, double stdev, u64 seed)
{
    __shared__ unsigned red[8];
    const long long s = blockIdx.x;
    unsigned sum = 0;
    for (int t = threadIdx.x; t < n; t += 256)
    {
        const unsigned r = (unsigned) t_rnd64(seed, 0xE00, (u64) s * n + t);
        out_a[s * n + t] = (int) r;
        sum += r * (unsigned) lwe_key[t];
    }
    for (int o = 16; o; o >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned tot = 0;
        for (int w = 0; w < 8; ++w)
            tot += red[w];
        const int e = t_double_to_torus32(stdev * t_normal(seed, 0xE01, (u64) s));
        out_b[s] = (int) (tot + (unsigned) msg[s] + (unsigned) e);
    }
}

// phase = b - <a, s> (decryption.cu:441-478)
__global__ void __launch_bounds__(256)
    k_tfhe_phase(const int* __restrict__ in_a, const int* __restrict__ in_b, int* __restrict__ phase,
                 const int* __restrict__ lwe_key, int n)
{
    __shared__ unsigned red[8];
    const long long s = blockIdx.x;
    unsigned sum = 0;
    for (int t = threadIdx.x; t < n; t += 256)
        sum += (unsigned) in_a[s * n + t] * (unsigned) lwe_key[t];
    for (int o = 16; o; o >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned tot = 0;
        for (int w = 0; w < 8; ++w)
            tot += red[w];
        phase[s] = (int) ((unsigned) in_b[s] - tot);
    }
}

void t_check_launch()
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("kernel launch: ") + cudaGetErrorString(e));
}

} // namespace

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct TfheContext {
    int device = 0;
    int n = 512, N = 1024, k = 1, bk_l = 2, bk_bg_bit = 10, ks_base_bit = 2, ks_length = 8;
    double ks_stdev, bk_stdev, max_stdev;
    u64 prime = 1152921504606877697ull, psi = 1689264667710614ull;
    TwPair *d_fwd = nullptr, *d_inv = nullptr;
    TfheDev dev;
};

static int tfhe_encode(unsigned mu, unsigned m_size) // encode_to_torus32 (tfhe/operator.cu:316-322)
{
    const unsigned long long interval = ((1ull << 63) / m_size) * 2;
    return (int) ((mu * interval) >> 32);
}

TfheContext* tfhe_create(int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        throw std::runtime_error("heon_tfhe_create: no such CUDA device (the TFHE path has no CPU fallback)");
    auto* c = new TfheContext();
    c->device = device;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    const double s2pi = std::sqrt(2.0 / M_PI);
    c->ks_stdev = (1.0 / 32768.0) * s2pi;
    c->bk_stdev = 9e-9 * s2pi;
    c->max_stdev = (1.0 / 64.0) * s2pi;
    const u64 p = c->prime;
    // tables in bit-reversed order (tfhe/context.cu:80-104)
    auto table = [&](u64 root) {
        std::vector<u64> pw(TN);
        pw[0] = 1;
        for (int j = 1; j < TN; ++j)
            pw[j] = mulmod(pw[j - 1], root, p);
        std::vector<TwPair> t(TN);
        for (int j = 0; j < TN; ++j)
        {
            int r = 0;
            for (int b = 0; b < TLOGN; ++b)
                r |= ((j >> b) & 1) << (TLOGN - 1 - b);
            t[j].w = pw[r];
            t[j].ws = shoup(pw[r], p);
        }
        return t;
    };
    const u64 psi_inv = invmod(c->psi, p);
    std::vector<TwPair> fwd = table(c->psi), inv = table(psi_inv);
    cudaMalloc(&c->d_fwd, sizeof(TwPair) * TN);
    cudaMalloc(&c->d_inv, sizeof(TwPair) * TN);
    cudaMemcpy(c->d_fwd, fwd.data(), sizeof(TwPair) * TN, cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_inv, inv.data(), sizeof(TwPair) * TN, cudaMemcpyHostToDevice);
    TfheDev& d = c->dev;
    memset(&d, 0, sizeof(d));
    d.pc.p = p;
    d.pc.inv64 = shoup(1, p);
    d.pc.r64 = (u64) ((((u128) 1) << 64) % p);
    d.pc.r64s = shoup(d.pc.r64, p);
    d.pc.bits = 61;
    d.pc.fin_shift = d.pc.bits - 25;
    d.pc.fin_m = (unsigned) ((((u128) 1) << (d.pc.bits + 31)) / p);
    d.pc.pinv = 1.0 / (double) p;
    const u64 ninv = invmod((u64) TN, p);
    d.ninv.w = ninv;
    d.ninv.ws = shoup(ninv, p);
    d.wninv.w = mulmod(inv[1].w, ninv, p);
    d.wninv.ws = shoup(d.wninv.w, p);
    d.mu = tfhe_encode(1, 8);
    d.bk_bit = c->bk_bg_bit;
    d.bk_half = (1 << c->bk_bg_bit) >> 1;
    d.bk_mask = (1 << c->bk_bg_bit) - 1;
    long long sum = 0;
    for (int i = 1; i <= c->bk_l; ++i)
        sum += 1ll << (32 - i * c->bk_bg_bit);
    d.bk_offset = (int) (sum * d.bk_half); // compute_offset (tfhe/context.cu:68-78)
    d.n = c->n;
    cudaFuncSetAttribute(k_tfhe_blind_rotate, cudaFuncAttributeMaxDynamicSharedMemorySize, kBrSmem);
    cudaSetDevice(prev);
    return c;
}

void tfhe_destroy(TfheContext* c)
{
    if (!c)
        return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaFree(c->d_fwd);
    cudaFree(c->d_inv);
    cudaSetDevice(prev);
    delete c;
}

int tfhe_device(const TfheContext* c) { return c->device; }

void tfhe_params(const TfheContext* c, int* out)
{
    out[0] = c->n;
    out[1] = c->N;
    out[2] = c->k;
    out[3] = c->bk_l;
    out[4] = c->bk_bg_bit;
    out[5] = c->ks_base_bit;
    out[6] = c->ks_length;
}

// gate codes: the linear part before the bootstrap, out = enc + s1*in1 + s2*in2
// (tfhe/operator.cu:24-196).  0 NAND, 1 AND, 2 NOR, 3 OR, 4 XNOR, 5 XOR, 6 AND with the first input negated,
// 7 NOT (no bootstrap), 8 MUX.
static void gate_coeffs(int gate, int& enc, int& s1, int& s2)
{
    const int e8 = tfhe_encode(1, 8), e4 = tfhe_encode(1, 4);
    switch (gate)
    {
    case 0: enc = e8, s1 = -1, s2 = -1; break;
    case 1: enc = -e8, s1 = 1, s2 = 1; break;
    case 2: enc = -e8, s1 = -1, s2 = -1; break;
    case 3: enc = e8, s1 = 1, s2 = 1; break;
    case 4: enc = -e4, s1 = -2, s2 = -2; break;
    case 5: enc = e4, s1 = 2, s2 = 2; break;
    case 6: enc = -e8, s1 = -1, s2 = 1; break;
    default: throw std::invalid_argument("unknown gate");
    }
}

void tfhe_gate_linear(const TfheContext& c, int gate, const int* a1, const int* b1, const int* a2, const int* b2, int* oa,
                      int* ob, int n, int shape, cudaStream_t st)
{
    if (shape < 1 || n < 1)
        throw std::invalid_argument("empty ciphertext");
    const long long total = (long long) shape * n;
    LaunchScope scope(KC_ELEMENTWISE, st);
    if (gate == 7)
        k_tfhe_linear<<<(unsigned) ((total + 255) / 256), 256, 0, st>>>(a1, b1, nullptr, nullptr, oa, ob, 0, -1, 0, n, shape);
    else
    {
        int enc, s1, s2;
        gate_coeffs(gate, enc, s1, s2);
        k_tfhe_linear<<<(unsigned) ((total + 255) / 256), 256, 0, st>>>(a1, b1, a2, b2, oa, ob, enc, s1, s2, n, shape);
    }
    t_check_launch();
}

// in: LWE samples under the n-bit key ([shape][n], [shape]); out: LWE samples under the extracted N-bit key
void tfhe_bootstrap(const TfheContext& c, const int* in_a, const int* in_b, int* out_a, int* out_b, const u64* bk, int shape,
                    cudaStream_t st)
{
    if (shape < 1)
        throw std::invalid_argument("empty ciphertext");
    LaunchScope scope(KC_TFHE_BLIND_ROTATE, st);
    k_tfhe_blind_rotate<<<shape, kBrThreads, kBrSmem, st>>>(in_a, in_b, out_a, out_b, bk, c.d_fwd, c.d_inv, c.dev);
    t_check_launch();
}

void tfhe_keyswitch(const TfheContext& c, const int* in_a, const int* in_b, int* out_a, int* out_b, const int* ks_a,
                    const int* ks_b, int shape, cudaStream_t st)
{
    if (shape < 1)
        throw std::invalid_argument("empty ciphertext");
    LaunchScope scope(KC_TFHE_KEYSWITCH, st);
    k_tfhe_keyswitch<<<(shape + KS_S - 1) / KS_S, 512, 0, st>>>(in_a, in_b, out_a, out_b, ks_a, ks_b, c.ks_base_bit, c.ks_length, c.n,
                                                               c.k * c.N, shape);
    t_check_launch();
}

namespace {
struct TScratch {
    void* p = nullptr;
    cudaStream_t st;
    TScratch(size_t bytes, cudaStream_t s) : st(s)
    {
        if (cudaMallocAsync(&p, bytes, s) != cudaSuccess)
            throw std::runtime_error("cudaMallocAsync failed");
    }
    ~TScratch() { cudaFreeAsync(p, st); }
    int* i() const { return (int*) p; }
};
} // namespace

// a complete gate: linear part, bootstrap, key switch (tfhe/operator.cuh:53-812)
void tfhe_gate(const TfheContext& c, int gate, const int* a1, const int* b1, const int* a2, const int* b2, const int* a3,
               const int* b3, int* oa, int* ob, const u64* bk, const int* ks_a, const int* ks_b, int shape, cudaStream_t st)
{
    const int n = c.n, Nk = c.k * c.N;
    if (gate == 7)
    {
        tfhe_gate_linear(c, 7, a1, b1, nullptr, nullptr, oa, ob, n, shape, st);
        return;
    }
    TScratch t1a((size_t) shape * n * 4, st), t1b((size_t) shape * 4, st);
    TScratch t2a((size_t) shape * Nk * 4, st), t2b((size_t) shape * 4, st);
    if (gate != 8)
    {
        tfhe_gate_linear(c, gate, a1, b1, a2, b2, t1a.i(), t1b.i(), n, shape, st);
        tfhe_bootstrap(c, t1a.i(), t1b.i(), t2a.i(), t2b.i(), bk, shape, st);
        tfhe_keyswitch(c, t2a.i(), t2b.i(), oa, ob, ks_a, ks_b, shape, st);
        return;
    }
    // MUX(in1, in2, control) = OR(AND(control, in1), AND(NOT control, in2)), the OR taken on the extracted
    // samples before ONE key switch (tfhe/operator.cuh:688-812)
    if (!a3 || !b3)
        throw std::invalid_argument("MUX needs a control ciphertext");
    TScratch t4a((size_t) shape * Nk * 4, st), t4b((size_t) shape * 4, st);
    TScratch t5a((size_t) shape * Nk * 4, st), t5b((size_t) shape * 4, st);
    tfhe_gate_linear(c, 1, a3, b3, a1, b1, t1a.i(), t1b.i(), n, shape, st);
    tfhe_bootstrap(c, t1a.i(), t1b.i(), t2a.i(), t2b.i(), bk, shape, st);
    tfhe_gate_linear(c, 6, a3, b3, a2, b2, t1a.i(), t1b.i(), n, shape, st);
    tfhe_bootstrap(c, t1a.i(), t1b.i(), t4a.i(), t4b.i(), bk, shape, st);
    tfhe_gate_linear(c, 3, t2a.i(), t2b.i(), t4a.i(), t4b.i(), t5a.i(), t5b.i(), Nk, shape, st);
    tfhe_keyswitch(c, t5a.i(), t5b.i(), oa, ob, ks_a, ks_b, shape, st);
}

void tfhe_keygen_secret(const TfheContext& c, u64 seed, int* lwe_key, int* tlwe_key, cudaStream_t st)
{
    k_tfhe_bits<<<(c.n + 255) / 256, 256, 0, st>>>(lwe_key, c.n, seed, 0x5E0);
    k_tfhe_bits<<<(c.k * c.N + 255) / 256, 256, 0, st>>>(tlwe_key, c.k * c.N, seed, 0x5E1);
    t_check_launch();
}

// bk: [n][k+1][l][k+1][N] words in the NTT domain; ks_a: [kN][length][base-1][n], ks_b: [kN][length][base-1]
void tfhe_keygen_boot(const TfheContext& c, const int* lwe_key, const int* tlwe_key, u64 seed, u64* bk, int* ks_a, int* ks_b,
                      cudaStream_t st)
{
    if (c.k != 1)
        throw std::invalid_argument("k = 1 only");
    k_tfhe_bootkey_row<<<c.n * (c.k + 1) * c.bk_l, 64, 0, st>>>(bk, lwe_key, tlwe_key, c.d_fwd, c.d_inv, c.dev, c.bk_stdev, seed);
    t_check_launch();
    const int rows = c.k * c.N * c.ks_length * ((1 << c.ks_base_bit) - 1);
    k_tfhe_switchkey<<<rows, 256, 0, st>>>(ks_a, ks_b, lwe_key, tlwe_key, c.n, c.ks_base_bit, c.ks_length, c.ks_stdev, seed);
    t_check_launch();
}

void tfhe_encrypt(const TfheContext& c, const int* lwe_key, const int* d_messages, u64 seed, int* out_a, int* out_b, int shape,
                  cudaStream_t st)
{
    if (shape < 1)
        throw std::invalid_argument("empty ciphertext");
    k_tfhe_encrypt<<<shape, 256, 0, st>>>(out_a, out_b, d_messages, lwe_key, c.n, c.ks_stdev, seed);
    t_check_launch();
}

void tfhe_phase(const TfheContext& c, const int* lwe_key, const int* in_a, const int* in_b, int* d_phase, int n, int shape,
                cudaStream_t st)
{
    if (shape < 1)
        throw std::invalid_argument("empty ciphertext");
    k_tfhe_phase<<<shape, 256, 0, st>>>(in_a, in_b, d_phase, lwe_key, n);
    t_check_launch();
}

// forward / inverse 1024-point transform of `count` polynomials in place (tests, key import)
namespace {
__global__ void __launch_bounds__(64) k_tfhe_ntt(u64* data, const TwPair* __restrict__ tw, const TfheDev P, int inverse)
{
    __shared__ __align__(16) u64 buf[TN];
    const int j = threadIdx.x;
    u64* poly = data + (size_t) blockIdx.x * TN;
    const BflyConst bc = make_bc(P.pc);
    u64 v[16];
    if (!inverse)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = poly[j + 64 * k];
        t_ntt_fwd(v, buf, j, 1, tw, bc);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            poly[16 * j + k] = csub(csub(csub(v[k], bc.p4), bc.p2), bc.p);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = poly[16 * j + k];
        t_ntt_inv(v, buf, j, 1, tw, P.ninv, P.wninv, bc);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            poly[j + 64 * k] = v[k];
    }
}
} // namespace

void tfhe_ntt(const TfheContext& c, u64* data, int count, bool inverse, cudaStream_t st)
{
    if (count < 1)
        return;
    k_tfhe_ntt<<<count, 64, 0, st>>>(data, inverse ? c.d_inv : c.d_fwd, c.dev, inverse ? 1 : 0);
    t_check_launch();
}

} // namespace heon
