// CKKS hot-path operators: tensor product, key-switch core (mod-up, inner
// product, mod-down), relinearize, rescale, mod-drop, Galois automorphism,
// element-wise add/sub/negate.  Batched: every ciphertext pointer comes with a
// batch stride (in 64-bit words).
//
// Launch sequencing replaces src/lib/host/ckks/operator.cu:796-1720 of the
// reference; kernels replace src/lib/kernel/{switchkey,multiplication,
// addition}.cu (exact lines cited at each kernel).  All element-wise
// arithmetic uses the reference's Barrett sequence (modarith.cuh), so stored
// words are identical to the reference's.
#include <cstdlib>
#include "modarith.cuh"
#include "ntt_core.cuh"
#include "ops.hpp"

namespace heon {

// ---------------------------------------------------------------------------
// element-wise kernels
// ---------------------------------------------------------------------------

// (c0,c1) x (d0,d1) -> (c0d0, c0d1+c1d0, c1d1) per limb.
// reference: src/lib/kernel/multiplication.cu:102-126 (cross_multiplication)
// For primes below 2^50 the four products run on the FP64 pipe (fp_mulmod with the quotient
// multiplier rebuilt as b*RN(1/p): |T| <= p(1/2 + 3/16)); same canonical residues.
__global__ void __launch_bounds__(256)
    k_cross_multiply(const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ out,
                     long long a_bs, long long b_bs, long long o_bs, const Mod64* __restrict__ mods,
                     const PrimeConst* __restrict__ pcs, int logn, int L)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z;
    const long long loc = idx + ((long long) y << logn);
    const long long comp = (long long) L << logn;
    const u64* pa = a + bz * a_bs;
    const u64* pb = b + bz * b_bs;
    u64* po = out + bz * o_bs;
    const u64 a0 = pa[loc], a1 = pa[loc + comp];
    const u64 b0 = pb[loc], b1 = pb[loc + comp];
    const PrimeConst* pc = pcs + y;
    if (pc->fp_var != 0 && ((a0 | a1 | b0 | b1) >> 50) == 0)
    {
        const double dp = fp_from_u64(pc->p), dnp = -dp, pinv = pc->pinv;
        const double x0 = fp_from_u64(a0), x1 = fp_from_u64(a1), y0 = fp_from_u64(b0), y1 = fp_from_u64(b1);
        const double i0 = __dmul_rn(y0, pinv), i1 = __dmul_rn(y1, pinv);
        const double t00 = fp_mulmod(x0, y0, i0, dnp), t01 = fp_mulmod(x0, y1, i1, dnp);
        const double t10 = fp_mulmod(x1, y0, i0, dnp), t11 = fp_mulmod(x1, y1, i1, dnp);
        po[loc] = fp_canon(t00, pinv, dnp, dp);
        po[loc + comp] = fp_canon(__dadd_rn(t01, t10), pinv, dnp, dp);
        po[loc + 2 * comp] = fp_canon(t11, pinv, dnp, dp);
        return;
    }
    const Mod64 m = mods[y];
    u64 o0 = barrett_mul(a0, b0, m);
    u64 o10 = barrett_mul(a0, b1, m);
    u64 o11 = barrett_mul(a1, b0, m);
    u64 o2 = barrett_mul(a1, b1, m);
    po[loc] = o0;
    po[loc + comp] = mod_add(o10, o11, m.value);
    po[loc + 2 * comp] = o2;
}

// reference: src/lib/kernel/addition.cu:10-49 (addition / substraction / negation)
template <int OP>
__global__ void __launch_bounds__(256)
    k_addsub(const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ out,
             long long a_bs, long long b_bs, long long o_bs, const Mod64* __restrict__ mods, int logn,
             int L, int comps)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z / comps;
    const int c = blockIdx.z % comps;
    const u64 p = mods[y].value;
    const long long loc = idx + ((long long) (c * L + y) << logn);
    u64 x = a[bz * a_bs + loc];
    u64 r;
    if (OP == 0)
        r = mod_add(x, b[bz * b_bs + loc], p);
    else if (OP == 1)
        r = mod_sub(x, b[bz * b_bs + loc], p);
    else
        r = mod_sub(0, x, p);
    out[bz * o_bs + loc] = r;
}

// ct (x) pt per limb, every component.
// reference: src/lib/kernel/multiplication.cu:313-331 (cipherplain_multiplication_kernel)
__global__ void __launch_bounds__(256)
    k_multiply_plain(const u64* __restrict__ ct, const u64* __restrict__ pt, u64* __restrict__ out,
                     long long ct_bs, long long pt_bs, long long o_bs, const Mod64* __restrict__ mods,
                     int logn, int L, int comps)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z / comps;
    const int c = blockIdx.z % comps;
    const long long loc = idx + ((long long) (c * L + y) << logn);
    const u64 x = ct[bz * ct_bs + loc];
    const u64 m = pt[bz * pt_bs + idx + ((long long) y << logn)];
    out[bz * o_bs + loc] = barrett_mul(x, m, mods[y]);
}

// component 0 +/- pt, the other components copied.
// reference: src/lib/kernel/addition.cu:175-217 (addition_plain_ckks_poly, substraction_plain_ckks_poly)
template <int OP>
__global__ void __launch_bounds__(256)
    k_addsub_plain(const u64* __restrict__ ct, const u64* __restrict__ pt, u64* __restrict__ out,
                   long long ct_bs, long long pt_bs, long long o_bs, const Mod64* __restrict__ mods,
                   int logn, int L, int comps)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z / comps;
    const int c = blockIdx.z % comps;
    const long long loc = idx + ((long long) (c * L + y) << logn);
    u64 x = ct[bz * ct_bs + loc];
    if (c == 0)
    {
        const u64 m = pt[bz * pt_bs + idx + ((long long) y << logn)];
        x = OP == 0 ? mod_add(x, m, mods[y].value) : mod_sub(x, m, mods[y].value);
    }
    out[bz * o_bs + loc] = x;
}

// ---------------------------------------------------------------------------
// key-switch inner product
// ---------------------------------------------------------------------------

// out[b][c][y] = sum_i in[b][i][y] * key[i][c][prime(y)]  (NTT domain).
// Products are accumulated lazily in 128 bits and reduced once; the result is
// the canonical residue, identical to the reference's per-term Barrett sum.
// reference: src/lib/kernel/switchkey.cu:164-285 (Method I), 287-398 (Method II)
// One thread owns two adjacent coefficients of one limb of one ciphertext.  The key (larger
// than L2 at the BASELINE sizes) should cross HBM once per batch, not once per ciphertext:
//  * blockDim = (256/BY, BY): the BY warp-rows of a CTA work on BY ciphertexts of the batch
//    and read the SAME key words, which the first reader leaves in L1 (HEON_MAC_BY, default 1);
//  * the batch group is the fastest-varying block coordinate, so CTAs that are resident
//    together share the key tile through L2.
#ifndef HEON_MAC_MINBLOCKS
#define HEON_MAC_MINBLOCKS 4
#endif
#ifndef HEON_MAC_UNROLL
#define HEON_MAC_UNROLL 4
#endif
__global__ void __launch_bounds__(256, HEON_MAC_MINBLOCKS)
    k_keyswitch_mac(const u64* __restrict__ in, const u64* __restrict__ key, u64* __restrict__ out,
                    const PrimeConst* __restrict__ pcs, int logn, int d, int L, int Qpl, int Qp0,
                    int depth, int batch)
{
    const int idx = (blockIdx.y * blockDim.x + threadIdx.x) * 2;
    const int y = blockIdx.z;
    const long long bz = (long long) blockIdx.x * blockDim.y + threadIdx.y;
    if (bz >= batch)
        return;
    const int prime = level_prime(y, L, depth);
    const PrimeConst pc = pcs[prime];
    const u64* pin = in + ((bz * d * Qpl + y) << logn) + idx;
    const u64* pk = key + ((long long) prime << logn) + idx;
    const long long in_step = (long long) Qpl << logn;
    const long long key_c = (long long) Qp0 << logn;
    const long long key_step = 2 * key_c;

    u64 a0l = 0, a0h = 0, a1l = 0, a1h = 0; // coefficient idx
    u64 b0l = 0, b0h = 0, b1l = 0, b1h = 0; // coefficient idx+1
    constexpr int kUnroll = HEON_MAC_UNROLL;
#pragma unroll kUnroll
    for (int i = 0; i < d; ++i)
    {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(pin + i * in_step);
        const ulonglong2 k0 = __ldg(reinterpret_cast<const ulonglong2*>(pk + i * key_step));
        const ulonglong2 k1 = __ldg(reinterpret_cast<const ulonglong2*>(pk + i * key_step + key_c));
        mac128(a0l, a0h, x.x, k0.x);
        mac128(a1l, a1h, x.x, k1.x);
        mac128(b0l, b0h, x.y, k0.y);
        mac128(b1l, b1h, x.y, k1.y);
    }
    u64* po = out + ((bz * 2 * Qpl + y) << logn) + idx;
    ulonglong2 r0, r1;
    r0.x = reduce_u128(a0l, a0h, pc);
    r0.y = reduce_u128(b0l, b0h, pc);
    r1.x = reduce_u128(a1l, a1h, pc);
    r1.y = reduce_u128(b1l, b1h, pc);
    *reinterpret_cast<ulonglong2*>(po) = r0;
    *reinterpret_cast<ulonglong2*>(po + ((long long) Qpl << logn)) = r1;
}

// ---------------------------------------------------------------------------
// Method-II mod-up (HPS fast base conversion with the fp32 correction)
// ---------------------------------------------------------------------------

// reference: src/lib/kernel/switchkey.cu:985-1046
// (base_conversion_DtoQtilde_relin_leveled_kernel).  The float sequence
// (u64->f32 rn, IEEE divide, sequential adds, round half away) is reproduced
// operation for operation; the integer part is restructured:
//   * partial_j = x_j * Mi_inv_j uses a Shoup product (canonical result),
//   * sum_j partial_j * M_{j,k} is accumulated lazily in 128 bits and reduced
//     once per output word,
//   * r * prod_k comes from a small table of multiples,
//   * output limbs that belong to the digit itself equal the input residue
//     (M_{j,k} = 0 for j != k, partial_k * M_{k,k} = x_k, prod_k = 0) and are
//     copied.
//   * when the digit's primes and the target prime are below 2^50 the products run on the
//     FP64 pipe (fp_mulmod, ntt_core.cuh): five DFMA-class instructions per term instead of
//     nine integer multiplies; the table then holds the doubles {M, RN(M/t_k)}.
// All of these are exact, so every output word equals the reference's.
// One thread converts CW adjacent coefficients (CW = 2: 16-byte loads and stores, the
// per-target constants and table words are fetched once for both, and the two dependent
// FP64 chains interleave).
template <int IJ, int CW>
__device__ __forceinline__ void modup2_body(const u64* __restrict__ pc_in, u64* __restrict__ po,
                                            const PrimeConst* __restrict__ pcs,
                                            const TwPair* __restrict__ base_change,
                                            const TwPair* __restrict__ mi_inv,
                                            const u64* __restrict__ rprod, int I_loc, int dg, int d,
                                            int logn, int Qpl, int L, int depth)
{
    u64 x[IJ][CW], partial[IJ][CW];
    double pd[IJ][CW];
    bool dfp = IJ <= 4; // every prime of the digit is FP64-capable (and the lazy sum stays below 2^52)
    float r[CW];
#pragma unroll
    for (int e = 0; e < CW; ++e)
        r[e] = 0;
#pragma unroll
    for (int i = 0; i < IJ; ++i)
    {
        const PrimeConst pi = pcs[I_loc + i];
        dfp = dfp && pi.fp_var != 0;
        if (CW == 2)
        {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(pc_in + ((long long) i << logn));
            x[i][0] = t.x;
            x[i][CW - 1] = t.y;
        }
        else
            x[i][0] = pc_in[(long long) i << logn];
        const TwPair mi = mi_inv[I_loc + i];
        const float mod = __ull2float_rn(pi.p);
#pragma unroll
        for (int e = 0; e < CW; ++e)
        {
            partial[i][e] = csub(shoup_mul_lazy(x[i][e], mi.w, mi.ws, pi.p), pi.p);
            const float div = __ull2float_rn(partial[i][e]);
            r[e] = __fadd_rn(r[e], __fdiv_rn(div, mod));
        }
    }
    const u64* rp[CW];
#pragma unroll
    for (int e = 0; e < CW; ++e)
    {
        const unsigned r_ = (unsigned) roundf(r[e]);
        rp[e] = rprod + ((long long) r_ * d + dg) * Qpl;
    }
    if (dfp)
    {
#pragma unroll
        for (int i = 0; i < IJ; ++i)
#pragma unroll
            for (int e = 0; e < CW; ++e)
                pd[i][e] = fp_from_u64(partial[i][e]);
    }
    const int matrix_index = I_loc * Qpl;
#pragma unroll 2
    for (int k = 0; k < Qpl; ++k)
    {
        u64 res[CW];
        if (k >= I_loc && k < I_loc + IJ)
        {
#pragma unroll
            for (int e = 0; e < CW; ++e)
                res[e] = 0;
#pragma unroll
            for (int i = 0; i < IJ; ++i)
                if (k == I_loc + i)
                {
#pragma unroll
                    for (int e = 0; e < CW; ++e)
                        res[e] = x[i][e];
                }
        }
        else
        {
            const PrimeConst* ppk = pcs + level_prime(k, L, depth);
            const u64 pkp = ppk->p;
            if (dfp && ppk->fp_var != 0)
            {
                const double dp = fp_from_u64(pkp), dnp = -dp, dpinv = ppk->pinv;
                double acc[CW]; // |acc| <= IJ * 0.6p, exact
#pragma unroll
                for (int e = 0; e < CW; ++e)
                    acc[e] = 0.0;
#pragma unroll
                for (int j = 0; j < IJ; ++j)
                {
                    const TwPair m = ld_tw(base_change + j + k * IJ + matrix_index);
#pragma unroll
                    for (int e = 0; e < CW; ++e)
                        acc[e] = __dadd_rn(acc[e], fp_mulmod(pd[j][e], u2d(m.w), u2d(m.ws), dnp));
                }
#pragma unroll
                for (int e = 0; e < CW; ++e)
                    res[e] = fp_canon(__dsub_rn(acc[e], fp_from_u64(rp[e][k])), dpinv, dnp, dp);
            }
            else
            {
                // sum_j partial_j * M_{j,k}: one Shoup product per term (constant multiplier with
                // its companion word, any 64-bit operand allowed), lazily in [0,4p)
                const u64 p4 = 4 * pkp, np = 0 - pkp;
                u64 acc[CW];
#pragma unroll
                for (int e = 0; e < CW; ++e)
                    acc[e] = 0;
#pragma unroll
                for (int j = 0; j < IJ; ++j)
                {
                    const TwPair m = ld_tw(base_change + j + k * IJ + matrix_index);
#pragma unroll
                    for (int e = 0; e < CW; ++e)
                        acc[e] = csub(acc[e] + shoup_lazy_ptx(partial[j][e], m.w, m.ws, np), p4);
                }
#pragma unroll
                for (int e = 0; e < CW; ++e)
                {
                    acc[e] = csub(csub(acc[e], 2 * pkp), pkp);
                    res[e] = mod_sub(acc[e], rp[e][k], pkp);
                }
            }
        }
        if (CW == 2)
        {
            ulonglong2 t;
            t.x = res[0];
            t.y = res[CW - 1];
            *reinterpret_cast<ulonglong2*>(po + ((long long) k << logn)) = t;
        }
        else
            po[(long long) k << logn] = res[0];
    }
}

// Fast form for short digits (I_j <= 4) and Q'_l <= 128: the per-target constants of this digit
// (prime, 1/p, the I_j conversion factors, whether the FP64 path applies, which input limb a
// digit-own target copies) and the r*prod table are staged in shared memory once per CTA, so the
// 34-target loop of the BASELINE parameters runs without global loads or 64-bit address arithmetic.
struct __align__(16) Mu2Rec {
    TwPair m[4];
    u64 p;
    double dp, pinv;
    int fp, self, pad0, pad1;
};
constexpr int kMu2MaxQ = 128;

template <int IJ, int CW>
__device__ __forceinline__ void modup2_fast_body(const u64* __restrict__ pc_in, u64* __restrict__ po,
                                                 const PrimeConst* __restrict__ pcs,
                                                 const TwPair* __restrict__ mi_inv, const Mu2Rec* rec,
                                                 const u64* srp, int I_loc, int logn, int Qpl, bool skip_own, bool emit_doubles)
{
    // CW adjacent coefficients per thread: 2 by default; 4 (HEON_MODUP_CW=4) pays the per-target record, the loop and
    // the address arithmetic once for four words but needs 119 registers -- measured 54.8 against 51.2 us/op at C3-II
    u64 x[IJ][CW], partial[IJ][CW];
    double pd[IJ][CW];
    bool dfp = true;
    float r[CW];
#pragma unroll
    for (int e = 0; e < CW; ++e)
        r[e] = 0.f;
#pragma unroll
    for (int i = 0; i < IJ; ++i)
    {
        const PrimeConst pi = pcs[I_loc + i];
        dfp = dfp && pi.fp_var != 0;
#pragma unroll
        for (int h = 0; h < CW / 2; ++h)
        {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(pc_in + ((long long) i << logn) + 2 * h);
            x[i][2 * h] = t.x;
            x[i][2 * h + 1] = t.y;
        }
        const TwPair mi = mi_inv[I_loc + i];
        const float mod = __ull2float_rn(pi.p);
#pragma unroll
        for (int e = 0; e < CW; ++e)
        {
            partial[i][e] = csub(shoup_mul_lazy(x[i][e], mi.w, mi.ws, pi.p), pi.p);
            const float div = __ull2float_rn(partial[i][e]);
            r[e] = __fadd_rn(r[e], __fdiv_rn(div, mod));
        }
    }
    const u64* rp[CW];
#pragma unroll
    for (int e = 0; e < CW; ++e)
        rp[e] = srp + (unsigned) roundf(r[e]) * Qpl;
    if (dfp)
    {
#pragma unroll
        for (int i = 0; i < IJ; ++i)
#pragma unroll
            for (int e = 0; e < CW; ++e)
                pd[i][e] = fp_from_u64(partial[i][e]);
    }
    const long long lstep = 1ll << logn;
#pragma unroll 2
    for (int k = 0; k < Qpl; ++k, po += lstep)
    {
        const Mu2Rec& rc = rec[k];
        u64 res[CW];
        if (rc.self >= 0)
        {
            if (skip_own)
                continue; // the slot already holds the original NTT-domain words
#pragma unroll
            for (int e = 0; e < CW; ++e)
                res[e] = 0;
#pragma unroll
            for (int i = 0; i < IJ; ++i)
                if (rc.self == i)
                {
#pragma unroll
                    for (int e = 0; e < CW; ++e)
                        res[e] = x[i][e];
                }
        }
        else if (dfp && rc.fp)
        {
            const double dnp = -rc.dp;
            double a[CW]; // |acc| <= IJ * 0.6p, exact
#pragma unroll
            for (int e = 0; e < CW; ++e)
                a[e] = 0.0;
#pragma unroll
            for (int j = 0; j < IJ; ++j)
            {
                const double w = u2d(rc.m[j].w), wi = u2d(rc.m[j].ws);
#pragma unroll
                for (int e = 0; e < CW; ++e)
                    a[e] = __dadd_rn(a[e], fp_mulmod(pd[j][e], w, wi, dnp));
            }
#pragma unroll
            for (int e = 0; e < CW; ++e)
            {
                const double v = __dsub_rn(a[e], fp_from_u64(rp[e][k]));
                // emit_doubles: the column pass that follows works on integer-valued doubles: leave |v| <= p/2 as it is
                res[e] = emit_doubles ? d2u(fp_reduce(v, rc.pinv, dnp)) : fp_canon(v, rc.pinv, dnp, rc.dp);
            }
        }
        else
        {
            const u64 pkp = rc.p, p4 = 4 * pkp, np = 0 - pkp;
            u64 a[CW];
#pragma unroll
            for (int e = 0; e < CW; ++e)
                a[e] = 0;
#pragma unroll
            for (int j = 0; j < IJ; ++j)
#pragma unroll
                for (int e = 0; e < CW; ++e)
                    a[e] = csub(a[e] + shoup_lazy_ptx(partial[j][e], rc.m[j].w, rc.m[j].ws, np), p4);
#pragma unroll
            for (int e = 0; e < CW; ++e)
                res[e] = mod_sub(csub(csub(a[e], 2 * pkp), pkp), rp[e][k], pkp);
        }
#pragma unroll
        for (int h = 0; h < CW / 2; ++h)
        {
            ulonglong2 t;
            t.x = res[2 * h];
            t.y = res[2 * h + 1];
            *reinterpret_cast<ulonglong2*>(po + 2 * h) = t;
        }
    }
}

template <int CW>
__global__ void __launch_bounds__(256)
    k_modup2_fast(const u64* __restrict__ coef, long long coef_bs, u64* __restrict__ out,
                  const PrimeConst* __restrict__ pcs, const TwPair* __restrict__ base_change,
                  const TwPair* __restrict__ mi_inv, const u64* __restrict__ rprod,
                  const int* __restrict__ I_j_, const int* __restrict__ I_loc_, int logn, int d, int Qpl,
                  int L, int depth, int K, int skip_own)
{
    __shared__ Mu2Rec rec[kMu2MaxQ];
    __shared__ u64 srp[kMu2MaxQ * 5];
    const int idx = (blockIdx.x * 256 + threadIdx.x) * CW;
    const int dg = blockIdx.y;
    const long long bz = blockIdx.z;
    const int I_j = I_j_[dg];
    const int I_loc = I_loc_[dg];
    for (int k = threadIdx.x; k < Qpl; k += 256)
    {
        const PrimeConst pk = pcs[level_prime(k, L, depth)];
        Mu2Rec rc;
        for (int j = 0; j < 4; ++j)
            rc.m[j] = j < I_j ? base_change[j + k * I_j + I_loc * Qpl] : TwPair{0, 0};
        rc.p = pk.p;
        rc.dp = (double) pk.p;
        rc.pinv = pk.pinv;
        rc.fp = pk.fp_var != 0;
        rc.self = (k >= I_loc && k < I_loc + I_j) ? k - I_loc : -1;
        rc.pad0 = rc.pad1 = 0;
        rec[k] = rc;
    }
    for (int t = threadIdx.x; t < (K + 1) * Qpl; t += 256)
        srp[t] = rprod[((long long) (t / Qpl) * d + dg) * Qpl + (t % Qpl)];
    __syncthreads();
    const u64* pin = coef + bz * coef_bs + idx + ((long long) I_loc << logn);
    u64* po = out + (((bz * d + dg) * Qpl) << logn) + idx;
    switch (I_j)
    {
        case 1: modup2_fast_body<1, CW>(pin, po, pcs, mi_inv, rec, srp, I_loc, logn, Qpl, (skip_own & 1) != 0, (skip_own & 2) != 0); break;
        case 2: modup2_fast_body<2, CW>(pin, po, pcs, mi_inv, rec, srp, I_loc, logn, Qpl, (skip_own & 1) != 0, (skip_own & 2) != 0); break;
        case 3: modup2_fast_body<3, CW>(pin, po, pcs, mi_inv, rec, srp, I_loc, logn, Qpl, (skip_own & 1) != 0, (skip_own & 2) != 0); break;
        case 4: modup2_fast_body<4, CW>(pin, po, pcs, mi_inv, rec, srp, I_loc, logn, Qpl, (skip_own & 1) != 0, (skip_own & 2) != 0); break;
    }
}

template <int CW>
__global__ void __launch_bounds__(256)
    k_modup2(const u64* __restrict__ coef, long long coef_bs, u64* __restrict__ out,
             const PrimeConst* __restrict__ pcs, const TwPair* __restrict__ base_change,
             const TwPair* __restrict__ mi_inv, const u64* __restrict__ rprod,
             const int* __restrict__ I_j_, const int* __restrict__ I_loc_, int logn, int d, int Qpl,
             int L, int depth)
{
    const int idx = (blockIdx.x * 256 + threadIdx.x) * CW;
    const int dg = blockIdx.y;
    const long long bz = blockIdx.z;
    const int I_j = I_j_[dg];
    const int I_loc = I_loc_[dg];
    const u64* pin = coef + bz * coef_bs + idx + ((long long) I_loc << logn);
    u64* po = out + (((bz * d + dg) * Qpl) << logn) + idx;
#define HEON_MU2(n)                                                                                \
    case n:                                                                                        \
        modup2_body<n, CW>(pin, po, pcs, base_change, mi_inv, rprod, I_loc, dg, d, logn, Qpl, L, depth); \
        break;
    if constexpr (CW == 2)
    {
        switch (I_j)
        {
            HEON_MU2(1)
            HEON_MU2(2)
            HEON_MU2(3)
            HEON_MU2(4)
        }
    }
    else
    {
        switch (I_j)
        {
            HEON_MU2(1)
            HEON_MU2(2)
            HEON_MU2(3)
            HEON_MU2(4)
            HEON_MU2(5)
            HEON_MU2(6)
            HEON_MU2(7)
            HEON_MU2(8)
            HEON_MU2(9)
            HEON_MU2(10)
            HEON_MU2(11)
            HEON_MU2(12)
            HEON_MU2(13)
            HEON_MU2(14)
            HEON_MU2(15)
        }
    }
#undef HEON_MU2
}

// ---------------------------------------------------------------------------
// mod-down
// ---------------------------------------------------------------------------

// Method I, stage two (NTT domain): ct[c][y] += (acc[c][y] - corr[c][y]) * P^-1.
// reference: src/lib/kernel/switchkey.cu:707-736 (stage_two), 738-771
// (switchkey variant: `add_mask` selects which components add the old ct).
__global__ void __launch_bounds__(256)
    k_moddown1_stage2(const u64* __restrict__ corr, const u64* __restrict__ acc,
                      const u64* __restrict__ ct, long long ct_bs, u64* __restrict__ out,
                      long long out_bs, const Mod64* __restrict__ mods,
                      const u64* __restrict__ last_q_modinv, int logn, int L, int Qpl, int add_mask)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z >> 1;
    const int c = blockIdx.z & 1;
    const Mod64 m = mods[y];
    u64 last = corr[((bz * 2 + c) * L + y << logn) + idx];
    u64 x = acc[((bz * 2 + c) * Qpl + y << logn) + idx];
    x = mod_sub(x, last, m.value);
    x = barrett_mul(x, last_q_modinv[y], m);
    u64 cin = 0;
    if ((add_mask >> c) & 1)
        cin = ct[bz * ct_bs + ((long long) (c * L + y) << logn) + idx];
    out[bz * out_bs + ((long long) (c * L + y) << logn) + idx] = mod_add(cin, x, m.value);
}

// Method II / Galois mod-down in the coefficient domain: peel the K special
// primes one at a time (last first).  Optional fused epilogue for
// apply_galois: add c0 to component 0 and scatter through the automorphism
// i -> i*g mod 2N with sign flip (no zero check, as in the reference).
// reference: src/lib/kernel/switchkey.cu:1222-1282
// (divide_round_lastq_extended_leveled_kernel) and 1621-1718
// (divide_round_lastq_permute_ckks_kernel).
// Restructured: one thread owns a coefficient of one component, runs the
// P-limb chain once (the reference redoes it for every Q limb) and then walks
// all Q limbs; x mod q uses an exact 64-bit reduction and the constant
// multipliers use Shoup words.  Same exact values, so identical words.
template <int K, bool PERMUTE>
__device__ __forceinline__ void moddown_body(const u64* __restrict__ pin, u64* __restrict__ pout,
                                             const u64* __restrict__ c0, const PrimeConst* __restrict__ pcs,
                                             const u64* __restrict__ half, const u64* __restrict__ half_mod,
                                             const TwPair* __restrict__ lqm, unsigned galois_elt, int idx,
                                             int comp, int logn, int L, int Qp0, int Q0)
{
    u64 last_ct[K], lh[K];
#pragma unroll
    for (int i = 0; i < K; ++i)
        last_ct[i] = pin[(long long) (L + i) << logn];
    int loc = 0;
#pragma unroll
    for (int i = 0; i < K; ++i)
    {
        lh[i] = mod_add(last_ct[K - 1 - i], half[i], pcs[Qp0 - 1 - i].p);
#pragma unroll
        for (int j = 0; j < K - 1 - i; ++j)
        {
            const PrimeConst pj = pcs[Q0 + j];
            u64 t = reduce_u64(lh[i], pj);
            t = mod_sub(t, half_mod[loc + Q0 + j], pj.p);
            t = mod_sub(last_ct[j], t, pj.p);
            const TwPair w = lqm[loc + Q0 + j];
            last_ct[j] = csub(shoup_mul_lazy(t, w.w, w.ws, pj.p), pj.p);
        }
        loc += Qp0 - 1 - i;
    }
    for (int y = 0; y < L; ++y)
    {
        const PrimeConst py = pcs[y];
        u64 x = pin[(long long) y << logn];
        int l2 = 0;
#pragma unroll
        for (int i = 0; i < K; ++i)
        {
            u64 t = reduce_u64(lh[i], py);
            t = mod_sub(t, half_mod[l2 + y], py.p);
            t = mod_sub(x, t, py.p);
            const TwPair w = lqm[l2 + y];
            x = csub(shoup_mul_lazy(t, w.w, w.ws, py.p), py.p);
            l2 += Qp0 - 1 - i;
        }
        if (!PERMUTE)
        {
            pout[((long long) y << logn) + idx] = x;
        }
        else
        {
            if (comp == 0)
                x = mod_add(c0[(long long) y << logn], x, py.p);
            const unsigned raw = (unsigned) idx * galois_elt; // low n+1 bits are all that matter
            const unsigned dst = raw & ((1u << logn) - 1);
            if ((raw >> logn) & 1)
                x = py.p - x;
            pout[((long long) y << logn) + dst] = x;
        }
    }
}

template <bool PERMUTE>
__global__ void __launch_bounds__(256)
    k_moddown_ext(const u64* __restrict__ in, u64* __restrict__ out, long long out_bs,
                  const u64* __restrict__ c0coef, const PrimeConst* __restrict__ pcs,
                  const u64* __restrict__ half, const u64* __restrict__ half_mod,
                  const TwPair* __restrict__ lqm, unsigned galois_elt, int logn, int Qpl, int L, int Qp0,
                  int Q0, int K, long long c0_bs)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const long long bz = blockIdx.y >> 1;
    const int c = blockIdx.y & 1;
    const u64* pin = in + (((bz * 2 + c) * Qpl) << logn) + idx;
    u64* pout = out + bz * out_bs + ((long long) (c * L) << logn);
    const u64* c0 = PERMUTE ? c0coef + bz * c0_bs + idx : nullptr;
#define HEON_MD(n)                                                                                 \
    case n:                                                                                        \
        moddown_body<n, PERMUTE>(pin, pout, c0, pcs, half, half_mod, lqm, galois_elt, idx, c, logn, L, \
                                 Qp0, Q0);                                                          \
        break;
    switch (K)
    {
        HEON_MD(1)
        HEON_MD(2)
        HEON_MD(3)
        HEON_MD(4)
        HEON_MD(5)
        HEON_MD(6)
        HEON_MD(7)
        HEON_MD(8)
        HEON_MD(9)
        HEON_MD(10)
        HEON_MD(11)
        HEON_MD(12)
        HEON_MD(13)
        HEON_MD(14)
        HEON_MD(15)
    }
#undef HEON_MD
}

// Method-II mod-down without leaving the NTT domain for the Q limbs.
// The reference's coefficient-domain peel (divide_round_lastq_extended_leveled_kernel,
// switchkey.cu:1222-1282) computes, for every Q limb y,
//     x' = (((x - t_0) m_0 - t_1) m_1 ... - t_{K-1}) m_{K-1}  (mod q_y),
// where the t_i depend only on the K special-prime limbs.  That is x' = x*M_y + c_y with
// M_y = prod m_i and c_y = the same chain started from x = 0 -- exact arithmetic mod q_y, and the
// NTT is linear, so NTT(x') = NTT(x)*M_y + NTT(c_y).  NTT(x) is what the inner product produced:
// only the 2K special limbs go through the inverse NTT (instead of all 2Q'), the correction c_y
// costs the 2L forward NTTs the old path spent on x', and the final combination also adds the
// old ciphertext.  Every word equals the reference's.
// For Q primes below 2^50 the chain is evaluated on the FP64 pipe in its expanded form
//   c_y = Cst_y - sum_i lh_i * B_{i,y}  (mod q_y),   lh_i = hi_i*2^30 + lo_i  (lh_i < 2^61),
// two fp_mulmod per special prime with operands below 2^31 (ntt_core.cuh); exact, same residue.
template <int K>
__device__ __forceinline__ void moddown2_corr_body(const u64* __restrict__ pin, u64* __restrict__ pout,
                                                   const PrimeConst* __restrict__ pcs,
                                                   const u64* __restrict__ half, const u64* __restrict__ half_mod,
                                                   const TwPair* __restrict__ lqm, const TwPair* __restrict__ btab,
                                                   const u64* __restrict__ cst, int logn, int L, int Qp0, int Q0)
{
    u64 last_ct[K], lh[K];
#pragma unroll
    for (int i = 0; i < K; ++i)
        last_ct[i] = pin[(long long) i << logn];
    int loc = 0;
#pragma unroll
    for (int i = 0; i < K; ++i)
    {
        lh[i] = mod_add(last_ct[K - 1 - i], half[i], pcs[Qp0 - 1 - i].p);
#pragma unroll
        for (int j = 0; j < K - 1 - i; ++j)
        {
            const PrimeConst pj = pcs[Q0 + j];
            u64 t = reduce_u64(lh[i], pj);
            t = mod_sub(t, half_mod[loc + Q0 + j], pj.p);
            t = mod_sub(last_ct[j], t, pj.p);
            const TwPair w = lqm[loc + Q0 + j];
            last_ct[j] = csub(shoup_mul_lazy(t, w.w, w.ws, pj.p), pj.p);
        }
        loc += Qp0 - 1 - i;
    }
    double dh[K], dl[K];
#pragma unroll
    for (int i = 0; i < K; ++i)
    {
        dh[i] = fp_from_u64(lh[i] >> 30);
        dl[i] = fp_from_u64(lh[i] & 0x3FFFFFFFull);
    }
#pragma unroll 2
    for (int y = 0; y < L; ++y)
    {
        const PrimeConst* ppy = pcs + y;
        const u64 q = ppy->p;
        u64 x;
        if (ppy->fp_var != 0 && K <= 8)
        {
            const double dp = fp_from_u64(q), dnp = -dp;
            double acc = 0.0; // |acc| <= 2K * 0.51 q
#pragma unroll
            for (int i = 0; i < K; ++i)
            {
                const TwPair b0 = ld_tw(btab + ((long long) y * K + i) * 2);
                const TwPair b1 = ld_tw(btab + ((long long) y * K + i) * 2 + 1);
                acc = __dadd_rn(acc, fp_mulmod(dh[i], u2d(b0.w), u2d(b0.ws), dnp));
                acc = __dadd_rn(acc, fp_mulmod(dl[i], u2d(b1.w), u2d(b1.ws), dnp));
            }
            x = fp_canon(__dsub_rn(fp_from_u64(cst[y]), acc), ppy->pinv, dnp, dp);
        }
        else
        {
            const PrimeConst py = *ppy;
            x = 0;
            int l2 = 0;
#pragma unroll
            for (int i = 0; i < K; ++i)
            {
                u64 t = reduce_u64(lh[i], py);
                t = mod_sub(t, half_mod[l2 + y], py.p);
                t = mod_sub(x, t, py.p);
                const TwPair w = lqm[l2 + y];
                x = csub(shoup_mul_lazy(t, w.w, w.ws, py.p), py.p);
                l2 += Qp0 - 1 - i;
            }
        }
        pout[(long long) y << logn] = x;
    }
}

// in: acc[b][2][Qpl][N] whose P limbs are in the coefficient domain; out: corr[b][2][L][N]
__global__ void __launch_bounds__(256)
    k_moddown2_corr(const u64* __restrict__ acc, u64* __restrict__ corr, const PrimeConst* __restrict__ pcs,
                    const u64* __restrict__ half, const u64* __restrict__ half_mod,
                    const TwPair* __restrict__ lqm, const TwPair* __restrict__ btab, const u64* __restrict__ cst,
                    int logn, int Qpl, int L, int Qp0, int Q0, int K)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const long long bc = blockIdx.y; // b*2 + c
    const u64* pin = acc + ((bc * Qpl + L) << logn) + idx;
    u64* pout = corr + ((bc * L) << logn) + idx;
#define HEON_MD2(n)                                                                                \
    case n:                                                                                        \
        moddown2_corr_body<n>(pin, pout, pcs, half, half_mod, lqm, btab, cst, logn, L, Qp0, Q0);   \
        break;
    switch (K)
    {
        HEON_MD2(1)
        HEON_MD2(2)
        HEON_MD2(3)
        HEON_MD2(4)
        HEON_MD2(5)
        HEON_MD2(6)
        HEON_MD2(7)
        HEON_MD2(8)
    }
#undef HEON_MD2
}

// Fast form (K <= 4, L <= 128): per-limb constants staged in shared memory, two coefficients per thread.
struct __align__(16) Md2Rec {
    TwPair b[8]; // {2^30*B_i, B_i} pairs, i < K
    u64 p;
    double dp, pinv, cst;
    int fp, pad0;
};

template <int K>
__device__ __forceinline__ void moddown2_corr_fast_body(const u64* __restrict__ pin, u64* __restrict__ pout,
                                                        const PrimeConst* __restrict__ pcs,
                                                        const u64* __restrict__ half, const u64* __restrict__ half_mod,
                                                        const TwPair* __restrict__ lqm, const Md2Rec* rec, int logn, int L,
                                                        int Qp0, int Q0)
{
    u64 lh[K][2];
    {
        u64 last_ct[K][2];
#pragma unroll
        for (int i = 0; i < K; ++i)
        {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(pin + ((long long) i << logn));
            last_ct[i][0] = t.x;
            last_ct[i][1] = t.y;
        }
        int loc = 0;
#pragma unroll
        for (int i = 0; i < K; ++i)
        {
            const u64 pp = pcs[Qp0 - 1 - i].p;
#pragma unroll
            for (int e = 0; e < 2; ++e)
                lh[i][e] = mod_add(last_ct[K - 1 - i][e], half[i], pp);
#pragma unroll
            for (int j = 0; j < K - 1 - i; ++j)
            {
                const PrimeConst pj = pcs[Q0 + j];
                const TwPair w = lqm[loc + Q0 + j];
                const u64 hm = half_mod[loc + Q0 + j];
#pragma unroll
                for (int e = 0; e < 2; ++e)
                {
                    u64 t = reduce_u64(lh[i][e], pj);
                    t = mod_sub(t, hm, pj.p);
                    t = mod_sub(last_ct[j][e], t, pj.p);
                    last_ct[j][e] = csub(shoup_mul_lazy(t, w.w, w.ws, pj.p), pj.p);
                }
            }
            loc += Qp0 - 1 - i;
        }
    }
    double dh[K][2], dl[K][2];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e)
        {
            dh[i][e] = fp_from_u64(lh[i][e] >> 30);
            dl[i][e] = fp_from_u64(lh[i][e] & 0x3FFFFFFFull);
        }
    const long long lstep = 1ll << logn;
#pragma unroll 2
    for (int y = 0; y < L; ++y, pout += lstep)
    {
        const Md2Rec& rc = rec[y];
        ulonglong2 res;
        if (rc.fp)
        {
            const double dnp = -rc.dp;
            double a0 = 0.0, a1 = 0.0; // |acc| <= 2K * 0.51 q
#pragma unroll
            for (int i = 0; i < K; ++i)
            {
                const double w0 = u2d(rc.b[2 * i].w), i0 = u2d(rc.b[2 * i].ws);
                const double w1 = u2d(rc.b[2 * i + 1].w), i1 = u2d(rc.b[2 * i + 1].ws);
                a0 = __dadd_rn(a0, __dadd_rn(fp_mulmod(dh[i][0], w0, i0, dnp), fp_mulmod(dl[i][0], w1, i1, dnp)));
                a1 = __dadd_rn(a1, __dadd_rn(fp_mulmod(dh[i][1], w0, i0, dnp), fp_mulmod(dl[i][1], w1, i1, dnp)));
            }
            res.x = fp_canon(__dsub_rn(rc.cst, a0), rc.pinv, dnp, rc.dp);
            res.y = fp_canon(__dsub_rn(rc.cst, a1), rc.pinv, dnp, rc.dp);
        }
        else
        {
            const PrimeConst py = pcs[y];
            u64 x[2] = {0, 0};
            int l2 = 0;
#pragma unroll
            for (int i = 0; i < K; ++i)
            {
                const TwPair w = lqm[l2 + y];
                const u64 hm = half_mod[l2 + y];
#pragma unroll
                for (int e = 0; e < 2; ++e)
                {
                    u64 t = reduce_u64(lh[i][e], py);
                    t = mod_sub(t, hm, py.p);
                    t = mod_sub(x[e], t, py.p);
                    x[e] = csub(shoup_mul_lazy(t, w.w, w.ws, py.p), py.p);
                }
                l2 += Qp0 - 1 - i;
            }
            res.x = x[0];
            res.y = x[1];
        }
        *reinterpret_cast<ulonglong2*>(pout) = res;
    }
}

__global__ void __launch_bounds__(256)
    k_moddown2_corr_fast(const u64* __restrict__ acc, u64* __restrict__ corr, const PrimeConst* __restrict__ pcs,
                         const u64* __restrict__ half, const u64* __restrict__ half_mod,
                         const TwPair* __restrict__ lqm, const TwPair* __restrict__ btab,
                         const u64* __restrict__ cst, int logn, int Qpl, int L, int Qp0, int Q0, int K)
{
    __shared__ Md2Rec rec[128];
    for (int y = threadIdx.x; y < L; y += 256)
    {
        const PrimeConst py = pcs[y];
        Md2Rec rc;
        for (int i = 0; i < 8; ++i)
            rc.b[i] = i < 2 * K ? btab[(long long) y * K * 2 + i] : TwPair{0, 0};
        rc.p = py.p;
        rc.dp = (double) py.p;
        rc.pinv = py.pinv;
        rc.cst = (double) cst[y];
        rc.fp = py.fp_var != 0;
        rc.pad0 = 0;
        rec[y] = rc;
    }
    __syncthreads();
    const int idx = (blockIdx.x * 256 + threadIdx.x) * 2;
    const long long bc = blockIdx.y; // b*2 + c
    const u64* pin = acc + ((bc * Qpl + L) << logn) + idx;
    u64* pout = corr + ((bc * L) << logn) + idx;
    switch (K)
    {
        case 1: moddown2_corr_fast_body<1>(pin, pout, pcs, half, half_mod, lqm, rec, logn, L, Qp0, Q0); break;
        case 2: moddown2_corr_fast_body<2>(pin, pout, pcs, half, half_mod, lqm, rec, logn, L, Qp0, Q0); break;
        case 3: moddown2_corr_fast_body<3>(pin, pout, pcs, half, half_mod, lqm, rec, logn, L, Qp0, Q0); break;
        case 4: moddown2_corr_fast_body<4>(pin, pout, pcs, half, half_mod, lqm, rec, logn, L, Qp0, Q0); break;
    }
}

// out[c][y] = (ct[c][y] if selected) + acc[c][y]*M_y + corr[c][y]   (all NTT domain)
__global__ void __launch_bounds__(256)
    k_moddown2_final(const u64* __restrict__ acc, const u64* __restrict__ corr, const u64* __restrict__ ct,
                     long long ct_bs, u64* __restrict__ out, long long out_bs, const PrimeConst* __restrict__ pcs,
                     const TwPair* __restrict__ mprod, int logn, int L, int Qpl, int add_mask)
{
    const int idx = (blockIdx.x * 256 + threadIdx.x) * 2;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z >> 1;
    const int c = blockIdx.z & 1;
    const u64 p = pcs[y].p;
    const TwPair m = mprod[y];
    const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(acc + (((bz * 2 + c) * Qpl + y) << logn) + idx);
    const ulonglong2 k = *reinterpret_cast<const ulonglong2*>(corr + (((bz * 2 + c) * L + y) << logn) + idx);
    ulonglong2 r;
    r.x = mod_add(csub(shoup_mul_lazy(x.x, m.w, m.ws, p), p), k.x, p);
    r.y = mod_add(csub(shoup_mul_lazy(x.y, m.w, m.ws, p), p), k.y, p);
    const long long o = ((long long) (c * L + y) << logn) + idx;
    if ((add_mask >> c) & 1)
    {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(ct + bz * ct_bs + o);
        r.x = mod_add(t.x, r.x, p);
        r.y = mod_add(t.y, r.y, p);
    }
    *reinterpret_cast<ulonglong2*>(out + bz * out_bs + o) = r;
}

// ---------------------------------------------------------------------------
// rescale / mod-drop tails
// ---------------------------------------------------------------------------

// ct[c][y] = (ct[c][y] - corr[c][y]) * q_last^-1, compacted from [2][L] to
// [2][L-1] in place.  One thread walks all limbs of a coefficient in
// ascending order, so reads of slot c*L+y always precede the write that
// reuses it (no staging copy needed).
// reference: src/lib/kernel/switchkey.cu:776-815
// (move_cipher_leveled_kernel + divide_round_lastq_rescale_kernel)
__global__ void __launch_bounds__(256)
    k_rescale_tail(const u64* __restrict__ corr, u64* __restrict__ ct, long long ct_bs,
                   const Mod64* __restrict__ mods, const u64* __restrict__ inv, int logn, int L)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const long long bz = blockIdx.y;
    u64* p = ct + bz * ct_bs + idx;
    const int Lo = L - 1;
    for (int c = 0; c < 2; ++c)
        for (int y = 0; y < Lo; ++y)
        {
            const Mod64 m = mods[y];
            u64 x = p[(long long) (c * L + y) << logn];
            u64 last = corr[(((bz * 2 + c) * Lo + y) << logn) + idx];
            x = mod_sub(x, last, m.value);
            x = barrett_mul(x, inv[y], m);
            p[(long long) (c * Lo + y) << logn] = x;
        }
}

// [comps][L][N] -> [comps][L-1][N] in place (drop the last limb of each component).
// reference: ckks/operator.cu:1246-1276 (mod_drop_ckks_leveled_inplace)
__global__ void __launch_bounds__(256)
    k_mod_drop_inplace(u64* __restrict__ ct, long long ct_bs, int logn, int L, int comps)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    u64* p = ct + (long long) blockIdx.y * ct_bs + idx;
    for (int c = 1; c < comps; ++c)
        for (int y = 0; y < L - 1; ++y)
            p[(long long) (c * (L - 1) + y) << logn] = p[(long long) (c * L + y) << logn];
}

__global__ void __launch_bounds__(256)
    k_mod_drop(const u64* __restrict__ in, long long in_bs, u64* __restrict__ out, long long out_bs,
               int logn, int L)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z >> 1;
    const int c = blockIdx.z & 1;
    out[bz * out_bs + ((long long) (c * (L - 1) + y) << logn) + idx] =
        in[bz * in_bs + ((long long) (c * L + y) << logn) + idx];
}

// Galois automorphism X -> X^g applied IN THE NTT DOMAIN: a pure index permutation, no sign flips.
// With the reference's bit-reversed layout out[i] = a(psi^(2*brev(i)+1)), sigma_g(a) evaluated at the
// same point is a(psi^((2*brev(i)+1)*g)), i.e. word i' of the input with
//     2*brev(i') + 1 = (2*brev(i) + 1) * g  (mod 2N).
// Equal, word for word, to permuting in the coefficient domain and transforming afterwards
// (divide_round_lastq_permute_ckks_kernel + GPU_NTT, switchkey.cu:1621-1718) for canonical data.
// In natural order the map is k -> g*k + (g-1)/2 (mod N); the 32 lanes of a warp differ only in the
// top five bits of k, so their sources are a permutation of ONE aligned 256-byte block: the gather is
// fully coalesced.
__global__ void __launch_bounds__(256)
    k_galois_permute_ntt(const u64* __restrict__ in, long long in_bs, u64* __restrict__ out, long long out_bs,
                         int logn, int L, unsigned galois_elt)
{
    const unsigned idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z >> 1;
    const int c = blockIdx.z & 1;
    const unsigned k = __brev(idx) >> (32 - logn);
    const unsigned f = (((2u * k + 1u) * galois_elt) & ((2u << logn) - 1u)) >> 1;
    const unsigned src = __brev(f) >> (32 - logn);
    const long long limb = (long long) (c * L + y) << logn;
    out[bz * out_bs + limb + idx] = in[bz * in_bs + limb + src];
}

// ---------------------------------------------------------------------------
// workspace (stream-ordered, cached by the device's default memory pool)
// ---------------------------------------------------------------------------

struct Scratch {
    void* p = nullptr;
    cudaStream_t st;
    Scratch(size_t bytes, cudaStream_t s) : st(s)
    {
        cudaError_t e = cudaMallocAsync(&p, bytes, s);
        if (e != cudaSuccess)
            throw std::runtime_error(std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    }
    ~Scratch()
    {
        if (p)
            cudaFreeAsync(p, st);
    }
    u64* w() const { return (u64*) p; }
};

static void check_launch()
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("kernel launch: ") + cudaGetErrorString(e));
}

static void check_depth(const Context& c, int depth)
{
    if (depth < 0 || depth >= c.Q_size)
        throw std::invalid_argument("invalid depth");
}

// ---------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------

void op_add(const Context& c, const u64* a, long long a_bs, const u64* b, long long b_bs, u64* out,
            long long o_bs, int comps, int depth, int batch, int op, cudaStream_t st)
{
    check_depth(c, depth);
    const int L = c.Q_size - depth;
    dim3 g(c.n >> 8, L, batch * comps);
    if (op == 0)
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_addsub<0><<<g, 256, 0, st>>>(a, b, out, a_bs, b_bs, o_bs, c.d_mod, c.logn, L, comps);
        }
    else if (op == 1)
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_addsub<1><<<g, 256, 0, st>>>(a, b, out, a_bs, b_bs, o_bs, c.d_mod, c.logn, L, comps);
        }
    else
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_addsub<2><<<g, 256, 0, st>>>(a, a, out, a_bs, a_bs, o_bs, c.d_mod, c.logn, L, comps);
        }
    check_launch();
}

void op_multiply(const Context& c, const u64* a, long long a_bs, const u64* b, long long b_bs,
                 u64* out, long long o_bs, int depth, int batch, cudaStream_t st)
{
    check_depth(c, depth);
    const int L = c.Q_size - depth;
    dim3 g(c.n >> 8, L, batch);
    {
        LaunchScope scope(KC_CROSS_MULTIPLY, st);
        k_cross_multiply<<<g, 256, 0, st>>>(a, b, out, a_bs, b_bs, o_bs, c.d_mod, c.d_pc, c.logn, L);
    }
    check_launch();
}

// multiply_plain_ckks / add_plain_ckks / sub_plain_ckks (ckks/operator.cu:839-871, 302-345, 434-477):
// op 0 multiply (every component), 1 add, 2 subtract (component 0 only, the rest copied).
void op_plain(const Context& c, const u64* ct, long long ct_bs, const u64* pt, long long pt_bs, u64* out,
              long long o_bs, int comps, int depth, int batch, int op, cudaStream_t st)
{
    check_depth(c, depth);
    if (comps < 1 || comps > 3)
        throw std::invalid_argument("Invalid Ciphertexts size!");
    const int L = c.Q_size - depth;
    dim3 g(c.n >> 8, L, batch * comps);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        if (op == 0)
            k_multiply_plain<<<g, 256, 0, st>>>(ct, pt, out, ct_bs, pt_bs, o_bs, c.d_mod, c.logn, L, comps);
        else if (op == 1)
            k_addsub_plain<0><<<g, 256, 0, st>>>(ct, pt, out, ct_bs, pt_bs, o_bs, c.d_mod, c.logn, L, comps);
        else
            k_addsub_plain<1><<<g, 256, 0, st>>>(ct, pt, out, ct_bs, pt_bs, o_bs, c.d_mod, c.logn, L, comps);
    }
    check_launch();
}

// tmp[b][digit(y)][y] = src[b][y] for every Q limb y: the digit's own limbs of the mod-up output,
// NTT domain, taken from the ciphertext component BEFORE it is brought to the coefficient domain.
__global__ void __launch_bounds__(256)
    k_stash_own_limbs(const u64* __restrict__ src, long long src_bs, u64* __restrict__ tmp,
                      const int* __restrict__ I_j_, const int* __restrict__ I_loc_, int logn, int d, int Qpl)
{
    const int idx = (blockIdx.x * 256 + threadIdx.x) * 2;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z;
    int i = 0;
    while (i + 1 < d && y >= I_loc_[i] + I_j_[i])
        ++i;
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + bz * src_bs + ((long long) y << logn) + idx);
    *reinterpret_cast<ulonglong2*>(tmp + (((bz * d + i) * Qpl + y) << logn) + idx) = v;
}

// Method II with the fast mod-up: true when the own-limb shortcut applies (see keyswitch_stash_own)
static bool own_limb_shortcut(const Context& c, int depth)
{
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    return c.method == 2 && c.skip_own && K <= 4 && Qpl <= kMu2MaxQ && c.n >= 512 && c.lvl2[depth].d <= 64;
}

// `src`: the key-switched component in the NTT domain, [b][L][N] with batch stride src_bs.
static bool keyswitch_stash_own(const Context& c, const u64* src, long long src_bs, u64* tmp, int depth, int batch,
                                cudaStream_t st)
{
    if (!own_limb_shortcut(c, depth) || (src_bs & 1) || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(tmp)) & 15))
        return false;
    const int L = c.Q_size - depth, Qpl = L + c.P_size;
    const LevelTablesII& t = c.lvl2[depth];
    dim3 g(c.n >> 9, L, batch);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_stash_own_limbs<<<g, 256, 0, st>>>(src, src_bs, tmp, t.d_I_j, t.d_I_loc, c.logn, t.d, Qpl);
    }
    check_launch();
    return true;
}

// Key-switch core shared by relinearize / apply_galois / keyswitch:
// takes the digit polynomial(s) in the coefficient domain and leaves
// acc[b][2][Qpl][N] (NTT domain).  `tmp` must hold batch*d*Qpl*N words.
// Part one: mod-up of the digits into every prime of Q'_l and forward NTT (tmp[b][d][Qpl][N]).
// This half does not depend on the key: hoisted rotations run it once for many keys.
static int keyswitch_modup_ntt(const Context& c, const u64* coef, long long coef_bs, u64* tmp, int depth,
                               int batch, cudaStream_t st, bool own_stashed = false, bool col_only = false)
{
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    int d;
    if (c.method == 1)
    {
        d = L;
        if (d > 64)
            throw std::invalid_argument("too many key-switch digits");
        launch_modup1_ntt(c, coef, coef_bs, tmp, L, depth, batch, st, col_only);
    }
    else
    {
        const LevelTablesII& t = c.lvl2[depth];
        d = t.d;
        if (d > 64)
            throw std::invalid_argument("too many key-switch digits");
        if (modup2_col_available(c, depth, coef, coef_bs, tmp, batch, own_stashed, col_only))
        {
            // conversion + column stages in one kernel, the digit's source tiles staged once for all targets
            launch_modup2_col(c, coef, coef_bs, tmp, depth, batch, st);
            check_launch();
            return d;
        }
        if (modup2_fused_available(c, depth, coef, coef_bs))
        {
            // the conversion runs inside the column-pass load: the converted digits are never stored
            Scratch part((size_t) batch * L * c.n * 8, st);
            Scratch rq((size_t) batch * d * c.n, st);
            launch_modup2_ntt(c, coef, coef_bs, tmp, part.w(), (unsigned char*) rq.p, depth, batch, own_stashed, col_only, st);
            check_launch();
            return d;
        }
        // digits whose source primes all have an FP64 form leave their FP64-prime target words as integer-valued
        // doubles (the column pass of MapDigitSkip takes them as they are): only with the fast kernel and the skip map
        unsigned long long dbl_mask = 0;
        // two coefficients per thread when the digits are short and the buffers 16-byte aligned
        const bool wide = K <= 4 && c.n >= 512 && (coef_bs & 1) == 0 &&
                          ((reinterpret_cast<uintptr_t>(coef) | reinterpret_cast<uintptr_t>(tmp)) & 15) == 0;
        dim3 g(wide ? c.n >> 9 : c.n >> 8, d, batch);
        if (own_stashed && wide && Qpl <= kMu2MaxQ && K <= 4 && c.modup_doubles && c.use_fp64)
            for (int i = 0; i < d; ++i)
            {
                bool fp = true;
                for (int j = 0; j < t.I_j[i]; ++j)
                    fp = fp && c.mod[t.I_loc[i] + j].bit <= 50;
                if (fp)
                    dbl_mask |= 1ull << i;
            }
        {
            LaunchScope scope(KC_MODUP2, st);
            if (own_stashed && !(wide && Qpl <= kMu2MaxQ && K <= 4))
                throw std::logic_error("own-limb shortcut needs the fast mod-up");
            if (wide && Qpl <= kMu2MaxQ && K <= 4)
            {
                // opt-in: four coefficients per thread (only when the grid keeps every SM busy with half the CTAs)
                const bool cw4 = c.modup_cw == 4 && c.n >= 1024 && (long long) (c.n >> 10) * d * batch >= 2ll * c.num_sms;
                if (cw4)
                    k_modup2_fast<4><<<dim3(c.n >> 10, d, batch), 256, 0, st>>>(
                        coef, coef_bs, tmp, c.d_pc, t.d_base_change_pair, t.d_mi_inv_pair, t.d_rprod, t.d_I_j, t.d_I_loc,
                        c.logn, d, Qpl, L, depth, K, (own_stashed ? 1 : 0) | (dbl_mask ? 2 : 0));
                else
                    k_modup2_fast<2><<<g, 256, 0, st>>>(coef, coef_bs, tmp, c.d_pc, t.d_base_change_pair, t.d_mi_inv_pair,
                                                     t.d_rprod, t.d_I_j, t.d_I_loc, c.logn, d, Qpl, L, depth, K,
                                                     (own_stashed ? 1 : 0) | (dbl_mask ? 2 : 0));
            }
            else if (wide)
                k_modup2<2><<<g, 256, 0, st>>>(coef, coef_bs, tmp, c.d_pc, t.d_base_change_pair, t.d_mi_inv_pair,
                                            t.d_rprod, t.d_I_j, t.d_I_loc, c.logn, d, Qpl, L, depth);
            else
                k_modup2<1><<<g, 256, 0, st>>>(coef, coef_bs, tmp, c.d_pc, t.d_base_change_pair, t.d_mi_inv_pair,
                                            t.d_rprod, t.d_I_j, t.d_I_loc, c.logn, d, Qpl, L, depth);
        }
        check_launch();
        if (own_stashed)
            launch_ntt_digit_skip(c, tmp, d, t.I_loc.data(), t.I_j.data(), L, depth, batch, st, col_only, dbl_mask);
        else
            launch_ntt(c, tmp, tmp, (long long) batch * d * Qpl, level_primes(L, K, depth), false, st, col_only);
    }
    return d;
}

// Part two: inner product of the NTT-domain digits with the key.
static void keyswitch_mac(const Context& c, const u64* tmp, const u64* key, u64* acc, int d, int depth,
                          int batch, cudaStream_t st)
{
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    {
        LaunchScope scope(KC_KEYSWITCH_MAC, st);
        static const int by_max = [] {
            const char* v = getenv("HEON_MAC_BY");
            return v ? atoi(v) : 1; // measured on B200: sharing through L2 alone is as fast (71.7 vs 73.3 us/op)
        }();
        int by = 1;
        while (by < by_max && by * 2 <= batch && (c.n >> 1) >= 256 / by * 2)
            by *= 2;
        const int tx = 256 / by; // threads along the coefficient axis, two coefficients each
        dim3 g((batch + by - 1) / by, (c.n >> 1) / tx, Qpl), blk(tx, by);
        k_keyswitch_mac<<<g, blk, 0, st>>>(tmp, key, acc, c.d_pc, c.logn, d, L, Qpl, c.Qp, depth, batch);
    }
    check_launch();
}

static int keyswitch_core(const Context& c, const u64* coef, long long coef_bs, const u64* key,
                          u64* tmp, u64* acc, int depth, int batch, cudaStream_t st, bool own_stashed = false)
{
    const int L = c.Q_size - depth;
    const int d0 = (c.method == 1) ? L : c.lvl2[depth].d;
    if (row_mac_available(c, tmp, key, acc, d0))
    {
        // column stages only; the row stages run inside the inner-product kernel and the transformed
        // digits never travel through HBM
        const int d = keyswitch_modup_ntt(c, coef, coef_bs, tmp, depth, batch, st, own_stashed, true);
        const LevelTablesII* t = c.method == 2 ? &c.lvl2[depth] : nullptr;
        launch_row_mac(c, tmp, key, acc, d, depth, batch, own_stashed, t ? t->I_loc.data() : nullptr,
                       t ? t->I_j.data() : nullptr, st);
        check_launch();
        return d;
    }
    const int d = keyswitch_modup_ntt(c, coef, coef_bs, tmp, depth, batch, st, own_stashed);
    keyswitch_mac(c, tmp, key, acc, d, depth, batch, st);
    return d;
}

static size_t ks_tmp_words(const Context& c, int depth, int batch)
{
    const int L = c.Q_size - depth, Qpl = L + c.P_size;
    const int d = (c.method == 1) ? L : c.lvl2[depth].d;
    return (size_t) batch * d * Qpl * c.n;
}

// Mod-down of acc[b][2][Qpl][N] (NTT domain) and accumulation into ct.
//  Method I : INTT the single P limb, stage one fused into the forward NTT of
//             the corrections, stage two in the NTT domain.
//  Method II: INTT everything, coefficient-domain peel, NTT, add.
static void moddown_add(const Context& c, u64* acc, u64* tmp, const u64* ct_in, long long ct_bs,
                        u64* out, long long out_bs, int depth, int batch, int add_mask,
                        cudaStream_t st)
{
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    const long long N = c.n;
    if (c.method == 1)
    {
        // P limb of both components: acc + (b*2+c)*Qpl*N + L*N
        launch_ntt_strided(c, acc + (long long) L * N, Qpl * N, 1, 0, (long long) batch * 2,
                           range_primes(c.Q_size, 1), true, st);
        if (row_final_available(c, tmp, acc, ct_in, ct_bs, out, out_bs, L, batch))
        {
            // column stages of the corrections (divide-and-round stage one in their load), then row stages +
            // stage two in one kernel
            launch_divround1_ntt(c, acc + (long long) L * N, 2 * Qpl * N, Qpl * N, tmp, L, c.half[0],
                                 c.mod[c.Q_size].value, c.d_half_mod, batch, st, true);
            launch_row_final(c, tmp, acc, ct_in, ct_bs, out, out_bs, depth, batch, add_mask, st);
            check_launch();
            return;
        }
        launch_divround1_ntt(c, acc + (long long) L * N, 2 * Qpl * N, Qpl * N, tmp, L, c.half[0],
                             c.mod[c.Q_size].value, c.d_half_mod, batch, st);
        dim3 g(c.n >> 8, L, batch * 2);
        {
            LaunchScope scope(KC_MODDOWN, st);
            k_moddown1_stage2<<<g, 256, 0, st>>>(tmp, acc, ct_in, ct_bs, out, out_bs, c.d_mod,
                                              c.d_last_q_modinv, c.logn, L, Qpl, add_mask);
        }
        check_launch();
    }
    else
    {
        const bool aligned = ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(tmp) |
                               reinterpret_cast<uintptr_t>(ct_in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                             ((ct_bs | out_bs) & 1) == 0;
        if (K <= 8 && aligned && c.n >= 512)
        {
            // inverse NTT of the K special limbs of both components only
            launch_ntt_strided(c, acc + (long long) L * N, Qpl * N, K, 0, (long long) batch * 2,
                               range_primes(c.Q_size, K), true, st);
            {
                LaunchScope scope(KC_MODDOWN, st);
                if (K <= 4 && L <= 128)
                {
                    dim3 g(c.n >> 9, batch * 2);
                    k_moddown2_corr_fast<<<g, 256, 0, st>>>(acc, tmp, c.d_pc, c.d_half, c.d_half_mod, c.d_lqm_pair,
                                                         c.d_md2_B, c.d_md2_cst, c.logn, Qpl, L, c.Qp, c.Q_size, K);
                }
                else
                {
                    dim3 g(c.n >> 8, batch * 2);
                    k_moddown2_corr<<<g, 256, 0, st>>>(acc, tmp, c.d_pc, c.d_half, c.d_half_mod, c.d_lqm_pair, c.d_md2_B,
                                                    c.d_md2_cst, c.logn, Qpl, L, c.Qp, c.Q_size, K);
                }
            }
            check_launch();
            if (row_final_available(c, tmp, acc, ct_in, ct_bs, out, out_bs, L, batch))
            {
                // column stages of the corrections, then row stages + final combination in one kernel: the
                // transformed corrections never travel through HBM
                launch_ntt(c, tmp, tmp, (long long) batch * 2 * L, range_primes(0, L), false, st, true);
                launch_row_final(c, tmp, acc, ct_in, ct_bs, out, out_bs, depth, batch, add_mask, st);
                check_launch();
                return;
            }
            launch_ntt(c, tmp, tmp, (long long) batch * 2 * L, range_primes(0, L), false, st);
            {
                dim3 g(c.n >> 9, L, batch * 2);
                LaunchScope scope(KC_MODDOWN, st);
                k_moddown2_final<<<g, 256, 0, st>>>(acc, tmp, ct_in, ct_bs, out, out_bs, c.d_pc, c.d_md2_M, c.logn, L,
                                                 Qpl, add_mask);
            }
            check_launch();
            return;
        }
        launch_ntt(c, acc, acc, (long long) batch * 2 * Qpl, level_primes(L, K, depth), true, st);
        dim3 g(c.n >> 8, batch * 2);
        {
            LaunchScope scope(KC_MODDOWN, st);
            k_moddown_ext<false><<<g, 256, 0, st>>>(acc, tmp, 2 * L * N, nullptr, c.d_pc, c.d_half,
                                                c.d_half_mod, c.d_lqm_pair, 0, c.logn, Qpl, L,
                                                c.Qp, c.Q_size, K, 0);
        }
        check_launch();
        launch_ntt(c, tmp, tmp, (long long) batch * 2 * L, range_primes(0, L), false, st);
        // out = tmp + ct (components selected by add_mask)
        for (int comp = 0; comp < 2; ++comp)
        {
            dim3 g2(c.n >> 8, L, batch);
            const u64* t = tmp + (long long) comp * L * N;
            u64* o = out + (long long) comp * L * N;
            if ((add_mask >> comp) & 1)
                {
                    LaunchScope scope(KC_ELEMENTWISE, st);
                    k_addsub<0><<<g2, 256, 0, st>>>(t, ct_in + (long long) comp * L * N, o, 2 * L * N,
                                               ct_bs, out_bs, c.d_mod, c.logn, L, 1);
                }
            else
                cudaMemcpy2DAsync(o, out_bs * 8, t, 2 * L * N * 8, L * N * 8, batch,
                                  cudaMemcpyDeviceToDevice, st);
            check_launch();
        }
    }
}

// Coefficient-domain divide-and-round by every special prime of acc-shaped input [b][2][Q'][N] -> [b][2][Q][N]
// (depth 0).  Used by public-key encryption (enc_div_lastq_*_kernel, encryption.cu:30-210).
void op_moddown_coeff(const Context& c, const u64* in, u64* out, long long out_bs, int batch, cudaStream_t st)
{
    const int L = c.Q_size, K = c.P_size, Qpl = L + K;
    dim3 g(c.n >> 8, batch * 2);
    {
        LaunchScope scope(KC_MODDOWN, st);
        k_moddown_ext<false><<<g, 256, 0, st>>>(in, out, out_bs, nullptr, c.d_pc, c.d_half, c.d_half_mod, c.d_lqm_pair, 0, c.logn,
                                                Qpl, L, c.Qp, c.Q_size, K, 0);
    }
    check_launch();
}

// ct: [b][3][L][N] NTT domain, in place; on return components 0,1 hold the
// relinearized ciphertext and component 2 holds INTT(c2) (as in the reference).
// reference: ckks/operator.cu:899-1023 (Method I), 1025-1154 (Method II)
void op_relinearize(const Context& c, u64* ct, long long ct_bs, const u64* relin_key, int depth,
                    int batch, cudaStream_t st)
{
    check_depth(c, depth);
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    const long long N = c.n;
    Scratch tmp(ks_tmp_words(c, depth, batch) * 8, st);
    Scratch acc((size_t) batch * 2 * Qpl * N * 8, st);
    // the digits' own limbs of the mod-up output are c2 itself (NTT domain): keep them before the INTT
    const bool own = keyswitch_stash_own(c, ct + 2LL * L * N, ct_bs, tmp.w(), depth, batch, st);
    // INTT c2 in place
    launch_ntt_strided(c, ct, ct_bs, L, 2 * L, batch, range_primes(0, L), true, st);
    keyswitch_core(c, ct + 2LL * L * N, ct_bs, relin_key, tmp.w(), acc.w(), depth, batch, st, own);
    moddown_add(c, acc.w(), tmp.w(), ct, ct_bs, ct, ct_bs, depth, batch, 3, st);
}

// out = (c0, 0) + KeySwitch(c1): re-encrypts `in` under the key the switch key targets.
// in, out: [b][2][L][N], NTT domain, distinct buffers.
// reference: ckks/operator.cu:1722-1863 (switchkey_ckks_method_I), 1865-2025 (method_II)
void op_keyswitch(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                  const u64* switch_key, int depth, int batch, cudaStream_t st)
{
    check_depth(c, depth);
    if (c.scheme != SCHEME_CKKS)
        throw std::invalid_argument("not a CKKS context");
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    const long long N = c.n;
    // INTT(c1) into scratch (the input stays untouched)
    Scratch coef((size_t) batch * L * N * 8, st);
    launch_ntt_strided_copy(c, in + (long long) L * N, in_bs, coef.w(), L, batch, range_primes(0, L), true, st);
    Scratch tmp(ks_tmp_words(c, depth, batch) * 8, st);
    Scratch acc((size_t) batch * 2 * Qpl * N * 8, st);
    const bool own = keyswitch_stash_own(c, in + (long long) L * N, in_bs, tmp.w(), depth, batch, st);
    keyswitch_core(c, coef.w(), L * N, switch_key, tmp.w(), acc.w(), depth, batch, st, own);
    moddown_add(c, acc.w(), tmp.w(), in, in_bs, out, out_bs, depth, batch, 1, st);
}

// BFV relinearize: ct [b][3][Q][N] in the COEFFICIENT domain, in place.
//   mod-up of c2 (Method I: exact reduction into every prime of Q', Method II: HPS base
//   conversion) -> NTT -> inner product with the key -> INTT of all 2*Q' limbs ->
//   coefficient-domain divide-and-round by P -> add to (c0, c1).
// reference: bfv/operator.cu:505-590 (relinearize_seal_method_inplace),
//            :592-671 (relinearize_external_product_method2_inplace)
void op_bfv_relinearize(const Context& c, u64* ct, long long ct_bs, const u64* relin_key, int batch,
                        cudaStream_t st)
{
    if (c.scheme != SCHEME_BFV)
        throw std::invalid_argument("not a BFV context");
    const int L = c.Q_size, K = c.P_size, Qpl = L + K;
    const long long N = c.n;
    Scratch tmp(ks_tmp_words(c, 0, batch) * 8, st);
    Scratch acc((size_t) batch * 2 * Qpl * N * 8, st);
    keyswitch_core(c, ct + 2LL * L * N, ct_bs, relin_key, tmp.w(), acc.w(), 0, batch, st);
    launch_ntt(c, acc.w(), acc.w(), (long long) batch * 2 * Qpl, level_primes(L, K, 0), true, st);
    dim3 g(c.n >> 8, batch * 2);
    {
        LaunchScope scope(KC_MODDOWN, st);
        k_moddown_ext<false><<<g, 256, 0, st>>>(acc.w(), tmp.w(), 2 * L * N, nullptr, c.d_pc, c.d_half,
                                                c.d_half_mod, c.d_lqm_pair, 0, c.logn, Qpl, L, c.Qp,
                                                c.Q_size, K, 0);
    }
    check_launch();
    dim3 g2(c.n >> 8, L, batch * 2);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_addsub<0><<<g2, 256, 0, st>>>(tmp.w(), ct, ct, 2 * L * N, ct_bs, ct_bs, c.d_mod, c.logn, L, 2);
    }
    check_launch();
}

// ct: [b][2][L][N] -> [b][2][L-1][N] compacted in place.
// reference: ckks/operator.cu:1156-1244 (rescale_inplace_ckks_leveled)
void op_rescale(const Context& c, u64* ct, long long ct_bs, int depth, int batch, cudaStream_t st)
{
    check_depth(c, depth);
    const int L = c.Q_size - depth;
    if (L < 2)
        throw std::logic_error("Ciphertext modulus can not be dropped!");
    const long long N = c.n;
    int location = 0;
    for (int i = 0, cnt = c.Q_size - 1; i < depth; ++i, --cnt)
        location += cnt;
    // INTT the last limb of both components: ct + b*bs + (c*L + L-1)*N
    // (two strided launches: component 0 and component 1)
    for (int comp = 0; comp < 2; ++comp)
        launch_ntt_strided(c, ct, ct_bs, 1, comp * L + L - 1, batch, range_primes(L - 1, 1), true,
                           st);
    Scratch tmp((size_t) batch * 2 * (L - 1) * N * 8, st);
    launch_divround1_ntt(c, ct + (long long) (L - 1) * N, ct_bs, L * N, tmp.w(), L - 1,
                         c.rescaled_half[depth], c.mod[L - 1].value,
                         c.d_rescaled_half_mod + location, batch, st);
    dim3 g(c.n >> 8, batch);
    {
        LaunchScope scope(KC_MODDOWN, st);
        k_rescale_tail<<<g, 256, 0, st>>>(tmp.w(), ct, ct_bs, c.d_mod,
                                       c.d_rescaled_last_q_modinv + location, c.logn, L);
    }
    check_launch();
}

void op_mod_drop_inplace(const Context& c, u64* ct, long long ct_bs, int comps, int depth, int batch,
                         cudaStream_t st)
{
    check_depth(c, depth);
    const int L = c.Q_size - depth;
    if (L < 2)
        throw std::logic_error("Ciphertext modulus can not be dropped!");
    dim3 g(c.n >> 8, batch);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_mod_drop_inplace<<<g, 256, 0, st>>>(ct, ct_bs, c.logn, L, comps);
    }
    check_launch();
}

void op_mod_drop(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                 int depth, int batch, cudaStream_t st)
{
    check_depth(c, depth);
    const int L = c.Q_size - depth;
    if (L < 2)
        throw std::logic_error("Ciphertext modulus can not be dropped!");
    dim3 g(c.n >> 8, L - 1, batch * 2);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_mod_drop<<<g, 256, 0, st>>>(in, in_bs, out, out_bs, c.logn, L);
    }
    check_launch();
}

// out = automorphism_g(in) key-switched back to the original key.
// CKKS (ciphertexts in the NTT domain): ckks/operator.cu:1422-1559 (Method I), 1561-1720 (II).
// BFV  (ciphertexts in the coefficient domain, depth 0): bfv/operator.cu:771-973
//       -- the same pipeline without the leading INTT and the trailing NTT.
void op_apply_galois(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                     const u64* galois_key, unsigned galois_elt, int depth, int batch,
                     cudaStream_t st)
{
    check_depth(c, depth);
    const bool coeff = c.scheme == SCHEME_BFV;
    if (coeff && depth != 0)
        throw std::invalid_argument("BFV ciphertexts have no levels");
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    const long long N = c.n;
    if (!coeff && c.galois_ntt)
    {
        // CKKS: key switch entirely in the NTT domain (only c1 and the 2K special limbs ever leave it),
        // then the automorphism as an index permutation of the NTT words
        Scratch ks((size_t) batch * 2 * L * N * 8, st);
        op_keyswitch(c, in, in_bs, ks.w(), 2 * L * N, galois_key, depth, batch, st);
        dim3 g(c.n >> 8, L, batch * 2);
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_galois_permute_ntt<<<g, 256, 0, st>>>(ks.w(), 2 * L * N, out, out_bs, c.logn, L, galois_elt);
        }
        check_launch();
        return;
    }
    Scratch coef(coeff ? 8 : (size_t) batch * 2 * L * N * 8, st);
    const u64* cp = in;
    long long cbs = in_bs;
    if (!coeff)
    {
        launch_ntt_strided_copy(c, in, in_bs, coef.w(), 2 * L, batch, range_primes(0, L), true, st);
        cp = coef.w();
        cbs = 2 * L * N;
    }
    Scratch tmp(ks_tmp_words(c, depth, batch) * 8, st);
    Scratch acc((size_t) batch * 2 * Qpl * N * 8, st);
    keyswitch_core(c, cp + (long long) L * N, cbs, galois_key, tmp.w(), acc.w(), depth, batch, st);
    launch_ntt(c, acc.w(), acc.w(), (long long) batch * 2 * Qpl, level_primes(L, K, depth), true, st);
    dim3 g(c.n >> 8, batch * 2);
    {
        LaunchScope scope(KC_MODDOWN, st);
        k_moddown_ext<true><<<g, 256, 0, st>>>(acc.w(), out, out_bs, cp, c.d_pc, c.d_half,
                                           c.d_half_mod, c.d_lqm_pair, galois_elt, c.logn, Qpl,
                                           L, c.Qp, c.Q_size, K, cbs);
    }
    check_launch();
    if (!coeff)
        launch_ntt_strided(c, out, out_bs, 2 * L, 0, batch, range_primes(0, L), false, st);
}

// Hoisted rotations: `count` automorphisms of the SAME ciphertext(s).  INTT, mod-up and the
// d*Q' forward NTTs -- more than half of a rotation -- run once; each rotation then costs one
// inner product with its own key, the INTT of the 2*Q' accumulator limbs, the mod-down fused with
// the permutation and the final NTT.  Rotation r is written to out + r*out_rs (+ b*out_bs) and is
// bit-identical to apply_galois(in, key_r, elt_r): the automorphism is applied after the key
// switch (as in the reference), so everything before the inner product is independent of it.
// This is the baby-step loop of the reference's BSGS matrix-vector product
// (fast_single_hoisting_rotation_ckks_method_I/II, ckks/operator.cu:4674-4954, 5092-5446), which
// repeats the full pipeline for every shift.
void op_rotate_hoisted(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                       long long out_rs, const u64* const* galois_keys, const unsigned* galois_elts, int count,
                       int depth, int batch, cudaStream_t st)
{
    check_depth(c, depth);
    if (c.scheme != SCHEME_CKKS)
        throw std::invalid_argument("not a CKKS context");
    if (count < 1)
        throw std::invalid_argument("no rotations requested");
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    const long long N = c.n;
    if (c.galois_ntt)
    {
        // NTT-domain form: only c1 is taken to the coefficient domain (once); per rotation the inner
        // product, the NTT-domain mod-down (+ c0) and the index permutation
        Scratch coef1((size_t) batch * L * N * 8, st);
        launch_ntt_strided_copy(c, in + (long long) L * N, in_bs, coef1.w(), L, batch, range_primes(0, L), true, st);
        Scratch tmp(ks_tmp_words(c, depth, batch) * 8, st);
        Scratch acc((size_t) batch * 2 * Qpl * N * 8, st);
        Scratch corr((size_t) batch * 2 * L * N * 8, st);
        Scratch ks((size_t) batch * 2 * L * N * 8, st);
        const bool own = keyswitch_stash_own(c, in + (long long) L * N, in_bs, tmp.w(), depth, batch, st);
        const int d = keyswitch_modup_ntt(c, coef1.w(), L * N, tmp.w(), depth, batch, st, own);
        for (int r = 0; r < count; ++r)
        {
            keyswitch_mac(c, tmp.w(), galois_keys[r], acc.w(), d, depth, batch, st);
            moddown_add(c, acc.w(), corr.w(), in, in_bs, ks.w(), 2 * L * N, depth, batch, 1, st);
            dim3 g(c.n >> 8, L, batch * 2);
            {
                LaunchScope scope(KC_ELEMENTWISE, st);
                k_galois_permute_ntt<<<g, 256, 0, st>>>(ks.w(), 2 * L * N, out + (long long) r * out_rs, out_bs, c.logn,
                                                     L, galois_elts[r]);
            }
            check_launch();
        }
        return;
    }
    Scratch coef((size_t) batch * 2 * L * N * 8, st);
    launch_ntt_strided_copy(c, in, in_bs, coef.w(), 2 * L, batch, range_primes(0, L), true, st);
    Scratch tmp(ks_tmp_words(c, depth, batch) * 8, st);
    Scratch acc((size_t) batch * 2 * Qpl * N * 8, st);
    const int d = keyswitch_modup_ntt(c, coef.w() + (long long) L * N, 2 * L * N, tmp.w(), depth, batch, st);
    for (int r = 0; r < count; ++r)
    {
        u64* o = out + (long long) r * out_rs;
        keyswitch_mac(c, tmp.w(), galois_keys[r], acc.w(), d, depth, batch, st);
        launch_ntt(c, acc.w(), acc.w(), (long long) batch * 2 * Qpl, level_primes(L, K, depth), true, st);
        dim3 g(c.n >> 8, batch * 2);
        {
            LaunchScope scope(KC_MODDOWN, st);
            k_moddown_ext<true><<<g, 256, 0, st>>>(acc.w(), o, out_bs, coef.w(), c.d_pc, c.d_half, c.d_half_mod,
                                                   c.d_lqm_pair, galois_elts[r], c.logn, Qpl, L, c.Qp, c.Q_size, K,
                                                   2 * L * N);
        }
        check_launch();
        launch_ntt_strided(c, o, out_bs, 2 * L, 0, batch, range_primes(0, L), false, st);
    }
}


// ---------------------------------------------------------------------------
// BSGS diagonal matrix-vector product with double hoisting in the PQ_l domain
// ---------------------------------------------------------------------------
// reference: HEOperator<CKKS>::multiply_matrix_v2 (ckks/operator.cu:2898-3390), the linear-transform core of
// CKKS bootstrapping (CoeffToSlot / SlotToCoeff), with the kernels galois_permute_ntt_pql_kernel,
// broadcast_scale_P_kernel, addition_pql_kernel (switchkey.cu:1482-1554) and
// cipherplain_multiply_accumulate_indexed_kernel (multiplication.cu:405-439).
// PQ_l = the limb set {q_0..q_{L-1}, p_0..p_{K-1}} of the key switch at this depth (pql = L + K limbs).

// out[y] = (P mod q_y) * c[y] for the Q limbs, 0 for the P limbs   (broadcast_scale_P_kernel)
__global__ void __launch_bounds__(256)
    k_pql_scale_P(const u64* __restrict__ c, u64* __restrict__ out, const Mod64* __restrict__ mods,
                  const u64* __restrict__ pmodq, int logn, int L)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    u64 r = 0;
    if (y < L)
        r = barrett_mul(c[((long long) y << logn) + idx], pmodq[y], mods[y]);
    out[((long long) y << logn) + idx] = r;
}
// out = a + b over `comps` components of pql limbs   (addition_pql_kernel)
__global__ void __launch_bounds__(256)
    k_pql_add(const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ out, const Mod64* __restrict__ mods,
              int logn, int L, int depth, int pql)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long off = idx + ((long long) y << logn) + (((long long) pql << logn) * blockIdx.z);
    out[off] = mod_add(a[off], b[off], mods[level_prime(y, L, depth)].value);
}
// out[idx] (+)= acc[pi(idx)] + (component 0 ? add0[pi(idx)] : 0): addition_pql_kernel on component 0, then
// galois_permute_ntt_pql_kernel, then (ACCUM) addition_pql_kernel into the giant-step accumulator, as one pass.
// Sums of canonical words reduced once per addition, exactly as the three reference kernels do.
template <bool ACCUM>
__global__ void __launch_bounds__(256)
    k_pql_add_permute(const u64* __restrict__ acc, const u64* __restrict__ add0, u64* __restrict__ out,
                      const Mod64* __restrict__ mods, int logn, int L, int depth, int pql, unsigned galois_elt)
{
    const unsigned idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, c = blockIdx.z;
    const unsigned k = __brev(idx) >> (32 - logn);
    const unsigned f = (((2u * k + 1u) * galois_elt) & ((2u << logn) - 1u)) >> 1;
    const unsigned src = __brev(f) >> (32 - logn);
    const long long limb = (long long) (c * pql + y) << logn;
    const u64 p = mods[level_prime(y, L, depth)].value;
    u64 v = acc[limb + src];
    if (c == 0)
        v = mod_add(v, add0[((long long) y << logn) + src], p);
    if (ACCUM)
        v = mod_add(out[limb + idx], v, p);
    out[limb + idx] = v;
}

// u[c][y] = sum_i baby[index[i]][c][y] * diag[i][y]   (cipherplain_multiply_accumulate_indexed_kernel);
// lazy 128-bit accumulation, one reduction: the canonical value of the reference's per-term Barrett sum
__global__ void __launch_bounds__(256)
    k_pql_mac_indexed(const u64* __restrict__ baby, const u64* __restrict__ diag, u64* __restrict__ out,
                      const PrimeConst* __restrict__ pcs, const int* __restrict__ index, int terms, int logn, int L,
                      int depth, int pql)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long loc = idx + ((long long) y << logn) + (((long long) pql << logn) * blockIdx.z);
    const long long ct_stride = (long long) pql << (logn + 1);
    const long long pt_stride = (long long) pql << logn;
    u64 lo = 0, hi = 0;
    for (int i = 0; i < terms; ++i)
        mac128(lo, hi, baby[loc + (index ? index[i] : i) * ct_stride], diag[idx + ((long long) y << logn) + i * pt_stride]);
    out[loc] = reduce_u128(lo, hi, pcs[level_prime(y, L, depth)]);
}

// out = sum_i cts[i] * pts[i] over the L limbs of the depth: the giant-step inner sum of the single-hoisting
// BSGS product, cipherplain_multiply_accumulate_kernel (multiplication.cu:374-403) as multiply_matrix launches it
// (ckks/operator.cu:2843-2853).  cts: [count][2][L][N], pts: [count][L][N], out: [2][L][N], NTT domain.
void op_multiply_plain_accumulate(const Context& c, const u64* cts, const u64* pts, u64* out, int count, int depth,
                                  cudaStream_t st)
{
    check_depth(c, depth);
    if (c.scheme != SCHEME_CKKS)
        throw std::invalid_argument("not a CKKS context");
    if (count < 1)
        throw std::invalid_argument("no terms");
    const int L = c.Q_size - depth;
    LaunchScope scope(KC_ELEMENTWISE, st);
    k_pql_mac_indexed<<<dim3(c.n >> 8, L, 2), 256, 0, st>>>(cts, pts, out, c.d_pc, nullptr, count, c.logn, L, depth, L);
    check_launch();
}

// in: [2][L][N] NTT domain; out: [2][L][N] NTT domain at the same depth (the reference rescales afterwards).
// baby step i: Galois element baby_elts[i] (0 = no rotation) with key baby_keys[i]; giant step j: element
// giant_elts[j] (0 = none), key giant_keys[j], group_sizes[j] terms; term t (in group order) multiplies baby
// step term_baby[t] with the plaintext diagonal diags + t*pql*N ([pql][N], NTT domain over PQ_l).
void op_bsgs_matvec(const Context& c, const u64* in, u64* out, const u64* diags, const unsigned* baby_elts,
                    const u64* const* baby_keys, int n1, const unsigned* giant_elts, const u64* const* giant_keys,
                    const int* group_sizes, const int* term_baby, int n2, int depth, cudaStream_t st)
{
    check_depth(c, depth);
    if (c.scheme != SCHEME_CKKS || c.method != 2)
        throw std::invalid_argument("multiply_matrix needs a CKKS context with key-switching Method II");
    if (n1 < 1 || n2 < 1)
        throw std::invalid_argument("empty BSGS plan");
    const int L = c.Q_size - depth, K = c.P_size, pql = L + K;
    const long long N = c.n;
    const size_t ct_words = (size_t) 2 * pql * N;
    // P mod q_y
    std::vector<u64> pmodq(L);
    for (int y = 0; y < L; ++y)
    {
        u64 f = 1;
        for (int k = 0; k < K; ++k)
            f = mulmod(f, c.mod[c.Q_size + k].value % c.mod[y].value, c.mod[y].value);
        pmodq[y] = f;
    }
    int total_terms = 0;
    for (int j = 0; j < n2; ++j)
        total_terms += group_sizes[j];
    for (int t = 0; t < total_terms; ++t)
        if (term_baby[t] < 0 || term_baby[t] >= n1)
            throw std::invalid_argument("baby-step index out of range");
    Scratch small((size_t) L * 8 + (size_t) total_terms * 4 + 64, st);
    u64* d_pmodq = small.w();
    int* d_terms = (int*) (small.w() + L);
    cudaMemcpyAsync(d_pmodq, pmodq.data(), (size_t) L * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_terms, term_baby, (size_t) total_terms * 4, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st); // the staging vector lives on this stack frame

    Scratch coef1((size_t) L * N * 8, st);
    Scratch tmp(ks_tmp_words(c, depth, 1) * 8, st);
    Scratch acc(ct_words * 8, st), Pc0((size_t) pql * N * 8, st), baby(ct_words * n1 * 8, st);
    Scratch accum(ct_words * 8, st), u(ct_words * 8, st), u1q((size_t) L * N * 8, st);
    const dim3 g1(c.n >> 8, pql, 1), g2(c.n >> 8, pql, 2);
    const PrimeList pl = level_primes(L, K, depth);

    // hoisted decomposition of c1 (shared by every baby step)
    launch_ntt_strided_copy(c, in + (long long) L * N, 0, coef1.w(), L, 1, range_primes(0, L), true, st);
    const bool own = keyswitch_stash_own(c, in + (long long) L * N, 0, tmp.w(), depth, 1, st);
    const int d = keyswitch_modup_ntt(c, coef1.w(), L * N, tmp.w(), depth, 1, st, own);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_pql_scale_P<<<g1, 256, 0, st>>>(in, Pc0.w(), c.d_mod, d_pmodq, c.logn, L);
    }
    for (int i = 0; i < n1; ++i)
    {
        u64* bi = baby.w() + (size_t) i * ct_words;
        if (baby_elts[i] == 0)
        {
            cudaMemcpyAsync(bi, Pc0.w(), (size_t) pql * N * 8, cudaMemcpyDeviceToDevice, st);
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_pql_scale_P<<<g1, 256, 0, st>>>(in + (long long) L * N, bi + (size_t) pql * N, c.d_mod, d_pmodq, c.logn, L);
            continue;
        }
        if (!baby_keys[i])
            throw std::logic_error("Galois key not present!");
        keyswitch_mac(c, tmp.w(), baby_keys[i], acc.w(), d, depth, 1, st);
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_pql_add_permute<false><<<g2, 256, 0, st>>>(acc.w(), Pc0.w(), bi, c.d_mod, c.logn, L, depth, pql, baby_elts[i]);
        }
    }
    check_launch();
    cudaMemsetAsync(accum.p, 0, ct_words * 8, st);
    int counter = 0;
    for (int j = 0; j < n2; ++j)
    {
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_pql_mac_indexed<<<g2, 256, 0, st>>>(baby.w(), diags + (size_t) counter * pql * N, u.w(), c.d_pc, d_terms + counter,
                                                 group_sizes[j], c.logn, L, depth, pql);
        }
        counter += group_sizes[j];
        if (giant_elts[j] == 0)
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_pql_add<<<g2, 256, 0, st>>>(accum.w(), u.w(), accum.w(), c.d_mod, c.logn, L, depth, pql);
            continue;
        }
        if (!giant_keys[j])
            throw std::logic_error("Galois key not present!");
        // u1: PQ_l NTT -> coefficients -> divide-round by P -> Q_l coefficients -> decompose -> NTT
        u64* u1 = u.w() + (size_t) pql * N;
        launch_ntt(c, u1, u1, pql, pl, true, st);
        {
            LaunchScope scope(KC_MODDOWN, st);
            k_moddown_ext<false><<<dim3(c.n >> 8, 1), 256, 0, st>>>(u1, u1q.w(), 0, nullptr, c.d_pc, c.d_half, c.d_half_mod,
                                                                 c.d_lqm_pair, 0, c.logn, pql, L, c.Qp, c.Q_size, K, 0);
        }
        // the giant step's own decomposition reuses the buffer of the hoisted digits (the reference keeps two,
        // temp3 / temp3_gs): the baby steps are complete at this point.  The digits are used by one key only, so
        // the fused key switch applies (row stages inside the inner product)
        keyswitch_core(c, u1q.w(), L * N, giant_keys[j], tmp.w(), acc.w(), depth, 1, st, false);
        {
            LaunchScope scope(KC_ELEMENTWISE, st);
            k_pql_add_permute<true><<<g2, 256, 0, st>>>(acc.w(), u.w(), accum.w(), c.d_mod, c.logn, L, depth, pql,
                                                        giant_elts[j]);
        }
        check_launch();
    }
    // final mod-down of both components: PQ_l -> Q_l
    launch_ntt(c, accum.w(), accum.w(), 2 * pql, pl, true, st);
    {
        LaunchScope scope(KC_MODDOWN, st);
        k_moddown_ext<false><<<dim3(c.n >> 8, 2), 256, 0, st>>>(accum.w(), out, 2 * L * N, nullptr, c.d_pc, c.d_half,
                                                             c.d_half_mod, c.d_lqm_pair, 0, c.logn, pql, L, c.Qp, c.Q_size, K, 0);
    }
    check_launch();
    launch_ntt(c, out, out, 2 * L, range_primes(0, L), false, st);
}

} // namespace heon
