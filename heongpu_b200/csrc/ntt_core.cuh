// Register-resident radix-2 stage groups ("rounds") on 16 coefficients/thread.
//
// Twiddle addressing (shared by forward and inverse, reference table order
// psi^bitrev(i), see util.cu:398-451 / ntt_cpu.cu:81-188 of the reference):
// the stage with 2^s butterfly groups uses table[2^s + (j >> (n - s))] for
// coefficient index j.  Inside one pass of S stages that started after s0
// stages on vector number q this becomes, for pass-stage u and vector index
// idx:  table[2^(s0+u) + q*2^u + (idx >> (S-u))].
#pragma once
#include "modarith.cuh"

namespace heon {

__device__ __forceinline__ TwPair ld_tw(const TwPair* p)
{
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(p));
    TwPair r;
    r.w = t.x;
    r.ws = t.y;
    return r;
}

// Cooley-Tukey lazy butterfly: inputs in [0,4p), outputs in [0,4p).
__device__ __forceinline__ void ct_bfly(u64& X, u64& Y, const TwPair& w, u64 p, u64 p2)
{
    u64 x = csub(X, p2);
    u64 t = shoup_mul_lazy(Y, w.w, w.ws, p);
    X = x + t;
    Y = x - t + p2;
}

// Gentleman-Sande lazy butterfly: inputs in [0,2p), outputs in [0,2p).
__device__ __forceinline__ void gs_bfly(u64& X, u64& Y, const TwPair& w, u64 p, u64 p2)
{
    u64 s = csub(X + Y, p2);
    u64 d = X - Y + p2;
    X = s;
    Y = shoup_mul_lazy(d, w.w, w.ws, p);
}

// One stage on the 16 registers; butterflies pair k and k + 2^LS, the
// twiddle changes every 2^(LS+1) registers.
template <int LS, bool INV>
__device__ __forceinline__ void stage16(u64 (&v)[16], const TwPair* __restrict__ tw, u64 p, u64 p2)
{
#pragma unroll
    for (int g = 0; g < (8 >> LS); ++g)
    {
        const TwPair w = ld_tw(tw + g);
#pragma unroll
        for (int j = 0; j < (1 << LS); ++j)
        {
            const int k = g * (2 << LS) + j;
            if (INV)
                gs_bfly(v[k], v[k + (1 << LS)], w, p, p2);
            else
                ct_bfly(v[k], v[k + (1 << LS)], w, p, p2);
        }
    }
}

// Round A: the four stages with register strides 8,4,2,1 when the thread
// holds idx = tt + T*k.  Twiddles do not depend on tt.
__device__ __forceinline__ void ct_round_a(u64 (&v)[16], const TwPair* __restrict__ tw, int s0,
                                           int q, u64 p, u64 p2)
{
    stage16<3, false>(v, tw + (1 << (s0 + 0)) + (q << 0), p, p2);
    stage16<2, false>(v, tw + (1 << (s0 + 1)) + (q << 1), p, p2);
    stage16<1, false>(v, tw + (1 << (s0 + 2)) + (q << 2), p, p2);
    stage16<0, false>(v, tw + (1 << (s0 + 3)) + (q << 3), p, p2);
}

__device__ __forceinline__ void gs_round_a(u64 (&v)[16], const TwPair* __restrict__ tw, int s0,
                                           int q, u64 p, u64 p2)
{
    stage16<0, true>(v, tw + (1 << (s0 + 3)) + (q << 3), p, p2);
    stage16<1, true>(v, tw + (1 << (s0 + 2)) + (q << 2), p, p2);
    stage16<2, true>(v, tw + (1 << (s0 + 1)) + (q << 1), p, p2);
    stage16<3, true>(v, tw + (1 << (s0 + 0)) + (q << 0), p, p2);
}

// Last round of the inverse transform (s0 = 0, q = 0): the final stage
// multiplies both outputs by N^-1 (folded into the twiddle) and canonicalises.
__device__ __forceinline__ void gs_round_a_final(u64 (&v)[16], const TwPair* __restrict__ tw, u64 p,
                                                 u64 p2, const TwPair& ninv, const TwPair& wninv)
{
    stage16<0, true>(v, tw + 8, p, p2);
    stage16<1, true>(v, tw + 4, p, p2);
    stage16<2, true>(v, tw + 2, p, p2);
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        u64 s = v[k] + v[k + 8]; // < 4p
        u64 d = v[k] - v[k + 8] + p2;
        v[k] = csub(shoup_mul_lazy(s, ninv.w, ninv.ws, p), p);
        v[k + 8] = csub(shoup_mul_lazy(d, wninv.w, wninv.ws, p), p);
    }
}

// Round B: the S-4 stages with strides < 16 when the thread holds the 16
// contiguous indices idx = 16*tt + k.
template <int S>
__device__ __forceinline__ void ct_round_b(u64 (&v)[16], const TwPair* __restrict__ tw, int s0,
                                           int q, int tt, u64 p, u64 p2)
{
    // pass-stage u = 4..S-1, register stride 2^(S-1-u)
    if constexpr (S >= 5)
        stage16<S - 5, false>(v, tw + (1 << (s0 + 4)) + (q << 4) + (tt << (8 - S)), p, p2);
    if constexpr (S >= 6)
        stage16<S - 6, false>(v, tw + (1 << (s0 + 5)) + (q << 5) + (tt << (9 - S)), p, p2);
    if constexpr (S >= 7)
        stage16<S - 7, false>(v, tw + (1 << (s0 + 6)) + (q << 6) + (tt << (10 - S)), p, p2);
    if constexpr (S >= 8)
        stage16<S - 8, false>(v, tw + (1 << (s0 + 7)) + (q << 7) + (tt << (11 - S)), p, p2);
}

template <int S>
__device__ __forceinline__ void gs_round_b(u64 (&v)[16], const TwPair* __restrict__ tw, int s0,
                                           int q, int tt, u64 p, u64 p2)
{
    if constexpr (S >= 8)
        stage16<S - 8, true>(v, tw + (1 << (s0 + 7)) + (q << 7) + (tt << (11 - S)), p, p2);
    if constexpr (S >= 7)
        stage16<S - 7, true>(v, tw + (1 << (s0 + 6)) + (q << 6) + (tt << (10 - S)), p, p2);
    if constexpr (S >= 6)
        stage16<S - 6, true>(v, tw + (1 << (s0 + 5)) + (q << 5) + (tt << (9 - S)), p, p2);
    if constexpr (S >= 5)
        stage16<S - 5, true>(v, tw + (1 << (s0 + 4)) + (q << 4) + (tt << (8 - S)), p, p2);
}

} // namespace heon
