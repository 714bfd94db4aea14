// Device modular arithmetic for 64-bit RNS limbs (p < 2^62).
//
// Two families:
//  * barrett_*: bit-for-bit the reference's Barrett sequence
//      (thirdparty/GPU-NTT/src/include/gpuntt/common/modular_arith.cuh:312-339,343-369,409-418)
//      mu = floor(2^(2*bit+1)/p), q = ((z >> (bit-2)) * mu) >> (bit+3), one
//      conditional subtract.  Used by the element-wise kernels so that even
//      non-canonical operands behave exactly like the reference.
//  * shoup_*: constant-multiplier products with a precomputed companion word
//      ws = floor(w*2^64/p); results are lazy (in [0,2p)).  Used inside the
//      NTT butterflies; every value is canonicalised before it is stored, so
//      stored words equal the reference's for canonical inputs.
#pragma once
#include <cstdint>
#include "heon_internal.hpp"

namespace heon {

__device__ __forceinline__ u64 mod_add(u64 a, u64 b, u64 p)
{
    u64 s = a + b;
    return (s >= p) ? (s - p) : s;
}

__device__ __forceinline__ u64 mod_sub(u64 a, u64 b, u64 p)
{
    u64 d = a + p - b;
    return (d >= p) ? (d - p) : d;
}

// low 64 bits of ((hi:lo) >> s), 0 < s <= 64
__device__ __forceinline__ u64 shr128_lo(u64 lo, u64 hi, unsigned s)
{
    return (s >= 64) ? hi : ((lo >> s) | (hi << (64 - s)));
}

__device__ __forceinline__ u64 barrett_fold(u64 zlo, u64 zhi, const Mod64& m)
{
    u64 w = shr128_lo(zlo, zhi, (unsigned) m.bit - 2);
    u64 wl = w * m.mu;
    u64 wh = __umul64hi(w, m.mu);
    w = shr128_lo(wl, wh, (unsigned) m.bit + 3);
    u64 r = zlo - w * m.value; // low word of z - w*p
    return (r >= m.value) ? (r - m.value) : r;
}

__device__ __forceinline__ u64 barrett_mul(u64 a, u64 b, const Mod64& m)
{
    return barrett_fold(a * b, __umul64hi(a, b), m);
}

__device__ __forceinline__ u64 barrett_reduce(u64 a, const Mod64& m)
{
    return barrett_fold(a, 0, m);
}

// reference reduce_forced: iterate until canonical
__device__ __forceinline__ u64 reduce_forced(u64 a, const Mod64& m)
{
    u64 r = a;
    while (r >= m.value)
        r = barrett_reduce(r, m);
    return r;
}

// x*w mod p, lazily: result in [0,2p) for ANY 64-bit x (w < p).
__device__ __forceinline__ u64 shoup_mul_lazy(u64 x, u64 w, u64 ws, u64 p)
{
    u64 q = __umul64hi(x, ws);
    return x * w - q * p;
}

__device__ __forceinline__ u64 csub(u64 x, u64 p) { return (x >= p) ? (x - p) : x; }


// exact x mod p for any 64-bit x
__device__ __forceinline__ u64 reduce_u64(u64 x, const PrimeConst& c)
{
    u64 q = __umul64hi(x, c.inv64);
    return csub(x - q * c.p, c.p);
}

// exact (hi*2^64 + lo) mod p
__device__ __forceinline__ u64 reduce_u128(u64 lo, u64 hi, const PrimeConst& c)
{
    u64 h = reduce_u64(hi, c);
    u64 t = shoup_mul_lazy(h, c.r64, c.r64s, c.p); // [0,2p)
    u64 ql = __umul64hi(lo, c.inv64);
    u64 l = lo - ql * c.p; // [0,2p)
    u64 s = t + l; // [0,4p), fits because p < 2^62
    s = csub(s, 2 * c.p);
    return csub(s, c.p);
}

// 128-bit lazy accumulator: acc += a*b
__device__ __forceinline__ void mac128(u64& lo, u64& hi, u64 a, u64 b)
{
    u64 pl = a * b;
    u64 ph = __umul64hi(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

} // namespace heon
