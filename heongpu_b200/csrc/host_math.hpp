// Host-side modular arithmetic and parameter search for the B200 RNS engine.
//
// Everything here is one-time setup work (prime search, roots of unity, table
// constants).  Values are defined mathematically (exact residues), which is
// what the reference's host Barrett code produces for canonical operands:
//   reference Modulus64 record   thirdparty/GPU-NTT/src/include/gpuntt/common/modular_arith.cuh:28-60
//   reference prime search       src/lib/util/util.cu:219-276
//   reference minimal psi        src/lib/util/util.cu:312-380
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace heon {

typedef uint64_t u64;
typedef unsigned __int128 u128;

// Same 24-byte layout as the reference's Modulus64 {value, bit, mu}.
struct Mod64 {
    u64 value;
    u64 bit;
    u64 mu;
};

inline int bit_length(u64 v) { return v ? 64 - __builtin_clzll(v) : 0; }

inline Mod64 make_mod(u64 p)
{
    Mod64 m;
    m.value = p;
    m.bit = (u64) bit_length(p);
    m.mu = (u64) ((((u128) 1) << (2 * m.bit + 1)) / p);
    return m;
}

inline u64 mulmod(u64 a, u64 b, u64 p) { return (u64) (((u128) a * b) % p); }

// The reference's HOST Barrett product, operation for operation
// (gpuntt/common/modular_arith.cuh:90-107, OPERATOR<Data64>::mult): all intermediates in 128 bits, ONE
// conditional subtraction.  For operands below the modulus it is the exact residue; the Method-II table
// generators (contextpool.cpp:160-191, 193-236, 361-438) call it with an UNREDUCED prime as one operand,
// and then the word it leaves can be a non-canonical representative.  Those generators are reproduced with
// this function so that every exported table word equals the reference's on every modulus chain.
inline u64 barrett_mult_host(u64 a, u64 b, const Mod64& m)
{
    u128 mult = (u128) a * (u128) b;
    u128 r = mult >> (m.bit - 2);
    r = r * (u128) m.mu;
    r = r >> (m.bit + 3);
    r = r * (u128) m.value;
    mult = mult - r;
    const u64 res = (u64) mult;
    return res >= m.value ? res - m.value : res;
}
// OPERATOR<Data64>::exp / modinv of the reference (modular_arith.cuh:111-136): square-and-multiply over
// the host Barrett product, bit count from log2() of the exponent as a double.  Equal to the exact inverse
// for a canonical base; reproduced so that a non-canonical base gives the reference's word.
inline u64 barrett_exp_host(u64 base, u64 exponent, const Mod64& m)
{
    u64 result = 1;
    if (exponent == 0)
        return result;
    const int exponent_bit = (int) (__builtin_log2((double) exponent) + 1);
    for (int i = exponent_bit - 1; i >= 0; --i)
    {
        result = barrett_mult_host(result, result, m);
        if (i < 64 && ((exponent >> i) & 1))
            result = barrett_mult_host(result, base, m);
    }
    return result;
}
inline u64 barrett_modinv_host(u64 a, const Mod64& m) { return barrett_exp_host(a, m.value - 2, m); }

inline u64 addmod(u64 a, u64 b, u64 p)
{
    u64 s = a + b;
    return s >= p ? s - p : s;
}
inline u64 submod(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }

inline u64 powmod(u64 b, u64 e, u64 p)
{
    u64 r = 1 % p;
    b %= p;
    while (e)
    {
        if (e & 1)
            r = mulmod(r, b, p);
        b = mulmod(b, b, p);
        e >>= 1;
    }
    return r;
}
inline u64 invmod(u64 a, u64 p) { return powmod(a, p - 2, p); }

// floor(w * 2^64 / p): the Shoup companion word of a constant multiplier w<p.
inline u64 shoup(u64 w, u64 p) { return (u64) ((((u128) w) << 64) / p); }

// Deterministic Miller-Rabin, exact for all 64-bit inputs.
inline bool is_prime_u64(u64 n)
{
    if (n < 2)
        return false;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 s : small)
    {
        if (n == s)
            return true;
        if (n % s == 0)
            return false;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1))
    {
        d >>= 1;
        ++r;
    }
    for (u64 a : small)
    {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1)
            continue;
        bool comp = true;
        for (int i = 1; i < r; ++i)
        {
            x = mulmod(x, x, n);
            if (x == n - 1)
            {
                comp = false;
                break;
            }
        }
        if (comp)
            return false;
    }
    return true;
}

// The `count` largest primes of exactly `bits` bits that are 1 mod `factor`,
// in descending order.
inline std::vector<u64> largest_ntt_primes(u64 factor, int bits, size_t count)
{
    std::vector<u64> out;
    u64 v = ((((u64) 1) << bits) - 1) / factor * factor + 1;
    u64 lo = ((u64) 1) << (bits - 1);
    while (out.size() < count && v > lo)
    {
        if (is_prime_u64(v))
            out.push_back(v);
        v -= factor;
    }
    if (out.size() < count)
        throw std::logic_error("failed to find enough qualifying primes");
    return out;
}

// Prime chain for a list of bit sizes: each size class takes the largest
// primes of that size; within a class the list is handed out from its back
// (smallest first), in the order the sizes appear.
inline std::vector<u64> primes_for_bit_sizes(u64 n, const std::vector<int>& bits)
{
    std::vector<u64> out;
    std::vector<int> classes;
    std::vector<std::vector<u64>> pools;
    for (int b : bits)
    {
        if (b > 60 || b < 30)
            throw std::logic_error("invalid modulus bit size");
        bool seen = false;
        for (int c : classes)
            seen |= (c == b);
        if (!seen)
        {
            size_t cnt = 0;
            for (int b2 : bits)
                cnt += (b2 == b);
            classes.push_back(b);
            pools.push_back(largest_ntt_primes(2 * n, b, cnt));
        }
    }
    for (int b : bits)
    {
        for (size_t c = 0; c < classes.size(); ++c)
            if (classes[c] == b)
            {
                out.push_back(pools[c].back());
                pools[c].pop_back();
            }
    }
    return out;
}

// Smallest primitive `degree`-th root of unity mod p (degree a power of two).
inline u64 minimal_primitive_root(u64 degree, u64 p)
{
    if ((p - 1) % degree)
        throw std::logic_error("no sufficient root unity");
    u64 cof = (p - 1) / degree;
    u64 root = 0;
    for (u64 g = 2; g < 1000; ++g)
    {
        u64 r = powmod(g, cof, p);
        if (powmod(r, degree >> 1, p) == p - 1)
        {
            root = r;
            break;
        }
    }
    if (!root)
        throw std::logic_error("no sufficient root unity");
    // all primitive roots are the odd powers of any one of them
    u64 sq = mulmod(root, root, p);
    u64 cur = root, best = root;
    for (u64 i = 0; i < degree; i += 2)
    {
        if (cur < best)
            best = cur;
        cur = mulmod(cur, sq, p);
    }
    return best;
}

inline uint32_t bitrev(uint32_t x, int bits)
{
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i)
    {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

} // namespace heon
