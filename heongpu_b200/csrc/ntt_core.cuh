// Register-resident radix-2 stage groups ("rounds") on 16 coefficients/thread.
//
// Twiddle addressing (shared by forward and inverse, reference table order
// psi^bitrev(i), see util.cu:398-451 / ntt_cpu.cu:81-188 of the reference):
// the stage with 2^s butterfly groups uses table[2^s + (j >> (n - s))] for
// coefficient index j.  Inside one pass of S stages that started after s0
// stages on vector number q this becomes, for pass-stage u and vector index
// idx:  table[2^(s0+u) + q*2^u + (idx >> (S-u))].
//
// Lazy-reduction variants of the forward (Cooley-Tukey) butterfly.  VAR 1/2
// use the Shoup product T = Y*w - q*p with an APPROXIMATE quotient
// q in [floor(Y*ws/2^64) - 2, floor(Y*ws/2^64)] (three 32x32->64 multiplies
// instead of four), so T is in [0,4p):
//   VAR 0  Harvey, exact quotient: values in [0,4p), one conditional
//          subtraction of 2p per butterfly.
//   VAR 1  values in [0,8p), one conditional subtraction of 4p per butterfly;
//          valid for every p < 2^61.
//   VAR 2  no per-stage correction at all: every stage adds at most 4p to the
//          bound, 4p + 16*4p = 68p < 2^64 needs p <= 57 bits; the single
//          final reduction uses a 32-bit quotient estimate (PrimeConst::fin_m).
//   VAR 3  the whole butterfly on the FP64 pipe (fp_mulmod below): words are
//          integer-valued doubles in balanced form, eight DFMA-class
//          instructions per butterfly and no integer multiply at all; no
//          per-stage correction, |v| grows by ~p/2 per stage and must stay
//          below 2^51: p <= 47 bits.
//   VAR 4  the same with X reduced on every other stage (three more FP64
//          instructions), which keeps |v| < 1.7p < 2^51: p <= 50 bits.
// For VAR 3/4 a twiddle pair holds the doubles {w, RN(w/p)}; the buffer
// between the column pass and the row pass holds the doubles' bit patterns.
// Stored words are always canonical, so all variants give identical results.
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

// floor(a*b / 2^64) minus at most 2: drops the low x low partial product and
// the carry out of the low halves of the two cross products.
__device__ __forceinline__ u64 mulhi_approx(u64 a, u64 b)
{
    const unsigned a0 = (unsigned) a, a1 = (unsigned) (a >> 32);
    const unsigned b0 = (unsigned) b, b1 = (unsigned) (b >> 32);
    const u64 m1 = (u64) a0 * b1;
    const u64 m2 = (u64) a1 * b0;
    return (u64) a1 * b1 + ((m1 >> 32) + (m2 >> 32));
}

// x*w mod p, lazily in [0,4p), for ANY 64-bit x.
__device__ __forceinline__ u64 shoup_mul_lazy3(u64 x, u64 w, u64 ws, u64 p)
{
    const u64 q = mulhi_approx(x, ws);
    return x * w - q * p;
}

struct BflyConst {
    u64 p, p2, p4;
    u64 np; // 2^64 - p
    double dp, dnp, dpinv, dpinv_lo; // p, -p, RN(1/p), RN(1/p - RN(1/p)) for the FP64 variants
};

__device__ __forceinline__ BflyConst make_bc(const PrimeConst& pc)
{
    BflyConst c;
    c.p = pc.p;
    c.p2 = 2 * pc.p;
    c.p4 = 4 * pc.p;
    c.np = 0 - pc.p;
    c.dp = (double) pc.p;
    c.dnp = -c.dp;
    c.dpinv = pc.pinv;
    c.dpinv_lo = pc.pinv_lo;
    return c;
}

// Hand-scheduled instruction selection for the hot butterflies.  ptxas, left
// to itself, turns 64-bit additions into IMAD.WIDE (a*1+c) and so loads the
// multiplier pipe -- the bottleneck of this kernel -- with work the integer
// ALU can do.  Written with 32-bit carry-chain adds, the multiplier pipe only
// sees the nine real multiplies of a Shoup product:
//   q  ~ hi64(Y*ws)        2 x mul.hi.u32 + 1 x mul.wide.u32   (q >= exact - 2)
//   T  = lo64(Y*w) + lo64(q*(2^64-p))   2 x mad.wide.u32 + 4 x mad.lo.u32
// T is in [0,4p) for ANY 64-bit Y.
__device__ __forceinline__ u64 shoup_lazy_ptx(u64 y, u64 w, u64 ws, u64 np)
{
#ifndef __CUDA_ARCH__
    return y * w + mulhi_approx(y, ws) * np; // host emulation (tests/host_emul.cpp)
#else
    u64 t;
    asm("{\n\t"
        ".reg .u32 y0,y1,w0,w1,s0,s1,n0,n1,h1,h2,q0,q1,t0,t1;\n\t"
        ".reg .u64 q, tt;\n\t"
        "mov.b64 {y0,y1}, %1;\n\t"
        "mov.b64 {w0,w1}, %2;\n\t"
        "mov.b64 {s0,s1}, %3;\n\t"
        "mov.b64 {n0,n1}, %4;\n\t"
        "mul.hi.u32 h1, y0, s1;\n\t"
        "mul.hi.u32 h2, y1, s0;\n\t"
        "mul.wide.u32 q, y1, s1;\n\t"
        "mov.b64 {q0,q1}, q;\n\t"
        "add.cc.u32 q0, q0, h1;\n\t"
        "addc.u32 q1, q1, 0;\n\t"
        "add.cc.u32 q0, q0, h2;\n\t"
        "addc.u32 q1, q1, 0;\n\t"
        "mul.wide.u32 tt, y0, w0;\n\t"
        "mad.wide.u32 tt, q0, n0, tt;\n\t"
        "mov.b64 {t0,t1}, tt;\n\t"
        "mad.lo.u32 t1, y0, w1, t1;\n\t"
        "mad.lo.u32 t1, y1, w0, t1;\n\t"
        "mad.lo.u32 t1, q0, n1, t1;\n\t"
        "mad.lo.u32 t1, q1, n0, t1;\n\t"
        "mov.b64 %0, {t0,t1};\n\t"
        "}"
        : "=l"(t)
        : "l"(y), "l"(w), "l"(ws), "l"(np));
    return t;
#endif
}

// X' = X + T,  Y' = X + C - T   with 32-bit carry chains (integer ALU only)
__device__ __forceinline__ void addsub_ptx(u64& X, u64& Y, u64 T, u64 C)
{
#ifndef __CUDA_ARCH__
    const u64 x = X;
    X = x + T;
    Y = x + C - T;
#else
    u64 xo, yo;
    asm("{\n\t"
        ".reg .u32 x0,x1,t0,t1,c0,c1,a0,a1,b0,b1;\n\t"
        "mov.b64 {x0,x1}, %2;\n\t"
        "mov.b64 {t0,t1}, %3;\n\t"
        "mov.b64 {c0,c1}, %4;\n\t"
        "add.cc.u32 a0, x0, t0;\n\t"
        "addc.u32 a1, x1, t1;\n\t"
        "add.cc.u32 b0, x0, c0;\n\t"
        "addc.u32 b1, x1, c1;\n\t"
        "sub.cc.u32 b0, b0, t0;\n\t"
        "subc.u32 b1, b1, t1;\n\t"
        "mov.b64 %0, {a0,a1};\n\t"
        "mov.b64 %1, {b0,b1};\n\t"
        "}"
        : "=l"(xo), "=l"(yo)
        : "l"(X), "l"(T), "l"(C));
    X = xo;
    Y = yo;
#endif
}

// ---------------------------------------------------------------------------
// FP64-pipe modular arithmetic (primes below 2^50).
//
// The multiplier pipe is the bottleneck of the integer butterfly (IMAD.WIDE
// issues once per ~4 cycles per SM sub-partition on B200) while the FP64 pipe
// (64 DFMA/clk/SM) idles.  Words are kept as integer-valued doubles in
// balanced form; with Y, w integers, |Y| < 2^51, 0 <= w < p < 2^50:
//   q  = rint(Y * RN(w/p))              fma against 1.5*2^52, one rounding
//   h  = RN(Y*w),  l = fma(Y, w, -h)    l = Y*w - h exactly (error-free product)
//   r  = fma(q, -p, h)                  h - q*p is an integer below 2^51: exact
//   T  = r + l = Y*w - q*p              exact, |T| <= p*(1/2 + |Y|*2^-54)
// Five FP64 instructions and no integer multiply; the butterfly adds two more
// (X + T, X - T).  Every step is exact integer arithmetic once q is fixed, so
// canonical results equal the integer path's bit for bit.
// ---------------------------------------------------------------------------
#define HEON_FP_MAGIC 6755399441055744.0 /* 1.5 * 2^52 */

// Alternative form of the quotient (compile with -DHEON_FP_FRND=1): q = rint(RN(Y * winv)) with the rounding
// done by cvt.rni.f64.f64 (FRND), which does not issue on the FP64 pipe -- seven FP64-pipe instructions per
// butterfly instead of eight.  Register-resident it measures 15.5 instead of 16.7 SM sub-partition cycles per
// warp-butterfly (profiles/r2_microbench4.txt), but inside the real kernels it gained nothing (column pass
// 93 instead of 89 us/op, fused row pass + inner product unchanged at 114 us/op at C3-II: the conversion
// unit becomes the next limiter), so the magic-constant rounding stays the default.  Bounds of the FRND form,
// kept because the host emulation (tests/host_emul.cpp) covers both: the product is rounded before the
// integer rounding, so for |Y| < 2^52
//   |q - Y*w/p| <= 1/2 + ulp(Y*winv)/2 + |Y|*2^-54 <= 1/2 + 1/4 + 1/4   ->   |T| <= p,
// and <= 0.75 p while |Y| < 2^51.  Every later step is exact as before (h - q*p is an integer below 2^52,
// l is the error-free remainder), values stay below 2^52 (VAR 4: at most 2.25 p, tests/host_emul.cpp), and
// canonical results are unchanged bit for bit (153 GPU parity tests passed with it).
#ifndef HEON_FP_FRND
#define HEON_FP_FRND 0
#endif
__device__ __forceinline__ double fp_rint(double x)
{
#ifndef __CUDA_ARCH__
    return nearbyint(x); // host emulation (round to nearest even, like cvt.rni)
#else
    double r;
    asm("cvt.rni.f64.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#endif
}

__device__ __forceinline__ double u2d(u64 x) { return __longlong_as_double((long long) x); }
__device__ __forceinline__ u64 d2u(double x) { return (u64) __double_as_longlong(x); }

__device__ __forceinline__ double fp_mulmod(double y, double w, double winv, double np)
{
#ifdef HEON_FP_TRACK
    heon_fp_track(y); // host emulation only: records max |Y| to check the 2^51 operand bound
#endif
#if HEON_FP_FRND
    const double q = fp_rint(__dmul_rn(y, winv));
#else
    const double q = __dsub_rn(__fma_rn(y, winv, HEON_FP_MAGIC), HEON_FP_MAGIC);
#endif
    const double h = __dmul_rn(y, w);
    const double l = __fma_rn(y, w, -h);
    const double r = __fma_rn(q, np, h);
    return __dadd_rn(r, l);
}

// x - rint(x/p)*p: |result| <= p/2 (+1), exact for |x| < 2^52
__device__ __forceinline__ double fp_reduce(double x, double pinv, double np)
{
#if HEON_FP_FRND
    const double q = fp_rint(__dmul_rn(x, pinv));
#else
    const double q = __dsub_rn(__fma_rn(x, pinv, HEON_FP_MAGIC), HEON_FP_MAGIC);
#endif
    return __fma_rn(q, np, x);
}

// integer word x < 2^52 -> double (exact)
__device__ __forceinline__ double fp_from_u64(u64 x)
{
    return __dsub_rn(__hiloint2double((int) (0x43300000u | (unsigned) (x >> 32)), (int) (unsigned) x),
                     4503599627370496.0);
}

// integer-valued double |v| < 2^52 -> canonical residue in [0,p), as an integer word
__device__ __forceinline__ u64 fp_canon(double v, double pinv, double np, double dp)
{
    double r = fp_reduce(v, pinv, np); // [-p/2, p/2]
    if (r < 0.0)
        r = __dadd_rn(r, dp);
    // r in [0,p), p < 2^50: the low 52 bits of r + 2^52 are the integer
    return d2u(__dadd_rn(r, 4503599627370496.0)) & 0x000FFFFFFFFFFFFFull;
}

// Word as loaded (an integer below 4p) -> working representation of variant VAR.
// `lazy`: the word may exceed p (fused mod-up); VAR 4 needs |v| <= p on entry.
template <int VAR> __device__ __forceinline__ u64 ct_prep(u64 x, const BflyConst& c, bool lazy)
{
    if (VAR < 3)
        return x;
    double d = fp_from_u64(x);
    if (VAR == 4 && lazy)
        d = fp_reduce(d, c.dpinv, c.dnp);
    return d2u(d);
}

template <int VAR, bool RED = false>
__device__ __forceinline__ void ct_bfly(u64& X, u64& Y, const TwPair& w, const BflyConst& c)
{
    if (VAR == 0)
    {
        const u64 x = csub(X, c.p2);
        const u64 t = shoup_mul_lazy(Y, w.w, w.ws, c.p);
        X = x + t;
        Y = x - t + c.p2;
    }
    else if (VAR == 1)
    {
        X = csub(X, c.p4);
        const u64 t = shoup_lazy_ptx(Y, w.w, w.ws, c.np);
        addsub_ptx(X, Y, t, c.p4);
    }
    else if (VAR == 2)
    {
        const u64 t = shoup_lazy_ptx(Y, w.w, w.ws, c.np);
        addsub_ptx(X, Y, t, c.p4);
    }
    else
    {
        // VAR 3 / 4: FP64 pipe only (see fp_mulmod); VAR 4 reduces X on the stages marked RED
        double x = u2d(X);
        if (VAR == 4 && RED)
            x = fp_reduce(x, c.dpinv, c.dnp);
        const double t = fp_mulmod(u2d(Y), u2d(w.w), u2d(w.ws), c.dnp);
        X = d2u(__dadd_rn(x, t));
        Y = d2u(__dsub_rn(x, t));
    }
}

// canonical word loaded by the first pass of the inverse transform -> working representation
template <int GVAR> __device__ __forceinline__ u64 gs_prep(u64 x)
{
    return GVAR >= 3 ? d2u(fp_from_u64(x)) : x;
}

// canonical value of a lazy forward word
template <int VAR> __device__ __forceinline__ u64 ct_finish(u64 x, const BflyConst& c, const PrimeConst& pc)
{
    if (VAR == 3 || VAR == 4)
        return fp_canon(u2d(x), c.dpinv, c.dnp, c.dp);
    if (VAR == 0)
        return csub(csub(x, c.p2), c.p);
    if (VAR == 1)
        return csub(csub(csub(x, c.p4), c.p2), c.p);
    // x < 128p: 32-bit quotient estimate, remainder in [0,2p)
    const unsigned xt = (unsigned) (x >> pc.fin_shift);
    const unsigned q = (unsigned) (((u64) xt * pc.fin_m) >> 56);
    return csub(x - (u64) q * c.p, c.p);
}

// Gentleman-Sande lazy butterfly.  GVAR 0: values in [0,2p), exact quotient.
// GVAR 1: values in [0,4p), approximate quotient.
// GVAR 3 / 4: the FP64 pipe (words are integer-valued doubles, twiddle pairs {w, RN(w/p)}):
//   S = X + Y (reduced when RED), D = X - Y, Y' = D*w mod p by fp_mulmod.  The sum branch doubles
//   per stage, so it is reduced on every stage for p < 2^50 (GVAR 4: all words stay below 0.58p,
//   |D| < 1.2p < 2^51) and on every other stage for p < 2^47 (GVAR 3).
template <int GVAR, bool RED = true>
__device__ __forceinline__ void gs_bfly(u64& X, u64& Y, const TwPair& w, const BflyConst& c)
{
    if (GVAR >= 3)
    {
        const double x = u2d(X), y = u2d(Y);
        double s = __dadd_rn(x, y);
        const double d = __dsub_rn(x, y);
        if (GVAR == 4 || RED)
            s = fp_reduce(s, c.dpinv, c.dnp);
        X = d2u(s);
        Y = d2u(fp_mulmod(d, u2d(w.w), u2d(w.ws), c.dnp));
    }
    else if (GVAR == 0)
    {
        const u64 s = csub(X + Y, c.p2);
        const u64 d = X - Y + c.p2;
        X = s;
        Y = shoup_mul_lazy(d, w.w, w.ws, c.p);
    }
    else
    {
        u64 s = X, d = Y;
        addsub_ptx(s, d, Y, c.p4); // s = X + Y, d = X + 4p - Y
        X = csub(s, c.p4);
        Y = shoup_lazy_ptx(d, w.w, w.ws, c.np);
    }
}

// One stage on the 16 registers; butterflies pair k and k + 2^LS, the
// twiddle changes every 2^(LS+1) registers.
// SMTW: the twiddles were staged in shared memory (plain loads instead of ld.global.nc)
template <int LS, bool INV, int VAR, int GSTRIDE = 1, bool RED = false, bool SMTW = false>
__device__ __forceinline__ void stage16(u64 (&v)[16], const TwPair* __restrict__ tw, const BflyConst& c)
{
#pragma unroll
    for (int g = 0; g < (8 >> LS); ++g)
    {
        TwPair w;
        if constexpr (SMTW)
        {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(tw + g * GSTRIDE);
            w.w = t.x;
            w.ws = t.y;
        }
        else
            w = ld_tw(tw + g * GSTRIDE);
#pragma unroll
        for (int j = 0; j < (1 << LS); ++j)
        {
            const int k = g * (2 << LS) + j;
#ifdef HEON_NTT_NULL
            // diagnostic build: keep every load/store/transposition, drop the arithmetic
            v[k] ^= w.w;
            v[k + (1 << LS)] ^= w.ws;
            continue;
#endif
            if (INV)
                gs_bfly<VAR, RED>(v[k], v[k + (1 << LS)], w, c);
            else
                ct_bfly<VAR, RED>(v[k], v[k + (1 << LS)], w, c);
        }
    }
}

// FP64 row pass with the twiddles of the row staged in shared memory as bare doubles w
// (8 bytes instead of a 16-byte pair: half the L2 traffic and half the shared memory).  The
// quotient multiplier is rebuilt as winv = RN(RN(w*ph) + w*pl) with ph + pl = 1/p to 106 bits:
// |winv - w/p| <= 2^-53, so |T| <= p*(1/2 + |Y|*2^-53) -- with X reduced on every other stage
// (VAR 4, pattern R,-,R,-) values stay below 1.9p < 2^51 for p < 2^50; VAR 3 (p < 2^47) stays
// below 14p.  Row layout: [0..14] round-A entries (stage u: 2^u - 1 + g), [15] pad,
// [16 + e*16 + lane] round-B entries (lane-major, as ct_round_b_lm).
template <int LS, int VAR, int GSTRIDE, bool RED>
__device__ __forceinline__ void stage16_sm(u64 (&v)[16], const double* tws, const BflyConst& c)
{
#pragma unroll
    for (int g = 0; g < (8 >> LS); ++g)
    {
        const double wd = tws[g * GSTRIDE];
        TwPair w;
        w.w = d2u(wd);
        w.ws = d2u(__fma_rn(wd, c.dpinv_lo, __dmul_rn(wd, c.dpinv)));
#pragma unroll
        for (int j = 0; j < (1 << LS); ++j)
        {
            const int k = g * (2 << LS) + j;
            ct_bfly<VAR, RED>(v[k], v[k + (1 << LS)], w, c);
        }
    }
}
template <int VAR> __device__ __forceinline__ void ct_round_a_sm(u64 (&v)[16], const double* rowtw, const BflyConst& c)
{
    stage16_sm<3, VAR, 1, true>(v, rowtw + 0, c);
    stage16_sm<2, VAR, 1, false>(v, rowtw + 1, c);
    stage16_sm<1, VAR, 1, true>(v, rowtw + 3, c);
    stage16_sm<0, VAR, 1, false>(v, rowtw + 7, c);
}
template <int VAR>
__device__ __forceinline__ void ct_round_b_sm(u64 (&v)[16], const double* rowtw, int tt, const BflyConst& c)
{
    stage16_sm<3, VAR, 16, true>(v, rowtw + 16 + 0 * 16 + tt, c);
    stage16_sm<2, VAR, 16, false>(v, rowtw + 16 + 1 * 16 + tt, c);
    stage16_sm<1, VAR, 16, true>(v, rowtw + 16 + 3 * 16 + tt, c);
    stage16_sm<0, VAR, 16, false>(v, rowtw + 16 + 7 * 16 + tt, c);
}

// Round A: the four stages with register strides 8,4,2,1 when the thread
// holds idx = tt + T*k.  Twiddles do not depend on tt.
// PH: VAR 4 reduces X on every other stage; PH = 0 gives the pattern -,R,-,R (column pass, whose
// input is canonical), PH = 1 gives R,-,R,- (row pass: its first stage follows an unreduced or
// reduced column-pass stage alike).
template <int VAR, int PH = 0, bool SMTW = false>
__device__ __forceinline__ void ct_round_a(u64 (&v)[16], const TwPair* __restrict__ tw, int s0, int q,
                                           const BflyConst& c)
{
    stage16<3, false, VAR, 1, PH == 1, SMTW>(v, tw + (1 << (s0 + 0)) + (q << 0), c);
    stage16<2, false, VAR, 1, PH == 0, SMTW>(v, tw + (1 << (s0 + 1)) + (q << 1), c);
    stage16<1, false, VAR, 1, PH == 1, SMTW>(v, tw + (1 << (s0 + 2)) + (q << 2), c);
    stage16<0, false, VAR, 1, PH == 0, SMTW>(v, tw + (1 << (s0 + 3)) + (q << 3), c);
}

template <int GVAR>
__device__ __forceinline__ void gs_round_a(u64 (&v)[16], const TwPair* __restrict__ tw, int s0, int q,
                                           const BflyConst& c)
{
    // RED (FP64 variants only): the sum branch is reduced on alternate stages (GVAR 4: on every stage)
    stage16<0, true, GVAR, 1, true>(v, tw + (1 << (s0 + 3)) + (q << 3), c);
    stage16<1, true, GVAR, 1, false>(v, tw + (1 << (s0 + 2)) + (q << 2), c);
    stage16<2, true, GVAR, 1, true>(v, tw + (1 << (s0 + 1)) + (q << 1), c);
    stage16<3, true, GVAR, 1, false>(v, tw + (1 << (s0 + 0)) + (q << 0), c);
}

// Last round of the inverse transform (s0 = 0, q = 0): the final stage
// multiplies both outputs by N^-1 (folded into the twiddle) and canonicalises.
template <int GVAR>
__device__ __forceinline__ void gs_round_a_final(u64 (&v)[16], const TwPair* __restrict__ tw,
                                                 const BflyConst& c, const TwPair& ninv,
                                                 const TwPair& wninv)
{
    stage16<0, true, GVAR, 1, true>(v, tw + 8, c);
    stage16<1, true, GVAR, 1, false>(v, tw + 4, c);
    stage16<2, true, GVAR, 1, true>(v, tw + 2, c);
    if (GVAR >= 3)
    {
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            const double x = u2d(v[k]), y = u2d(v[k + 8]);
            const double sd = __dadd_rn(x, y), dd = __dsub_rn(x, y); // |.| < 2.4p
            v[k] = fp_canon(fp_mulmod(sd, u2d(ninv.w), u2d(ninv.ws), c.dnp), c.dpinv, c.dnp, c.dp);
            v[k + 8] = fp_canon(fp_mulmod(dd, u2d(wninv.w), u2d(wninv.ws), c.dnp), c.dpinv, c.dnp, c.dp);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        const u64 s = v[k] + v[k + 8]; // < 8p
        const u64 d = v[k] - v[k + 8] + c.p4;
        if (GVAR == 0)
        {
            v[k] = csub(shoup_mul_lazy(s, ninv.w, ninv.ws, c.p), c.p);
            v[k + 8] = csub(shoup_mul_lazy(d, wninv.w, wninv.ws, c.p), c.p);
        }
        else
        {
            v[k] = csub(csub(shoup_lazy_ptx(s, ninv.w, ninv.ws, c.np), c.p2), c.p);
            v[k + 8] = csub(csub(shoup_lazy_ptx(d, wninv.w, wninv.ws, c.np), c.p2), c.p);
        }
    }
}

// Round B: the S-4 stages with strides < 16 when the thread holds the 16
// contiguous indices idx = 16*tt + k.
template <int S, int VAR, int PH = 0, bool SMTW = false>
__device__ __forceinline__ void ct_round_b(u64 (&v)[16], const TwPair* __restrict__ tw, int s0, int q,
                                           int tt, const BflyConst& c)
{
    // pass-stage u = 4..S-1, register stride 2^(S-1-u)
    if constexpr (S >= 5)
        stage16<S - 5, false, VAR, 1, PH == 1, SMTW>(v, tw + (1 << (s0 + 4)) + (q << 4) + (tt << (8 - S)), c);
    if constexpr (S >= 6)
        stage16<S - 6, false, VAR, 1, PH == 0, SMTW>(v, tw + (1 << (s0 + 5)) + (q << 5) + (tt << (9 - S)), c);
    if constexpr (S >= 7)
        stage16<S - 7, false, VAR, 1, PH == 1, SMTW>(v, tw + (1 << (s0 + 6)) + (q << 6) + (tt << (10 - S)), c);
    if constexpr (S >= 8)
        stage16<S - 8, false, VAR, 1, PH == 0, SMTW>(v, tw + (1 << (s0 + 7)) + (q << 7) + (tt << (11 - S)), c);
}

// Round B of the 256-point row transform with the lane-major twiddle block of
// this (prime,row): entry e = 2^(u-4)-1+g of lane tt sits at blk[e*16 + tt], so
// the 16 lanes of a row read 256 contiguous bytes per butterfly group.
template <int VAR, int PH = 0>
__device__ __forceinline__ void ct_round_b_lm(u64 (&v)[16], const TwPair* __restrict__ blk, int tt,
                                              const BflyConst& c)
{
    stage16<3, false, VAR, 16, PH == 1>(v, blk + 0 * 16 + tt, c);
    stage16<2, false, VAR, 16, PH == 0>(v, blk + 1 * 16 + tt, c);
    stage16<1, false, VAR, 16, PH == 1>(v, blk + 3 * 16 + tt, c);
    stage16<0, false, VAR, 16, PH == 0>(v, blk + 7 * 16 + tt, c);
}
template <int GVAR>
__device__ __forceinline__ void gs_round_b_lm(u64 (&v)[16], const TwPair* __restrict__ blk, int tt,
                                              const BflyConst& c)
{
    stage16<0, true, GVAR, 16, true>(v, blk + 7 * 16 + tt, c);
    stage16<1, true, GVAR, 16, false>(v, blk + 3 * 16 + tt, c);
    stage16<2, true, GVAR, 16, true>(v, blk + 1 * 16 + tt, c);
    stage16<3, true, GVAR, 16, false>(v, blk + 0 * 16 + tt, c);
}

template <int S, int GVAR>
__device__ __forceinline__ void gs_round_b(u64 (&v)[16], const TwPair* __restrict__ tw, int s0, int q,
                                           int tt, const BflyConst& c)
{
    if constexpr (S >= 8)
        stage16<S - 8, true, GVAR, 1, true>(v, tw + (1 << (s0 + 7)) + (q << 7) + (tt << (11 - S)), c);
    if constexpr (S >= 7)
        stage16<S - 7, true, GVAR, 1, false>(v, tw + (1 << (s0 + 6)) + (q << 6) + (tt << (10 - S)), c);
    if constexpr (S >= 6)
        stage16<S - 6, true, GVAR, 1, true>(v, tw + (1 << (s0 + 5)) + (q << 5) + (tt << (9 - S)), c);
    if constexpr (S >= 5)
        stage16<S - 5, true, GVAR, 1, false>(v, tw + (1 << (s0 + 4)) + (q << 4) + (tt << (8 - S)), c);
}

} // namespace heon
