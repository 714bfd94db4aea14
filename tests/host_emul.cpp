// Host emulation of the device butterfly code (ntt_core.cuh compiled for the
// CPU with tiny shims): every lazy-reduction variant of the forward and
// inverse transform must reproduce the textbook merged NTT / INTT exactly.
// Built and run by tests/test_host_emulation.py (no GPU needed).
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b)
{
    return (unsigned long long) (((unsigned __int128) a * b) >> 64);
}
static inline ulonglong2 __ldg(const ulonglong2* p) { return *p; }
#include <cmath>
#include <cstring>
static inline double __hiloint2double(int hi, int lo)
{
    unsigned long long b = ((unsigned long long) (unsigned) hi << 32) | (unsigned) lo;
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}
static inline double __longlong_as_double(long long v)
{
    double d;
    std::memcpy(&d, &v, 8);
    return d;
}
static inline long long __double_as_longlong(double d)
{
    long long v;
    std::memcpy(&v, &d, 8);
    return v;
}
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static double g_fp_max = 0.0; // largest |Y| ever multiplied on the FP64 path
static inline void heon_fp_track(double y) { if (std::fabs(y) > g_fp_max) g_fp_max = std::fabs(y); }
#define HEON_FP_TRACK
#include "../heongpu_b200/csrc/ntt_core.cuh"
using namespace heon;

template <int VAR> static void fwd(std::vector<u64>& x, const std::vector<TwPair>& tw, const PrimeConst& pc)
{
    const BflyConst bc = make_bc(pc);
    for (int col = 0; col < 256; ++col) // column pass, S = 4
    {
        u64 v[16];
        for (int k = 0; k < 16; ++k) v[k] = ct_prep<VAR>(x[k * 256 + col], bc, true);
        ct_round_a<VAR>(v, tw.data(), 0, 0, bc);
        for (int k = 0; k < 16; ++k) x[k * 256 + col] = v[k];
    }
    for (int r = 0; r < 16; ++r) // row pass
    {
        u64 row[256];
        for (int tt = 0; tt < 16; ++tt)
        {
            u64 v[16];
            for (int k = 0; k < 16; ++k) v[k] = x[r * 256 + tt + 16 * k];
            ct_round_a<VAR, 1>(v, tw.data(), 4, r, bc);
            for (int k = 0; k < 16; ++k) row[tt + 16 * k] = v[k];
        }
        for (int tt = 0; tt < 16; ++tt)
        {
            u64 v[16];
            for (int k = 0; k < 16; ++k) v[k] = row[16 * tt + k];
            ct_round_b<8, VAR, 1>(v, tw.data(), 4, r, tt, bc);
            for (int k = 0; k < 16; ++k) x[r * 256 + 16 * tt + k] = ct_finish<VAR>(v[k], bc, pc);
        }
    }
}

// forward transform with the row pass of the TMA kernels: bare-double twiddles per row
// (stage16_sm: winv rebuilt from w with the two-word 1/p)
template <int VAR> static void fwd_sm(std::vector<u64>& x, const std::vector<TwPair>& twf, const std::vector<TwPair>& tw_int,
                                      const PrimeConst& pc)
{
    const BflyConst bc = make_bc(pc);
    const int S1 = 4;
    for (int col = 0; col < 256; ++col)
    {
        u64 v[16];
        for (int k = 0; k < 16; ++k) v[k] = ct_prep<VAR>(x[k * 256 + col], bc, true);
        ct_round_a<VAR>(v, twf.data(), 0, 0, bc);
        for (int k = 0; k < 16; ++k) x[k * 256 + col] = v[k];
    }
    for (int r = 0; r < 16; ++r)
    {
        double rowtw[256] = {0};
        for (int u = 0; u < 4; ++u)
            for (int g = 0; g < (1 << u); ++g)
                rowtw[(1 << u) - 1 + g] = (double) tw_int[(1u << (S1 + u)) + ((size_t) r << u) + g].w;
        for (int u = 4; u < 8; ++u)
            for (int g = 0; g < (1 << (u - 4)); ++g)
                for (int tt = 0; tt < 16; ++tt)
                    rowtw[16 + ((1 << (u - 4)) - 1 + g) * 16 + tt] =
                        (double) tw_int[(1u << (S1 + u)) + ((size_t) r << u) + (tt << (u - 4)) + g].w;
        u64 row[256];
        for (int tt = 0; tt < 16; ++tt)
        {
            u64 v[16];
            for (int k = 0; k < 16; ++k) v[k] = x[r * 256 + tt + 16 * k];
            ct_round_a_sm<VAR>(v, rowtw, bc);
            for (int k = 0; k < 16; ++k) row[tt + 16 * k] = v[k];
        }
        for (int tt = 0; tt < 16; ++tt)
        {
            u64 v[16];
            for (int k = 0; k < 16; ++k) v[k] = row[16 * tt + k];
            ct_round_b_sm<VAR>(v, rowtw, tt, bc);
            for (int k = 0; k < 16; ++k) x[r * 256 + 16 * tt + k] = ct_finish<VAR>(v[k], bc, pc);
        }
    }
}

template <int GVAR>
static void inv(std::vector<u64>& x, const std::vector<TwPair>& tw, const PrimeConst& pc, const TwPair& ninv,
                const TwPair& wninv)
{
    const BflyConst bc = make_bc(pc);
    for (int r = 0; r < 16; ++r) // row pass first
    {
        u64 row[256];
        for (int tt = 0; tt < 16; ++tt)
        {
            u64 v[16];
            for (int k = 0; k < 16; ++k) v[k] = gs_prep<GVAR>(x[r * 256 + 16 * tt + k]);
            gs_round_b<8, GVAR>(v, tw.data(), 4, r, tt, bc);
            for (int k = 0; k < 16; ++k) row[16 * tt + k] = v[k];
        }
        for (int tt = 0; tt < 16; ++tt)
        {
            u64 v[16];
            for (int k = 0; k < 16; ++k) v[k] = row[tt + 16 * k];
            gs_round_a<GVAR>(v, tw.data(), 4, r, bc);
            for (int k = 0; k < 16; ++k) x[r * 256 + tt + 16 * k] = v[k];
        }
    }
    for (int col = 0; col < 256; ++col)
    {
        u64 v[16];
        for (int k = 0; k < 16; ++k) v[k] = x[k * 256 + col];
        gs_round_a_final<GVAR>(v, tw.data(), bc, ninv, wninv);
        for (int k = 0; k < 16; ++k) x[k * 256 + col] = v[k];
    }
}

int main()
{
    const int logn = 12, N = 1 << logn;
    int failures = 0;
    const int bitsizes[] = {30, 40, 45, 46, 47, 49, 50, 57, 58, 60, 61};
    for (int bits : bitsizes)
    {
        const u64 p = bits == 61 ? 2305843009213554689ull : largest_ntt_primes(2 * N, bits, 1)[0];
        const u64 psi = minimal_primitive_root(2 * N, p), ipsi = invmod(psi, p);
        std::vector<TwPair> tw(N), itw(N), twf(N), itwf(N);
        std::vector<u64> pw(N), ipw(N);
        pw[0] = ipw[0] = 1;
        for (int j = 1; j < N; ++j)
        {
            pw[j] = mulmod(pw[j - 1], psi, p);
            ipw[j] = mulmod(ipw[j - 1], ipsi, p);
        }
        for (int j = 0; j < N; ++j)
        {
            tw[j] = TwPair{pw[bitrev(j, logn)], shoup(pw[bitrev(j, logn)], p)};
            itw[j] = TwPair{ipw[bitrev(j, logn)], shoup(ipw[bitrev(j, logn)], p)};
            const double wd = (double) tw[j].w, winv = wd / (double) p; // FP64 table format {w, RN(w/p)}
            std::memcpy(&twf[j].w, &wd, 8);
            std::memcpy(&twf[j].ws, &winv, 8);
            const double iwd = (double) itw[j].w, iwinv = iwd / (double) p;
            std::memcpy(&itwf[j].w, &iwd, 8);
            std::memcpy(&itwf[j].ws, &iwinv, 8);
        }
        PrimeConst pc;
        pc.p = p;
        pc.inv64 = shoup(1, p);
        pc.bits = bit_length(p);
        pc.fin_shift = pc.bits - 25;
        pc.fin_m = (unsigned) ((((u128) 1) << (pc.bits + 31)) / p);
        pc.nc_ok = pc.bits <= 57;
        pc.pinv = 1.0 / (double) p;
        pc.pinv_lo = std::fma(-(double) p, pc.pinv, 1.0) / (double) p;
        const u64 ni = invmod(N, p), wn = mulmod(itw[1].w, ni, p);
        const TwPair ninv{ni, shoup(ni, p)}, wninv{wn, shoup(wn, p)};
        for (int pattern = 0; pattern < 3; ++pattern)
        {
            std::vector<u64> a(N), ref;
            u64 s = 12345 + bits;
            for (int i = 0; i < N; ++i)
            {
                s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                a[i] = pattern == 0 ? (s >> 2) % p : pattern == 1 ? p - 1 : 0;
            }
            ref = a;
            int t = N, m = 1;
            while (m < N)
            {
                t >>= 1;
                for (int i = 0; i < m; ++i)
                    for (int j = 2 * i * t; j < 2 * i * t + t; ++j)
                    {
                        u64 U = ref[j], V = mulmod(ref[j + t], tw[m + i].w, p);
                        ref[j] = addmod(U, V, p);
                        ref[j + t] = submod(U, V, p);
                    }
                m <<= 1;
            }
            for (int var = 0; var < 5; ++var)
            {
                if (var == 2 && !pc.nc_ok)
                    continue;
                if ((var == 3 && pc.bits > 47) || (var == 4 && pc.bits > 50))
                    continue;
                std::vector<u64> x = a;
                // worst-case lazy input for the fused mod-up: words below 4p are legal inputs
                if (var != 0 && pattern == 0)
                    for (int i = 0; i < N; i += 3)
                        x[i] += 3 * p;
                var == 0   ? fwd<0>(x, tw, pc)
                : var == 1 ? fwd<1>(x, tw, pc)
                : var == 2 ? fwd<2>(x, tw, pc)
                : var == 3 ? fwd<3>(x, twf, pc)
                           : fwd<4>(x, twf, pc);
                int bad = 0;
                for (int i = 0; i < N; ++i) bad += x[i] != ref[i];
                if (bad)
                {
                    printf("FAIL fwd bits=%d var=%d pattern=%d mismatches=%d\n", bits, var, pattern, bad);
                    ++failures;
                }
            }
            for (int var = 3; var < 5; ++var) // shared-memory twiddle form of the row pass
            {
                if ((var == 3 && pc.bits > 47) || (var == 4 && pc.bits > 50))
                    continue;
                std::vector<u64> x = a;
                if (pattern == 0)
                    for (int i = 0; i < N; i += 3)
                        x[i] += 3 * p;
                var == 3 ? fwd_sm<3>(x, twf, tw, pc) : fwd_sm<4>(x, twf, tw, pc);
                int bad = 0;
                for (int i = 0; i < N; ++i) bad += x[i] != ref[i];
                if (bad)
                {
                    printf("FAIL fwd_sm bits=%d var=%d pattern=%d mismatches=%d\n", bits, var, pattern, bad);
                    ++failures;
                }
            }
            for (int gvar = 0; gvar < 5; ++gvar)
            {
                if (gvar == 2 || (gvar == 3 && pc.bits > 47) || (gvar == 4 && pc.bits > 50))
                    continue;
                std::vector<u64> x = ref;
                auto dpair = [&](u64 w) {
                    const double wd = (double) w, wi = wd / (double) p;
                    TwPair t;
                    std::memcpy(&t.w, &wd, 8);
                    std::memcpy(&t.ws, &wi, 8);
                    return t;
                };
                gvar == 0   ? inv<0>(x, itw, pc, ninv, wninv)
                : gvar == 1 ? inv<1>(x, itw, pc, ninv, wninv)
                : gvar == 3 ? inv<3>(x, itwf, pc, dpair(ni), dpair(wn))
                            : inv<4>(x, itwf, pc, dpair(ni), dpair(wn));
                int bad = 0;
                for (int i = 0; i < N; ++i) bad += x[i] != a[i];
                if (bad)
                {
                    printf("FAIL inv bits=%d gvar=%d pattern=%d mismatches=%d\n", bits, gvar, pattern, bad);
                    ++failures;
                }
            }
        }
    }
    const double ntt_fp_max = g_fp_max; // operands seen inside the transforms (incl. worst-case lazy inputs)
    // FP64 modular product: exact for ANY integer |Y| < 2^52 (random and edge operands)
    for (int bits : {30, 40, 46, 47, 49, 50})
    {
        const u64 p = largest_ntt_primes(2 * N, bits, 1)[0];
        const double dp = (double) p;
        u64 s = 777 + bits;
        int bad = 0;
        for (int it = 0; it < 400000; ++it)
        {
            s = s * 6364136223846793005ULL + 1442695040888963407ULL;
            u64 w = (it % 7 == 0) ? p - 1 : (it % 11 == 0) ? 1 : (s >> 4) % p;
            s = s * 6364136223846793005ULL + 1442695040888963407ULL;
            // operand range: 2^52 with the FRND rounding, 2^51 with the magic constant (its sum has to stay
            // inside one binade: |q| < 2^51)
            const int top = HEON_FP_FRND ? 52 : 51;
            u64 y = s >> (64 - top);
            if (it % 5 == 0) y = (y / p) * p + (it % 3) - 1 + (y < p ? p : 0); // multiples of p, +-1
            if (it % 13 == 0) y = (1ull << top) - 1 - (it & 7);
            if (it % 17 == 0) y = (1ull << 51) - 9 + (it & 7);
            if (y >= (1ull << top)) y = (1ull << top) - 1;
            const bool neg = (it & 1);
            const double yd = neg ? -(double) y : (double) y;
            const double wd = (double) w, winv = wd / dp;
            const double t = fp_mulmod(yd, wd, winv, -dp);
            // |T| <= p*(1/2 + ulp(Y*winv)/2 + |Y|*2^-54) and T == Y*w (mod p); ulp(Y*winv)/2 <= 1/4 below 2^52
            // (1/8 below 2^51; 0 for the magic-constant rounding)
            const double lim = dp * (0.5 + (HEON_FP_FRND ? (y < (1ull << 51) ? 0.125 : 0.25) : 0.0) + (double) y * 0x1p-54) + 1.0;
            u64 want = mulmod(y % p, w, p);
            if (neg && want) want = p - want;
            long long ti = (long long) t;
            u64 got = (u64) ((ti % (long long) p + (long long) p) % (long long) p);
            if (!(t <= lim && t >= -lim) || (double) ti != t || got != want)
                ++bad;
        }
        if (bad)
        {
            printf("FAIL fp_mulmod bits=%d bad=%d\n", bits, bad);
            ++failures;
        }
    }
    // the FP64 butterflies are exact only while every multiplied operand stays below 2^51 (2^52 with FRND)
    printf("max |Y| inside the FP64 transforms: 2^%.2f\n", std::log2(ntt_fp_max));
    if (!(ntt_fp_max < (HEON_FP_FRND ? 0x1p52 : 0x1p51)))
    {
        printf("FAIL: FP64 operand bound exceeded\n");
        ++failures;
    }
    printf(failures ? "FAILED %d\n" : "OK\n", failures);
    return failures != 0;
}
