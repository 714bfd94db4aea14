// Method-II mod-up (HPS fast base conversion with the reference's fp32 correction) fused into the
// column pass of the forward NTT.
//
// reference: base_conversion_DtoQtilde_relin_leveled_kernel (src/lib/kernel/switchkey.cu:985-1046) and
// base_conversion_DtoQtilde_relin_kernel (:872-927), followed by GPU_NTT_Modulus_Ordered_Inplace.
//
// The reference writes the converted digits tmp[b][d][Q'_l][N] to HBM and the transform reads them
// back.  Here a small preparation kernel leaves, per coefficient,
//     partial_m = x_m * Mi_inv_m mod q_m          (L*N words, the input scaled once)
//     r         = round(sum_j float(partial_j) / float(q_j))     (one byte per coefficient and digit,
//                 the fp32 sequence of the reference operation for operation)
// and the column pass of output polynomial (b, digit i, target y) computes its input words
//     sum_j partial_j * M_{j,y}  -  r * prod_y     (mod t_y)
// while it loads them: the converted digits never exist in memory.  For digits and targets below 2^50
// the products run on the FP64 pipe (fp_mulmod) and the value enters the butterflies directly as an
// integer-valued double (no canonicalisation, no integer round trip); other combinations use Shoup
// products.  Every variant is exact, so the transform's canonical output equals the reference's.
#include "ntt_impl.cuh"

namespace heon {

struct MapModUpII {
    const u64* part; // [b][L][N]: partial words; bit patterns of doubles for all-FP64 digits
    const unsigned char* rq; // [b][d][N]
    u64* out; // tmp [b][d][Qpl][N]
    const TwPair* bc_pair; // [digit][k over Q'_l][j in digit] (LevelTablesII::d_base_change_pair)
    const u64* rprod; // [r][digit][k]: r * prod mod t_k
    int d, Qpl, L, depth, logn, per_b, skip_own;
    unsigned long long dfp_mask; // bit i: every prime of digit i runs on the FP64 pipe
    short prefix[66], I_loc[65], I_j[65];

    static constexpr bool kXform = false;
    static constexpr bool kGather = true;
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&, int) const { return x; }

    __device__ __forceinline__ void locate(long long z, long long& b, int& i, int& y) const
    {
        b = z / per_b;
        const int zl = (int) (z % per_b);
        i = 0;
        while (i + 1 < d && zl >= prefix[i + 1])
            ++i;
        y = zl - prefix[i];
        if (skip_own && y >= I_loc[i])
            y += I_j[i];
    }
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& o, int& prime, int& aux) const
    {
        long long b;
        int i, y;
        locate(z, b, i, y);
        in = part + ((b * L + I_loc[i]) << logn);
        o = out + (((b * d + i) * Qpl + y) << logn);
        prime = level_prime(y, L, depth);
        aux = i;
    }

    struct Gather {
        const u64* src;
        const unsigned char* r;
        u64 mw[4], ms[4]; // conversion factors {M, companion}: doubles {M, RN(M/t)} or Shoup pairs
        u64 rp[5]; // r * prod mod t for r = 0..I_j (doubles' bit patterns on the FP64 path)
        u64 p;
        int ij;
        bool fp; // FP64 products (digit and target below 2^50)
        bool fp_src; // the digit's partial words are stored as doubles
    };
    __device__ __forceinline__ void gather_init(Gather& g, long long z, int prime, const PrimeConst& pc) const
    {
        long long b;
        int i, y;
        locate(z, b, i, y);
        g.src = part + ((b * L + I_loc[i]) << logn);
        g.r = rq + ((b * d + i) << logn);
        g.ij = I_j[i];
        g.p = pc.p;
        g.fp_src = (dfp_mask >> i) & 1;
        g.fp = g.fp_src && pc.fp_var != 0;
        const TwPair* m = bc_pair + (long long) I_loc[i] * Qpl + (long long) y * g.ij;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const TwPair t = j < g.ij ? ld_tw(m + j) : TwPair{0, 0};
            g.mw[j] = t.w;
            g.ms[j] = t.ws;
        }
#pragma unroll
        for (int r = 0; r < 5; ++r)
        {
            const u64 v = r <= g.ij ? __ldg(rprod + ((long long) r * d + i) * Qpl + y) : 0;
            g.rp[r] = g.fp ? d2u(fp_from_u64(v)) : v;
        }
    }
    // working representation (ct_prep form) of the thread's 16 input words idx = base + k*stride.
    // One source limb at a time: 16 loads in flight, then 16 products -- the load/compute shape of the
    // plain column pass, repeated per source limb.
    template <int VAR>
    __device__ __forceinline__ void gather16(const Gather& g, int base, int stride, u64 (&v)[16], const BflyConst& c) const
    {
        const long long lstep = 1ll << logn;
        unsigned rb[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            rb[k] = g.r[base + k * stride];
        if (VAR >= 3 && g.fp)
        {
            double acc[16]; // |acc| <= I_j * 0.57 t
#pragma unroll
            for (int k = 0; k < 16; ++k)
                acc[k] = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < g.ij)
                {
                    u64 raw[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        raw[k] = __ldg(g.src + j * lstep + base + k * stride);
                    const double w = u2d(g.mw[j]), wi = u2d(g.ms[j]);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        acc[k] = __dadd_rn(acc[k], fp_mulmod(u2d(raw[k]), w, wi, c.dnp));
                }
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                const unsigned r = rb[k];
                const u64 rp = r == 0 ? g.rp[0] : r == 1 ? g.rp[1] : r == 2 ? g.rp[2] : r == 3 ? g.rp[3] : g.rp[4];
                // |acc - rp| < 3.3 t; it enters the butterflies as a balanced residue |v| <= t/2 (+1): tighter
                // than a canonical word, so the operand bounds proven for canonical inputs
                // (tests/host_emul.cpp) hold for VAR 3 and VAR 4
                v[k] = d2u(fp_reduce(__dsub_rn(acc[k], u2d(rp)), c.dpinv, c.dnp));
            }
            return;
        }
        // integer form: Shoup products, canonical result
        const u64 p4 = 4 * g.p, np = 0 - g.p;
        u64 a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            a[k] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < g.ij)
            {
                u64 raw[16];
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    raw[k] = __ldg(g.src + j * lstep + base + k * stride);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                {
                    const u64 pj = g.fp_src ? (u64) __double2ll_rn(u2d(raw[k])) : raw[k];
                    a[k] = csub(a[k] + shoup_lazy_ptx(pj, g.mw[j], g.ms[j], np), p4);
                }
            }
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            const unsigned r = rb[k];
            const u64 rp = r == 0 ? g.rp[0] : r == 1 ? g.rp[1] : r == 2 ? g.rp[2] : r == 3 ? g.rp[3] : g.rp[4];
            u64 x = csub(csub(a[k], 2 * g.p), g.p);
            x = mod_sub(x, rp, g.p);
            v[k] = ct_prep<VAR>(x, c, false);
        }
    }
};


// Preparation: partial words and the fp32 correction, two coefficients per thread.
// reference: the first loop of base_conversion_DtoQtilde_relin_leveled_kernel (switchkey.cu:1002-1020);
// the float sequence (u64 -> f32 round-to-nearest, IEEE divide, sequential adds, round half away from
// zero) is reproduced operation for operation.
__global__ void __launch_bounds__(256)
    k_modup2_prep(const u64* __restrict__ coef, long long coef_bs, u64* __restrict__ part, unsigned char* __restrict__ rq,
                  const PrimeConst* __restrict__ pcs, const TwPair* __restrict__ mi_inv, const int* __restrict__ I_j_,
                  const int* __restrict__ I_loc_, int logn, int d, int L, unsigned long long dfp_mask)
{
    const int idx = (blockIdx.x * 256 + threadIdx.x) * 2;
    const int dg = blockIdx.y;
    const long long bz = blockIdx.z;
    const int I_j = I_j_[dg], I_loc = I_loc_[dg];
    const bool as_double = (dfp_mask >> dg) & 1;
    const u64* pin = coef + bz * coef_bs + ((long long) I_loc << logn) + idx;
    u64* po = part + ((bz * L + I_loc) << logn) + idx;
    float r0 = 0.f, r1 = 0.f;
    for (int i = 0; i < I_j; ++i)
    {
        const PrimeConst pi = pcs[I_loc + i];
        const TwPair mi = mi_inv[I_loc + i];
        const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(pin + ((long long) i << logn));
        const u64 p0 = csub(shoup_mul_lazy(x.x, mi.w, mi.ws, pi.p), pi.p);
        const u64 p1 = csub(shoup_mul_lazy(x.y, mi.w, mi.ws, pi.p), pi.p);
        const float mod = __ull2float_rn(pi.p);
        r0 = __fadd_rn(r0, __fdiv_rn(__ull2float_rn(p0), mod));
        r1 = __fadd_rn(r1, __fdiv_rn(__ull2float_rn(p1), mod));
        ulonglong2 o;
        o.x = as_double ? d2u(fp_from_u64(p0)) : p0;
        o.y = as_double ? d2u(fp_from_u64(p1)) : p1;
        *reinterpret_cast<ulonglong2*>(po + ((long long) i << logn)) = o;
    }
    uchar2 rr;
    rr.x = (unsigned char) (unsigned) roundf(r0);
    rr.y = (unsigned char) (unsigned) roundf(r1);
    *reinterpret_cast<uchar2*>(rq + ((bz * d + dg) << logn) + idx) = rr;
}

bool modup2_fused_available(const Context& c, int depth, const u64* coef, long long coef_bs)
{
    if (c.method != 2 || !c.modup_fused || c.n < 512 || c.P_size > 4)
        return false;
    const LevelTablesII& t = c.lvl2[depth];
    if (t.d > 64)
        return false;
    for (int i = 0; i < t.d; ++i)
        if (t.I_j[i] > 4)
            return false;
    return (coef_bs & 1) == 0 && (reinterpret_cast<uintptr_t>(coef) & 15) == 0;
}

// tmp[b][d][Q'_l][N] = NTT(mod-up(coef)) (all stages, or the column stages only when col_only);
// part: batch*L*N words, rq: batch*d*N bytes of scratch.  own_stashed: the digits' own limbs are
// neither converted nor transformed (they already hold the input's NTT words).
void launch_modup2_ntt(const Context& c, const u64* coef, long long coef_bs, u64* tmp, u64* part, unsigned char* rq,
                       int depth, long long batch, bool own_stashed, bool col_only, cudaStream_t st)
{
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    const LevelTablesII& t = c.lvl2[depth];
    MapModUpII m;
    m.part = part;
    m.rq = rq;
    m.out = tmp;
    m.bc_pair = t.d_base_change_pair;
    m.rprod = t.d_rprod;
    m.d = t.d;
    m.Qpl = Qpl;
    m.L = L;
    m.depth = depth;
    m.logn = c.logn;
    m.skip_own = own_stashed ? 1 : 0;
    m.dfp_mask = 0;
    int acc = 0;
    for (int i = 0; i < t.d; ++i)
    {
        m.prefix[i] = (short) acc;
        m.I_loc[i] = (short) t.I_loc[i];
        m.I_j[i] = (short) t.I_j[i];
        acc += own_stashed ? Qpl - t.I_j[i] : Qpl;
        bool dfp = c.use_fp64 && t.I_j[i] <= 4; // same rule as the table upload (context.cu)
        for (int j = 0; j < t.I_j[i]; ++j)
            dfp = dfp && c.mod[t.I_loc[i] + j].bit <= 50;
        if (dfp)
            m.dfp_mask |= 1ull << i;
    }
    m.prefix[t.d] = (short) acc;
    m.per_b = acc;
    {
        LaunchScope scope(KC_MODUP2, st);
        k_modup2_prep<<<dim3(c.n >> 9, t.d, (unsigned) batch), 256, 0, st>>>(coef, coef_bs, part, rq, c.d_pc, t.d_mi_inv_pair,
                                                                     t.d_I_j, t.d_I_loc, c.logn, t.d, L, m.dfp_mask);
    }
    const long long w = (batch * t.d * Qpl) << c.logn;
    run_ntt(c, m, batch * acc, false, Extent{tmp, w, tmp, w}, st, col_only);
}

} // namespace heon
