// Batched negacyclic NTT / INTT over RNS limbs for sm_100a.
//
// Replaces gpuntt::GPU_NTT / GPU_INTT / *_Modulus_Ordered / *_Poly_Ordered
// (reference: thirdparty/GPU-NTT/src/lib/ntt_merge/ntt.cu:596-763,1204-1320,
// 3106-3255,3405-3502,3785-3935,4085-4182; host dispatch 2563-3103,3603-3783,
// 4284-4466).  Same transform: psi-merged Cooley-Tukey forward (natural in,
// bit-reversed out), Gentleman-Sande inverse with the final N^-1, twiddles
// psi^bitrev(i) (reference table layout util.cu:398-451).
//
// Structure: N = 2^n is viewed as a (2^(n-8) x 256) matrix.  The forward
// transform is a column pass (first n-8 stages, stride >= 256) followed by a
// row pass (last 8 stages inside 2 KiB rows); the inverse runs the row pass
// first.  Each thread keeps 16 coefficients in registers and performs four
// radix-2 stages per round; rounds are separated by one shared-memory
// transpose.  Butterflies are Harvey/Shoup lazy butterflies (values kept in
// [0,4p) forward, [0,2p) inverse); every word is canonicalised before the
// final store, so results equal the reference's Barrett arithmetic bit for bit.
#include "modarith.cuh"
#include "ntt_core.cuh"
#include "ops.hpp"

namespace heon {

// ---------------------------------------------------------------------------
// poly -> (input pointer, output pointer, prime) maps
// ---------------------------------------------------------------------------

// contiguous polys; prime = list[z % count].  `src` may differ from `dst`
// (out-of-place transform), both are [n_polys][N].
struct MapContig {
    const u64* src;
    u64* dst;
    PrimeList pl;
    int logn;
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& prime) const
    {
        in = src + (z << logn);
        out = dst + (z << logn);
        prime = pl.idx[z % pl.count];
    }
    static constexpr bool kXform = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&) const { return x; }
};

// polys at explicit word offsets, one prime, in place
struct MapScatter {
    u64* base;
    const long long* offs; // device array
    int prime;
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& pr) const
    {
        in = out = base + offs[z];
        pr = prime;
    }
    static constexpr bool kXform = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&) const { return x; }
};

// Strided polys: poly z = (b, j) with j < per_batch lives at
// base + b*bstride + (first + j)*N; prime = list[j % count].  In place.
struct MapStrided {
    u64* base;
    long long bstride;
    int per_batch, first;
    PrimeList pl;
    int logn;
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& prime) const
    {
        long long b = z / per_batch;
        int j = (int) (z % per_batch);
        in = out = base + b * bstride + ((long long) (first + j) << logn);
        prime = pl.idx[j % pl.count];
    }
    static constexpr bool kXform = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&) const { return x; }
};

// Out-of-place strided source -> contiguous destination (used by apply_galois
// to leave the input ciphertext untouched).
struct MapStridedCopy {
    const u64* src;
    u64* dst;
    long long src_bstride;
    int per_batch;
    PrimeList pl;
    int logn;
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& prime) const
    {
        long long b = z / per_batch;
        int j = (int) (z % per_batch);
        in = src + b * src_bstride + ((long long) j << logn);
        out = dst + (z << logn);
        prime = pl.idx[j % pl.count];
    }
    static constexpr bool kXform = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&) const { return x; }
};

// Method-I mod-up fused into the first pass: output poly z = (b, i, y) reads
// digit i of ciphertext b (coefficient domain) and reduces it into prime y.
// Replaces cipher_broadcast_leveled_kernel / ckks_duplicate_kernel
// (reference: src/lib/kernel/switchkey.cu:29-59, 1558-1590).
struct MapModUpI {
    const u64* coef; // digits (coefficient domain): coef + b*bstride + i*N
    u64* out; // [b][L][Qpl][N]
    long long coef_bstride;
    int L, Qpl, depth, logn;
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& o, int& prime) const
    {
        int y = (int) (z % Qpl);
        long long t = z / Qpl;
        int i = (int) (t % L);
        long long b = t / L;
        in = coef + b * coef_bstride + ((long long) i << logn);
        o = out + (z << logn);
        prime = level_prime(y, L, depth);
    }
    static constexpr bool kXform = true;
    // x mod p, lazily in [0,2p): good enough as NTT input
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst& pc) const
    {
        return shoup_mul_lazy(x, 1, pc.inv64, pc.p);
    }
};

// Divide-and-round stage one fused into the first pass: output poly
// z = (b, c, i) reads the dropped limb of component c (coefficient domain),
// adds half, reduces into q_i and subtracts half mod q_i.
// Replaces divide_round_lastq_leveled_stage_one_kernel
// (reference: src/lib/kernel/switchkey.cu:678-705).
struct MapDivRoundOne {
    const u64* src; // dropped limb of comp c: src + b*bstride + c*cstride
    u64* out; // [b][2][Lout][N]
    long long bstride, cstride;
    int Lout, logn;
    u64 half, plast; // floor(p_last/2), p_last
    const u64* half_mod; // [Lout]
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& o, int& prime) const
    {
        int i = (int) (z % Lout);
        long long t = z / Lout;
        int c = (int) (t & 1);
        long long b = t >> 1;
        in = src + b * bstride + c * cstride;
        o = out + (z << logn);
        prime = i;
    }
    static constexpr bool kXform = true;
    __device__ __forceinline__ u64 xform(u64 x, int prime, const PrimeConst& pc) const
    {
        x = mod_add(x, half, plast);
        x = reduce_u64(x, pc);
        return mod_sub(x, half_mod[prime], pc.p);
    }
};

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

// Column pass: S stages on columns (stride 256 words).  T = 2^S/16 threads
// cooperate on one column, C = 256/T adjacent columns per CTA.
template <int S, bool INV, class Map>
__global__ void __launch_bounds__(256) ntt_col_pass(Map map, const TwPair* __restrict__ tw_all,
                                                    const PrimeConst* __restrict__ pcs,
                                                    const TwPair* __restrict__ inv_last, int logn,
                                                    bool first_pass)
{
    constexpr int T = (1 << S) / 16;
    constexpr int C = 256 / T;
    __shared__ u64 sm[(S > 4) ? (1 << S) * C : 1];

    const int tiles = T; // 256 / C
    long long z = blockIdx.x / tiles;
    int tile = blockIdx.x % tiles;
    const u64* in;
    u64* out;
    int prime;
    map.get(z, in, out, prime);
    if (!first_pass)
        in = out;
    const PrimeConst pc = pcs[prime];
    const u64 p = pc.p, p2 = 2 * pc.p;
    const TwPair* tw = tw_all + ((long long) prime << logn);

    const int c = threadIdx.x % C;
    const int tt = threadIdx.x / C;
    const int col = tile * C + c;
    u64 v[16];

    if constexpr (!INV)
    {
        // forward: first pass of the transform
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            u64 x = in[(long long) (tt + T * k) * 256 + col];
            if (Map::kXform && first_pass)
                x = map.xform(x, prime, pc);
            v[k] = x;
        }
        ct_round_a(v, tw, 0, 0, p, p2);
        if constexpr (S > 4)
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                sm[(tt + T * k) * C + c] = v[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = sm[(16 * tt + k) * C + c];
            ct_round_b<S>(v, tw, 0, 0, tt, p, p2);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                out[(long long) (16 * tt + k) * 256 + col] = v[k]; // lazy [0,4p)
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                out[(long long) (tt + T * k) * 256 + col] = v[k];
        }
    }
    else
    {
        // inverse: last pass of the transform, folds N^-1 into the last stage
        const TwPair ninv = inv_last[2 * prime], wninv = inv_last[2 * prime + 1];
        if constexpr (S > 4)
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = in[(long long) (16 * tt + k) * 256 + col];
            gs_round_b<S>(v, tw, 0, 0, tt, p, p2);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                sm[(16 * tt + k) * C + c] = v[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = sm[(tt + T * k) * C + c];
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = in[(long long) (tt + T * k) * 256 + col];
        }
        gs_round_a_final(v, tw, p, p2, ninv, wninv);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            out[(long long) (tt + T * k) * 256 + col] = v[k];
    }
}

// Row pass: the 8 stages that live inside one 256-word row.  16 threads per
// row, 16 rows per CTA.  S1 = n - 8 is the number of column-pass stages.
template <bool INV, class Map>
__global__ void __launch_bounds__(256) ntt_row_pass(Map map, const TwPair* __restrict__ tw_all,
                                                    const PrimeConst* __restrict__ pcs, int logn,
                                                    bool first_pass)
{
    constexpr int PITCH = 288; // 256 + 2 words of padding per 16
    __shared__ __align__(16) u64 sm[16 * PITCH];
    const int S1 = logn - 8;
    const int tiles = (1 << S1) / 16;
    long long z = blockIdx.x / tiles;
    int tile = blockIdx.x % tiles;
    const u64* in;
    u64* out;
    int prime;
    map.get(z, in, out, prime);
    if (!first_pass)
        in = out;
    const PrimeConst pc = pcs[prime];
    const u64 p = pc.p, p2 = 2 * pc.p;
    const TwPair* tw = tw_all + ((long long) prime << logn);

    const int tt = threadIdx.x & 15;
    const int rl = threadIdx.x >> 4;
    const int r = tile * 16 + rl;
    const u64* rin = in + (long long) r * 256;
    u64* rout = out + (long long) r * 256;
    u64* srow = sm + rl * PITCH;
    u64 v[16];

    if constexpr (!INV)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = rin[tt + 16 * k];
        ct_round_a(v, tw, S1, r, p, p2);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            srow[tt + 18 * k] = v[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(srow + 18 * tt + k);
            v[k] = t2.x;
            v[k + 1] = t2.y;
        }
        ct_round_b<8>(v, tw, S1, r, tt, p, p2);
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2;
            t2.x = csub(csub(v[k], p2), p);
            t2.y = csub(csub(v[k + 1], p2), p);
            *reinterpret_cast<ulonglong2*>(rout + 16 * tt + k) = t2;
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(rin + 16 * tt + k);
            v[k] = t2.x;
            v[k + 1] = t2.y;
        }
        gs_round_b<8>(v, tw, S1, r, tt, p, p2);
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2;
            t2.x = v[k];
            t2.y = v[k + 1];
            *reinterpret_cast<ulonglong2*>(srow + 18 * tt + k) = t2;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = srow[tt + 18 * k];
        gs_round_a(v, tw, S1, r, p, p2);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            rout[tt + 16 * k] = v[k]; // lazy [0,2p), finished by the column pass
    }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------

template <bool INV, class Map>
static void launch_col(const Context& c, const Map& m, long long n_polys, bool first, cudaStream_t st)
{
    const int S = c.logn - 8;
    const unsigned grid = (unsigned) (n_polys * ((1 << S) / 16));
    LaunchScope scope(INV ? KC_NTT_INV_COL : KC_NTT_FWD_COL, st);
#define HEON_COL(SS)                                                                               \
    case SS:                                                                                       \
        ntt_col_pass<SS, INV, Map><<<grid, 256, 0, st>>>(m, INV ? c.d_inv : c.d_fwd, c.d_pc,       \
                                                         c.d_inv_last, c.logn, first);             \
        break;
    switch (S)
    {
        HEON_COL(4)
        HEON_COL(5)
        HEON_COL(6)
        HEON_COL(7)
        HEON_COL(8)
        default:
            throw std::invalid_argument("unsupported ring size");
    }
#undef HEON_COL
}

template <bool INV, class Map>
static void launch_row(const Context& c, const Map& m, long long n_polys, bool first, cudaStream_t st)
{
    const int S = c.logn - 8;
    const unsigned grid = (unsigned) (n_polys * ((1 << S) / 16));
    LaunchScope scope(INV ? KC_NTT_INV_ROW : KC_NTT_FWD_ROW, st);
    ntt_row_pass<INV, Map><<<grid, 256, 0, st>>>(m, INV ? c.d_inv : c.d_fwd, c.d_pc, c.logn, first);
}

template <class Map>
static void run_ntt(const Context& c, const Map& m, long long n_polys, bool inverse, cudaStream_t st)
{
    if (n_polys <= 0)
        return;
    if (!inverse)
    {
        launch_col<false>(c, m, n_polys, true, st);
        launch_row<false>(c, m, n_polys, false, st);
    }
    else
    {
        launch_row<true>(c, m, n_polys, true, st);
        launch_col<true>(c, m, n_polys, false, st);
    }
}

void launch_ntt(const Context& c, const u64* src, u64* dst, long long n_polys, const PrimeList& pl,
                bool inverse, cudaStream_t st)
{
    MapContig m{src, dst, pl, c.logn};
    run_ntt(c, m, n_polys, inverse, st);
}

void launch_ntt_scattered(const Context& c, u64* base, const long long* d_offsets, int n_polys,
                          int prime, bool inverse, cudaStream_t st)
{
    MapScatter m{base, d_offsets, prime};
    run_ntt(c, m, n_polys, inverse, st);
}

void launch_ntt_strided(const Context& c, u64* base, long long bstride, int per_batch, int first,
                        long long batch, const PrimeList& pl, bool inverse, cudaStream_t st)
{
    MapStrided m{base, bstride, per_batch, first, pl, c.logn};
    run_ntt(c, m, batch * per_batch, inverse, st);
}

void launch_ntt_strided_copy(const Context& c, const u64* src, long long src_bstride, u64* dst,
                             int per_batch, long long batch, const PrimeList& pl, bool inverse,
                             cudaStream_t st)
{
    MapStridedCopy m{src, dst, src_bstride, per_batch, pl, c.logn};
    run_ntt(c, m, batch * per_batch, inverse, st);
}

void launch_modup1_ntt(const Context& c, const u64* coef, long long coef_bstride, u64* out, int L,
                       int depth, long long batch, cudaStream_t st)
{
    const int Qpl = L + c.P_size;
    MapModUpI m{coef, out, coef_bstride, L, Qpl, depth, c.logn};
    run_ntt(c, m, batch * L * Qpl, false, st);
}

void launch_divround1_ntt(const Context& c, const u64* src, long long bstride, long long cstride,
                          u64* out, int Lout, u64 half, u64 plast, const u64* d_half_mod,
                          long long batch, cudaStream_t st)
{
    MapDivRoundOne m{src, out, bstride, cstride, Lout, c.logn, half, plast, d_half_mod};
    run_ntt(c, m, batch * 2 * Lout, false, st);
}

} // namespace heon
