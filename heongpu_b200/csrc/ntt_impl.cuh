#pragma once
#include <type_traits>
// Template implementation of the batched NTT (maps, kernels, launch helpers).  Included by the ntt_*.cu
// translation units, each of which instantiates it for a few polynomial maps (parallel compilation).
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
#include "tma.cuh"
#include <algorithm>
#include <stdexcept>
#include <string>

#ifndef HEON_NTT_MINBLOCKS
#define HEON_NTT_MINBLOCKS 3
#endif
#ifndef HEON_COL_MINBLOCKS
#define HEON_COL_MINBLOCKS HEON_NTT_MINBLOCKS
#endif

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
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& prime, int& aux) const
    {
        in = src + (z << logn);
        out = dst + (z << logn);
        prime = pl.idx[z % pl.count];
        aux = 0;
    }
    static constexpr bool kXform = false;
    static constexpr bool kGather = false;
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&, int) const { return x; }
};

// polys at explicit word offsets, one prime, in place
struct MapScatter {
    u64* base;
    const long long* offs; // device array
    int prime;
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& pr, int& aux) const
    {
        in = out = base + offs[z];
        pr = prime;
        aux = 0;
    }
    static constexpr bool kXform = false;
    static constexpr bool kGather = false;
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&, int) const { return x; }
};

// Strided polys: poly z = (b, j) with j < per_batch lives at
// base + b*bstride + (first + j)*N; prime = list[j % count].  In place.
struct MapStrided {
    u64* base;
    long long bstride;
    int per_batch, first;
    PrimeList pl;
    int logn;
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& prime, int& aux) const
    {
        long long b = z / per_batch;
        int j = (int) (z % per_batch);
        in = out = base + b * bstride + ((long long) (first + j) << logn);
        prime = pl.idx[j % pl.count];
        aux = 0;
    }
    static constexpr bool kXform = false;
    static constexpr bool kGather = false;
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&, int) const { return x; }
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
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& prime, int& aux) const
    {
        long long b = z / per_batch;
        int j = (int) (z % per_batch);
        in = src + b * src_bstride + ((long long) j << logn);
        out = dst + (z << logn);
        prime = pl.idx[j % pl.count];
        aux = 0;
    }
    static constexpr bool kXform = false;
    static constexpr bool kGather = false;
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&, int) const { return x; }
};

// Method-II key-switch buffer tmp[b][digit][Q'_l][N] WITHOUT the digits' own limbs (those hold the
// original NTT-domain words already: NTT(INTT(x)) = x).  Poly z = (b, i, y') enumerates, per digit i,
// the Q'_l - I_j[i] limbs outside [I_loc[i], I_loc[i] + I_j[i]).  In place.
struct MapDigitSkip {
    u64* base;
    int d, Qpl, L, depth, logn, per_b;
    unsigned long long dbl_mask; // bit i: the FP64-prime words of digit i are integer-valued doubles
    short prefix[66], I_loc[65], I_j[65];
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& out, int& prime, int& aux) const
    {
        const long long b = z / per_b;
        const int zl = (int) (z % per_b);
        int i = 0;
        while (i + 1 < d && zl >= prefix[i + 1])
            ++i;
        int y = zl - prefix[i];
        if (y >= I_loc[i])
            y += I_j[i];
        in = out = base + (((b * d + i) * Qpl + y) << logn);
        prime = level_prime(y, L, depth);
        aux = (int) ((dbl_mask >> i) & 1ull);
    }
    static constexpr bool kXform = false;
    static constexpr bool kGather = false;
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst&, int) const { return x; }
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
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& o, int& prime, int& aux) const
    {
        int y = (int) (z % Qpl);
        long long t = z / Qpl;
        int i = (int) (t % L);
        long long b = t / L;
        in = coef + b * coef_bstride + ((long long) i << logn);
        o = out + (z << logn);
        prime = level_prime(y, L, depth);
        // The digit word x < 2^bits(q_i) is already a valid lazy NTT input (< 4p)
        // when the digit prime is at most one bit longer than the target prime.
        aux = pcs[i].bits > pcs[prime].bits + 1;
    }
    static constexpr bool kXform = true;
    static constexpr bool kGather = false;
    static constexpr bool kLazyIn = true; // words in [0,4p)
    const PrimeConst* pcs;
    __device__ __forceinline__ u64 xform(u64 x, int, const PrimeConst& pc, int need_reduce) const
    {
        return need_reduce ? shoup_mul_lazy3(x, 1, pc.inv64, pc.p) : x; // [0,4p) either way
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
    __device__ __forceinline__ void get(long long z, const u64*& in, u64*& o, int& prime, int& aux) const
    {
        int i = (int) (z % Lout);
        long long t = z / Lout;
        int c = (int) (t & 1);
        long long b = t >> 1;
        in = src + b * bstride + c * cstride;
        o = out + (z << logn);
        prime = i;
        aux = 0;
    }
    static constexpr bool kXform = true;
    static constexpr bool kGather = false;
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int prime, const PrimeConst& pc, int) const
    {
        x = mod_add(x, half, plast);
        x = reduce_u64(x, pc);
        return mod_sub(x, half_mod[prime], pc.p);
    }
};

// kDoubleAux<Map>: `aux` = 1 marks a polynomial whose FP64-prime words were left as integer-valued doubles
// (|v| <= p/2) by the producer (the fast Method-II mod-up): the column pass takes them as they are instead of
// converting canonical integers.
template <class Map> struct MapDoubleAux { static constexpr bool value = false; };
template <> struct MapDoubleAux<MapDigitSkip> { static constexpr bool value = true; };

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

// Column pass body: S stages on columns (stride 256 words).  T = 2^S/16 threads
// cooperate on one column, C = 256/T adjacent columns per CTA.
template <int S, bool INV, int VAR, class Map, int NT = 256>
__device__ __forceinline__ void col_pass_body(const Map& map, const u64* in, u64* out, int prime,
                                              const PrimeConst& pc, const TwPair* __restrict__ tw,
                                              const TwPair* __restrict__ inv_last, int tile,
                                              bool first_pass, int aux, u64* sm, long long z = 0)
{
    constexpr int T = (1 << S) / 16;
    constexpr int C = NT / T; // columns per CTA
    const BflyConst bc = make_bc(pc);
    const int c = threadIdx.x % C;
    const int tt = threadIdx.x / C;
    const int col = tile * C + c;
    u64 v[16];
    // transposition space: row r of the column tile at word smi(r).  With 8 columns a row is half a bank line and
    // the rows 16 apart that a half-warp reads back would share their banks: the row's lowest bit is flipped by
    // its bit 4, which keeps the writes (adjacent rows) and the reads (rows 16 apart) conflict-free
    auto smi = [&](int r) { return (C == 8 ? (r ^ ((r >> 4) & 1)) : r) * C + c; };

    if constexpr (!INV)
    {
        // forward: first pass of the transform
        if constexpr (Map::kGather)
        {
            // the map computes every input word from several source words (fused Method-II mod-up)
            typename Map::Gather g;
            map.gather_init(g, z, prime, pc);
            map.template gather16<VAR>(g, tt * 256 + col, T * 256, v, bc);
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                u64 x = in[(long long) (tt + T * k) * 256 + col];
                if (Map::kXform && first_pass)
                    x = map.xform(x, prime, pc, aux);
                if (MapDoubleAux<Map>::value && VAR >= 3 && aux)
                    v[k] = x;
                else
                    v[k] = ct_prep<VAR>(x, bc, Map::kLazyIn);
            }
        }
        ct_round_a<VAR>(v, tw, 0, 0, bc);
        if constexpr (S > 4)
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                sm[smi(tt + T * k)] = v[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = sm[smi(16 * tt + k)];
            ct_round_b<S, VAR>(v, tw, 0, 0, tt, bc);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                out[(long long) (16 * tt + k) * 256 + col] = v[k]; // lazy, finished by the row pass
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
            gs_round_b<S, VAR>(v, tw, 0, 0, tt, bc);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                sm[smi(16 * tt + k)] = v[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = sm[smi(tt + T * k)];
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = in[(long long) (tt + T * k) * 256 + col];
        }
        gs_round_a_final<VAR>(v, tw, bc, ninv, wninv);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            out[(long long) (tt + T * k) * 256 + col] = v[k];
    }
}

// NT threads per CTA: 256 (16 columns at N = 2^16, 32 KiB of transposition space) or 128 (8 columns).
// gather maps (fused mod-up) hold a second set of 16 source words next to the 16 accumulators: they get a
// larger register budget (2 CTAs of 256 / 5 CTAs of 128 threads per SM instead of 3 / 6)
template <int S, bool INV, class Map, int NT = 256>
__global__ void __launch_bounds__(NT, Map::kGather ? (NT == 256 ? 2 : 5) : HEON_COL_MINBLOCKS * 256 / NT) ntt_col_pass(Map map, const TwPair* __restrict__ tw_all,
                                                    const PrimeConst* __restrict__ pcs,
                                                    const TwPair* __restrict__ inv_last, int logn,
                                                    bool first_pass, int variant)
{
    constexpr int T = (1 << S) / 16;
    constexpr int C = NT / T;
    __shared__ u64 sm[(S > 4) ? (1 << S) * C : 1];
    const int tiles = 256 / C;
    long long z = blockIdx.x / tiles;
    int tile = blockIdx.x % tiles;
    const u64* in;
    u64* out;
    int prime, aux;
    map.get(z, in, out, prime, aux);
    if (!first_pass)
        in = out;
    const PrimeConst pc = pcs[prime];
    const TwPair* tw = tw_all + ((long long) prime << logn);
    if (pc.fp_var == 3)
        col_pass_body<S, INV, 3, Map, NT>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm, z);
    else if (pc.fp_var == 4)
        col_pass_body<S, INV, 4, Map, NT>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm, z);
    else if (INV || variant == 1 || !pc.nc_ok)
        col_pass_body<S, INV, 1, Map, NT>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm, z);
    else
        col_pass_body<S, INV, 2, Map, NT>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm, z);
}

// Row pass: the 8 stages that live inside one 256-word row.  16 threads per
// row, 16 rows per CTA.  S1 = n - 8 is the number of column-pass stages.
template <bool INV, int VAR>
__device__ __forceinline__ void row_pass_body(const u64* rin, u64* rout, const PrimeConst& pc,
                                              const TwPair* __restrict__ tw, int S1, int r, int tt,
                                              u64* srow)
{
    const BflyConst bc = make_bc(pc);
    u64 v[16];
    if constexpr (!INV)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = rin[tt + 16 * k];
        ct_round_a<VAR, 1>(v, tw, S1, r, bc);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            srow[tt + 18 * k] = v[k];
        __syncwarp(); // a row lives in one half-warp: the transpose is warp-local
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(srow + 18 * tt + k);
            v[k] = t2.x;
            v[k + 1] = t2.y;
        }
        ct_round_b<8, VAR, 1>(v, tw, S1, r, tt, bc);
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2;
            t2.x = ct_finish<VAR>(v[k], bc, pc);
            t2.y = ct_finish<VAR>(v[k + 1], bc, pc);
            *reinterpret_cast<ulonglong2*>(rout + 16 * tt + k) = t2;
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(rin + 16 * tt + k);
            v[k] = gs_prep<VAR>(t2.x);
            v[k + 1] = gs_prep<VAR>(t2.y);
        }
        gs_round_b<8, VAR>(v, tw, S1, r, tt, bc);
#pragma unroll
        for (int k = 0; k < 16; k += 2)
        {
            ulonglong2 t2;
            t2.x = v[k];
            t2.y = v[k + 1];
            *reinterpret_cast<ulonglong2*>(srow + 18 * tt + k) = t2;
        }
        __syncwarp(); // a row lives in one half-warp: the transpose is warp-local
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = srow[tt + 18 * k];
        gs_round_a<VAR>(v, tw, S1, r, bc);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            rout[tt + 16 * k] = v[k]; // lazy, finished by the column pass
    }
}

template <bool INV, class Map>
__global__ void __launch_bounds__(256, HEON_NTT_MINBLOCKS) ntt_row_pass(Map map, const TwPair* __restrict__ tw_all,
                                                    const PrimeConst* __restrict__ pcs, int logn,
                                                    bool first_pass, int variant)
{
    constexpr int PITCH = 288; // 256 + 2 words of padding per 16
    __shared__ __align__(16) u64 sm[16 * PITCH];
    const int S1 = logn - 8;
    const int tiles = (1 << S1) / 16;
    long long z = blockIdx.x / tiles;
    int tile = blockIdx.x % tiles;
    const u64* in;
    u64* out;
    int prime, aux;
    map.get(z, in, out, prime, aux);
    if (!first_pass)
        in = out;
    const PrimeConst pc = pcs[prime];
    const TwPair* tw = tw_all + ((long long) prime << logn);
    const int tt = threadIdx.x & 15;
    const int rl = threadIdx.x >> 4;
    const int r = tile * 16 + rl;
    const u64* rin = in + (long long) r * 256;
    u64* rout = out + (long long) r * 256;
    u64* srow = sm + rl * PITCH;
    if (pc.fp_var == 3)
        row_pass_body<INV, 3>(rin, rout, pc, tw, S1, r, tt, srow);
    else if (pc.fp_var == 4)
        row_pass_body<INV, 4>(rin, rout, pc, tw, S1, r, tt, srow);
    else if (INV || variant == 1 || !pc.nc_ok)
        row_pass_body<INV, 1>(rin, rout, pc, tw, S1, r, tt, srow);
    else
        row_pass_body<INV, 2>(rin, rout, pc, tw, S1, r, tt, srow);
}

// ---------------------------------------------------------------------------
// Row pass through TMA.  One CTA owns a tile of 16 rows (256 lines of 128 B,
// 32 KiB).  A 2-D tensor map over "lines of sixteen 64-bit words" brings the
// tile into shared memory with the 128-byte swizzle (UTMALDG), the threads
// run the eight stages out of registers with ONE in-place, warp-local,
// bank-conflict-free transpose, and the canonical result leaves through the
// same swizzled buffer with a TMA store (UTMASTG).  The load/store unit only
// sees shared-memory traffic and coalesced twiddle reads.
//
// Swizzled position of element e of line l of a row (row base 2 KiB aligned):
//   byte = l*128 + ((e>>1) ^ (l&7))*16 + (e&1)*8
// Round A (idx = tt + 16k) touches element tt of line k: a permutation inside
// one 128-byte line -> conflict free.  Round B (idx = 16tt + k) touches the
// eight 16-byte chunks of line tt at chunk positions c ^ (tt&7) -> the eight
// lanes of a quarter warp hit eight different bank groups.
// ---------------------------------------------------------------------------
constexpr int kRowTileBytes = 16 * 2048;

// Forward row stages on one swizzled row: loads the thread's 16 words in the round-A layout, runs the
// eight stages with the in-place warp-local transpose, and leaves the LAZY results in registers in the
// round-B layout (v[2c], v[2c+1] = 16-byte chunk c of line tt).
template <int VAR>
__device__ __forceinline__ void row_fwd_stages(unsigned char* rowp, const BflyConst& bc, const TwPair* __restrict__ tw,
                                               const TwPair* __restrict__ blk, int S1, int r, int tt,
                                               const double* rowtw, u64 (&v)[16])
{
    unsigned char* lineB = rowp + tt * 128;
    const int sw = tt & 7;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        v[k] = *reinterpret_cast<const u64*>(rowp + k * 128 + ((((tt >> 1) ^ (k & 7)) << 4) | ((tt & 1) << 3)));
    if ((VAR == 3 || VAR == 4) && rowtw)
        ct_round_a_sm<VAR>(v, rowtw, bc);
    else
        ct_round_a<VAR, 1>(v, tw, S1, r, bc);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        *reinterpret_cast<u64*>(rowp + k * 128 + ((((tt >> 1) ^ (k & 7)) << 4) | ((tt & 1) << 3))) = v[k];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c)
    {
        const ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(lineB + ((c ^ sw) << 4));
        v[2 * c] = t2.x;
        v[2 * c + 1] = t2.y;
    }
    if ((VAR == 3 || VAR == 4) && rowtw)
        ct_round_b_sm<VAR>(v, rowtw, tt, bc);
    else
        ct_round_b_lm<VAR, 1>(v, blk, tt, bc);
}

template <bool INV, int VAR>
__device__ __forceinline__ void row_pass_tma_body(unsigned char* rowp, const PrimeConst& pc,
                                                  const TwPair* __restrict__ tw,
                                                  const TwPair* __restrict__ blk, int S1, int r, int tt,
                                                  const double* rowtw)
{
    const BflyConst bc = make_bc(pc);
    u64 v[16];
    unsigned char* lineB = rowp + tt * 128;
    const int sw = tt & 7;
    if constexpr (!INV)
    {
        row_fwd_stages<VAR>(rowp, bc, tw, blk, S1, r, tt, rowtw, v);
#pragma unroll
        for (int c = 0; c < 8; ++c)
        {
            ulonglong2 t2;
            t2.x = ct_finish<VAR>(v[2 * c], bc, pc);
            t2.y = ct_finish<VAR>(v[2 * c + 1], bc, pc);
            *reinterpret_cast<ulonglong2*>(lineB + ((c ^ sw) << 4)) = t2;
        }
    }
    else
    {
#pragma unroll
        for (int c = 0; c < 8; ++c)
        {
            const ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(lineB + ((c ^ sw) << 4));
            v[2 * c] = gs_prep<VAR>(t2.x);
            v[2 * c + 1] = gs_prep<VAR>(t2.y);
        }
        gs_round_b_lm<VAR>(v, blk, tt, bc);
#pragma unroll
        for (int c = 0; c < 8; ++c)
        {
            ulonglong2 t2;
            t2.x = v[2 * c];
            t2.y = v[2 * c + 1];
            *reinterpret_cast<ulonglong2*>(lineB + ((c ^ sw) << 4)) = t2;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = *reinterpret_cast<const u64*>(rowp + k * 128 + ((((tt >> 1) ^ (k & 7)) << 4) | ((tt & 1) << 3)));
        gs_round_a<VAR>(v, tw, S1, r, bc);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            *reinterpret_cast<u64*>(rowp + k * 128 + ((((tt >> 1) ^ (k & 7)) << 4) | ((tt & 1) << 3))) = v[k]; // lazy
    }
}

// Persistent form: the grid is a few CTAs per SM; each CTA walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... with two shared-memory buffers, so the
// TMA load of tile i+1 and the TMA store of tile i-1 overlap the arithmetic of
// tile i (the pass is otherwise a load -> compute -> store chain whose memory
// time and multiplier-pipe time add up instead of overlapping).
// ROWS rows per CTA (16 threads each): 16 -> 32 KiB tiles, 3 CTAs/SM; 8 -> 16 KiB tiles, 6 CTAs/SM.
template <bool INV, class Map, int ROWS>
__global__ void __launch_bounds__(ROWS * 16, HEON_NTT_MINBLOCKS * 16 / ROWS)
    ntt_row_pass_tma(Map map, const __grid_constant__ CUtensorMap tm_in,
                     const __grid_constant__ CUtensorMap tm_out, const u64* in_base, const u64* out_base,
                     const TwPair* __restrict__ tw_all, const TwPair* __restrict__ rowb_all,
                     const PrimeConst* __restrict__ pcs, int logn, bool first_pass, int variant,
                     long long n_tiles, const double* __restrict__ rowc_all)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    // 1024-byte alignment for the 128B swizzle; plain offset arithmetic keeps the
    // pointer in the shared address space (LDS/STS instead of generic LD/ST)
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S1 = logn - 8;
    const int tiles = (1 << S1) / ROWS;
    constexpr int kTileBytes = ROWS * 2048;
    const CUtensorMap* tmi = first_pass ? &tm_in : &tm_out;
    if (!first_pass)
        in_base = out_base;
    const int tt = threadIdx.x & 15;
    const int rl = threadIdx.x >> 4;

    auto tile_lines = [&](long long t, int& line_in, int& line_out, int& prime, int& tile_idx) {
        const long long z = t / tiles;
        tile_idx = (int) (t % tiles);
        const u64* in;
        u64* out;
        int aux;
        map.get(z, in, out, prime, aux);
        if (!first_pass)
            in = out;
        line_in = (int) ((in - in_base) >> 4) + tile_idx * (ROWS * 16);
        line_out = (int) ((out - out_base) >> 4) + tile_idx * (ROWS * 16);
    };

    if (threadIdx.x == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    long long t = blockIdx.x;
    // one tile per CTA (rowc_all != nullptr): the second buffer receives the tile's FP64
    // twiddles (32 KiB, contiguous) through the same barrier
    if (threadIdx.x == 0 && t < n_tiles)
    {
        int li, lo, pr, ti;
        tile_lines(t, li, lo, pr, ti);
        const bool twsm = !INV && rowc_all && pcs[pr].fp_var != 0;
        mbar_arrive_expect_tx(&bar[0], twsm ? 2 * kTileBytes : kTileBytes);
        tma_load_2d(buf0, tmi, &bar[0], 0, li);
        if (twsm)
            tma_load_1d(buf0 + kTileBytes, rowc_all + ((((long long) pr << S1) + ti * ROWS) << 8), kTileBytes,
                        &bar[0]);
    }
    for (int it = 0; t < n_tiles; ++it, t += gridDim.x)
    {
        const int b = it & 1;
        unsigned char* tile = buf0 + b * kTileBytes;
        if (threadIdx.x == 0)
        {
            // the other buffer was handed to a TMA store one iteration ago: wait until that
            // store has finished reading it, then prefetch the next tile into it
            tma_store_wait_read<0>();
            const long long tn = t + gridDim.x;
            if (tn < n_tiles)
            {
                int li, lo, pr, ti;
                tile_lines(tn, li, lo, pr, ti);
                mbar_arrive_expect_tx(&bar[b ^ 1], kTileBytes);
                tma_load_2d(buf0 + (b ^ 1) * kTileBytes, tmi, &bar[b ^ 1], 0, li);
            }
        }
        int line_in, line_out, prime, tile_idx;
        tile_lines(t, line_in, line_out, prime, tile_idx);
        const PrimeConst pc = pcs[prime];
        const TwPair* tw = tw_all + ((long long) prime << logn);
        const int r = tile_idx * ROWS + rl;
        const TwPair* blk = rowb_all + ((((long long) prime << S1) + r) << 8);
        unsigned char* rowp = tile + rl * 2048;
        const double* rowtw =
            rowc_all ? reinterpret_cast<const double*>(buf0 + kTileBytes) + rl * 256 : nullptr;
        mbar_wait(&bar[b], (it >> 1) & 1);

        if (pc.fp_var == 3)
            row_pass_tma_body<INV, 3>(rowp, pc, tw, blk, S1, r, tt, rowtw);
        else if (pc.fp_var == 4)
            row_pass_tma_body<INV, 4>(rowp, pc, tw, blk, S1, r, tt, rowtw);
        else if (INV || variant == 1 || !pc.nc_ok)
            row_pass_tma_body<INV, 1>(rowp, pc, tw, blk, S1, r, tt, nullptr);
        else
            row_pass_tma_body<INV, 2>(rowp, pc, tw, blk, S1, r, tt, nullptr);

        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            tma_store_2d(&tm_out, tile, 0, line_out);
            tma_store_commit();
        }
    }
    if (threadIdx.x == 0)
        tma_store_wait_read<0>();
}


// Forward row pass, walking form (MapContig: polynomial z uses prime pl[z % count]): one CTA owns (row tile,
// prime slot) and walks the G polynomials z = y + count*(g*G + k) that share this prime.  The tile's FP64 twiddles
// are staged once per CTA; data tiles go through two buffers, so the load of polynomial k+1 and the store of
// polynomial k-1 overlap the arithmetic of polynomial k (the structure of k_row_mac's digit walk).  ROWS = 8:
// 16 KiB of twiddles + 2 x 16 KiB of data per CTA, four CTAs per SM.
template <class Map, int G>
__global__ void __launch_bounds__(128, 4)
    ntt_row_pass_tma_walk(Map map, const __grid_constant__ CUtensorMap tm_out, const u64* out_base,
                          const TwPair* __restrict__ tw_all, const TwPair* __restrict__ rowb_all,
                          const PrimeConst* __restrict__ pcs, int logn, int variant, long long n_polys,
                          const double* __restrict__ rowc_all)
{
    constexpr int ROWS = 8;
    constexpr int kTileBytes = ROWS * 2048;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ __align__(8) uint64_t twbar;
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* twbuf = buf0 + 2 * kTileBytes;
    const int S1 = logn - 8;
    const int tiles = (1 << S1) / ROWS;
    const int period = map.pl.count;
    const int tile_idx = blockIdx.x % tiles;
    const int y = (blockIdx.x / tiles) % period;
    const long long grp = blockIdx.x / ((long long) tiles * period);
    const int tt = threadIdx.x & 15, rl = threadIdx.x >> 4;
    auto poly = [&](int k) { return (long long) y + (long long) period * (grp * G + k); };
    int cnt = 0;
    while (cnt < G && poly(cnt) < n_polys)
        ++cnt;
    if (cnt == 0)
        return;
    const int prime = map.pl.idx[y];
    const PrimeConst pc = pcs[prime];
    const bool twsm = rowc_all && pc.fp_var != 0;
    auto line_of = [&](int k) {
        const u64* in;
        u64* out;
        int pr, aux;
        map.get(poly(k), in, out, pr, aux);
        return (int) ((out - out_base) >> 4) + tile_idx * (ROWS * 16);
    };
    if (threadIdx.x == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&twbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_arrive_expect_tx(&bar[0], kTileBytes);
        tma_load_2d(buf0, &tm_out, &bar[0], 0, line_of(0));
        if (twsm)
        {
            mbar_arrive_expect_tx(&twbar, kTileBytes);
            tma_load_1d(twbuf, rowc_all + ((((long long) prime << S1) + tile_idx * ROWS) << 8), kTileBytes, &twbar);
        }
    }
    const TwPair* tw = tw_all + ((long long) prime << logn);
    const int r = tile_idx * ROWS + rl;
    const TwPair* blk = rowb_all + ((((long long) prime << S1) + r) << 8);
    const double* rowtw = twsm ? reinterpret_cast<const double*>(twbuf) + rl * 256 : nullptr;
    if (twsm)
        mbar_wait(&twbar, 0);
#pragma unroll 1
    for (int k = 0; k < cnt; ++k)
    {
        const int b = k & 1;
        unsigned char* tile = buf0 + b * kTileBytes;
        if (threadIdx.x == 0 && k + 1 < cnt)
        {
            tma_store_wait_read<0>(); // the other buffer: its store (polynomial k-1) has read it
            mbar_arrive_expect_tx(&bar[b ^ 1], kTileBytes);
            tma_load_2d(buf0 + (b ^ 1) * kTileBytes, &tm_out, &bar[b ^ 1], 0, line_of(k + 1));
        }
        mbar_wait(&bar[b], (k >> 1) & 1);
        unsigned char* rowp = tile + rl * 2048;
        if (pc.fp_var == 3)
            row_pass_tma_body<false, 3>(rowp, pc, tw, blk, S1, r, tt, rowtw);
        else if (pc.fp_var == 4)
            row_pass_tma_body<false, 4>(rowp, pc, tw, blk, S1, r, tt, rowtw);
        else if (variant == 1 || !pc.nc_ok)
            row_pass_tma_body<false, 1>(rowp, pc, tw, blk, S1, r, tt, nullptr);
        else
            row_pass_tma_body<false, 2>(rowp, pc, tw, blk, S1, r, tt, nullptr);
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            tma_store_2d(&tm_out, tile, 0, line_of(k));
            tma_store_commit();
        }
    }
    if (threadIdx.x == 0)
        tma_store_wait_read<0>();
}

// ---------------------------------------------------------------------------
// Fused forward transform: ONE persistent kernel runs the column tiles and the
// row tiles of the whole batch.  CTAs draw tickets in order; the ticket stream
// is  col(g0) col(g1) row(g0) col(g2) row(g1) ...  over groups of `G`
// polynomials, so the column-pass output of a group (lazy words, written in
// place into the destination) is still in L2 when its row tiles read it and
// the transform costs one DRAM read and one DRAM write per word instead of
// two of each.  A row tile waits on a per-polynomial counter of finished
// column tiles (release/acquire at GPU scope); tickets are handed out in order
// and column tickets of a group precede its row tickets, so every CTA a
// waiter depends on is already running: no deadlock.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int S, class Map>
__global__ void __launch_bounds__(256, HEON_NTT_MINBLOCKS)
    ntt_fwd_fused(Map map, const __grid_constant__ CUtensorMap tm_out, const u64* out_base,
                  const TwPair* __restrict__ tw_all, const TwPair* __restrict__ rowb_all,
                  const double* __restrict__ rowc_all, const PrimeConst* __restrict__ pcs,
                  const TwPair* __restrict__ inv_last, int variant, long long n_polys, int G, int* sync)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ long long s_ticket;
    constexpr int logn = S + 8;
    constexpr int tiles = (1 << S) / 16; // tiles per polynomial, both passes
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const long long seg = (long long) G * tiles; // tickets per segment
    const long long n_groups = (n_polys + G - 1) / G;
    const long long n_tickets = (2 * n_groups + 1) * seg; // col(0) + pairs {col(k), row(k-1)}, k = 1..n_groups
    int* ticket = sync;
    int* done = sync + 1;
    const int tt = threadIdx.x & 15;
    const int rl = threadIdx.x >> 4;
    unsigned phase = 0;

    if (threadIdx.x == 0)
    {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    for (;;)
    {
        if (threadIdx.x == 0)
        {
            tma_store_wait_read<0>(); // the previous row tile has left shared memory
            s_ticket = atomicAdd(ticket, 1);
        }
        __syncthreads();
        const long long t = s_ticket;
        __syncthreads();
        if (t >= n_tickets)
            break;
        // segment s: 0 -> col(0); odd s -> col((s+1)/2); even s >= 2 -> row(s/2 - 1)
        const long long sgm = t / seg;
        const int w = (int) (t % seg);
        const bool is_row = sgm >= 2 && (sgm & 1) == 0;
        const long long g = sgm == 0 ? 0 : is_row ? sgm / 2 - 1 : (sgm + 1) / 2;
        const long long z = g * G + w / tiles;
        const int tile = w % tiles;
        if (g >= n_groups || z >= n_polys)
            continue;
        const u64* in;
        u64* out;
        int prime, aux;
        map.get(z, in, out, prime, aux);
        const PrimeConst pc = pcs[prime];
        const TwPair* tw = tw_all + ((long long) prime << logn);
        if (!is_row)
        {
            u64* sm = reinterpret_cast<u64*>(buf0);
            if (pc.fp_var == 3)
                col_pass_body<S, false, 3>(map, in, out, prime, pc, tw, inv_last, tile, true, aux, sm, z);
            else if (pc.fp_var == 4)
                col_pass_body<S, false, 4>(map, in, out, prime, pc, tw, inv_last, tile, true, aux, sm, z);
            else if (variant == 1 || !pc.nc_ok)
                col_pass_body<S, false, 1>(map, in, out, prime, pc, tw, inv_last, tile, true, aux, sm, z);
            else
                col_pass_body<S, false, 2>(map, in, out, prime, pc, tw, inv_last, tile, true, aux, sm, z);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0)
                atomicAdd(done + z, 1);
        }
        else
        {
            const bool twsm = rowc_all && pc.fp_var != 0;
            if (threadIdx.x == 0)
            {
                while (ld_acquire_gpu(done + z) < tiles)
                    __nanosleep(64);
                fence_proxy_async_all(); // generic-proxy writes of other CTAs -> this CTA's TMA read
                const int line = (int) ((out - out_base) >> 4) + tile * 256;
                mbar_arrive_expect_tx(&bar, twsm ? 2 * kRowTileBytes : kRowTileBytes);
                tma_load_2d(buf0, &tm_out, &bar, 0, line);
                if (twsm)
                    tma_load_1d(buf0 + kRowTileBytes, rowc_all + ((((long long) prime << S) + tile * 16) << 8),
                                kRowTileBytes, &bar);
            }
            const int r = tile * 16 + rl;
            const TwPair* blk = rowb_all + ((((long long) prime << S) + r) << 8);
            unsigned char* rowp = buf0 + rl * 2048;
            const double* rowtw = twsm ? reinterpret_cast<const double*>(buf0 + kRowTileBytes) + rl * 256 : nullptr;
            mbar_wait(&bar, phase & 1);
            ++phase;
            if (pc.fp_var == 3)
                row_pass_tma_body<false, 3>(rowp, pc, tw, blk, S, r, tt, rowtw);
            else if (pc.fp_var == 4)
                row_pass_tma_body<false, 4>(rowp, pc, tw, blk, S, r, tt, rowtw);
            else if (variant == 1 || !pc.nc_ok)
                row_pass_tma_body<false, 1>(rowp, pc, tw, blk, S, r, tt, nullptr);
            else
                row_pass_tma_body<false, 2>(rowp, pc, tw, blk, S, r, tt, nullptr);
            fence_proxy_async_smem();
            __syncthreads();
            if (threadIdx.x == 0)
            {
                const int line = (int) ((out - out_base) >> 4) + tile * 256;
                tma_store_2d(&tm_out, buf0, 0, line);
                tma_store_commit();
            }
        }
    }
    if (threadIdx.x == 0)
        tma_store_wait_read<0>();
}


// ---------------------------------------------------------------------------
// Pipelined fused forward transform for N = 2^16 (the BASELINE ring size).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0  producer : walks the ticket stream, issues the TMA load of every tile into a ring
//                      of kPipeStages 32 KiB shared-memory stages (full[] barriers);
//   warp 1  storer   : waits until a stage has been computed (comp[]), issues its TMA store,
//                      frees the stage (empty[]) and publishes finished column tiles;
//   2 x 8 consumer warps: two groups of 256 threads, each transforming one tile at a time out
//                      of registers, reading and writing the stage in place.
// Loads, stores and arithmetic of different tiles overlap freely; nothing on the arithmetic
// path waits for DRAM.  The ticket stream is  col(g0) col(g1) row(g0) col(g2) row(g1) ...  over
// groups of G polynomials (tickets are dealt round-robin to the CTAs), so the column-pass output
// of a group is still in L2 when its row tiles read it back: one DRAM read and one DRAM write
// per word.  The producer holds a row tile back until the 16 column tiles of its polynomial have
// been stored (per-polynomial counter, release/acquire at GPU scope).  The smallest unfinished
// ticket only depends on smaller tickets, so the scheme cannot deadlock.
//
// Column tile: 16 columns x 256 rows through a 3-D tensor map {16 words, 16 lines, rows} with
// box {16, 1, 256}; row tile: 16 rows = 256 consecutive 128-byte lines (2-D map); both land with
// the 128-byte swizzle, element e of line l at  l*128 + ((e>>1) ^ (l&7))*16 + (e&1)*8.
// ---------------------------------------------------------------------------
constexpr int kPipeStages = 6;
constexpr int kPipeGroups = 2;
constexpr int kPipeThreads = 64 + 256 * kPipeGroups;

template <int VAR, class Map, bool SMTW = false>
__device__ __forceinline__ void pipe_col_tile(unsigned char* tile, const Map& map, int prime, const PrimeConst& pc,
                                              const TwPair* __restrict__ tw, int aux, int tid, int bar_id)
{
    const BflyConst bc = make_bc(pc);
    const int c = tid & 15, tt = tid >> 4;
    u64 v[16];
    unsigned char* pa = tile + tt * 128 + ((((c >> 1) ^ (tt & 7)) << 4) | ((c & 1) << 3)); // rows tt + 16k
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
        u64 x = *reinterpret_cast<const u64*>(pa + k * 2048);
        if (Map::kXform)
            x = map.xform(x, prime, pc, aux);
        if (MapDoubleAux<Map>::value && VAR >= 3 && aux)
            v[k] = x;
        else
            v[k] = ct_prep<VAR>(x, bc, Map::kLazyIn);
    }
    ct_round_a<VAR, 0, SMTW>(v, tw, 0, 0, bc);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        *reinterpret_cast<u64*>(pa + k * 2048) = v[k];
    named_bar_sync(bar_id, 256);
    unsigned char* pb = tile + tt * 2048 + ((c & 1) << 3); // rows 16*tt + k
#pragma unroll
    for (int k = 0; k < 16; ++k)
        v[k] = *reinterpret_cast<const u64*>(pb + k * 128 + (((c >> 1) ^ (k & 7)) << 4));
    ct_round_b<8, VAR, 0, SMTW>(v, tw, 0, 0, tt, bc);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        *reinterpret_cast<u64*>(pb + k * 128 + (((c >> 1) ^ (k & 7)) << 4)) = v[k]; // lazy, finished by the row tile
}

// FP64 row tile: the thread's 15 last-four-stage twiddles (bare doubles, compact table) are
// requested before the first four stages run, so their L2 latency hides behind arithmetic.
template <int VAR>
__device__ __forceinline__ void pipe_row_tile_fp(unsigned char* rowp, const PrimeConst& pc,
                                                 const double* __restrict__ rowc, int tt)
{
    const BflyConst bc = make_bc(pc);
    double twb[15];
#pragma unroll
    for (int e = 0; e < 15; ++e)
        twb[e] = __ldg(rowc + 16 + e * 16 + tt);
    double twa[15];
#pragma unroll
    for (int e = 0; e < 15; ++e)
        twa[e] = __ldg(rowc + e);
    u64 v[16];
    unsigned char* lineB = rowp + tt * 128;
    const int sw = tt & 7;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        v[k] = *reinterpret_cast<const u64*>(rowp + k * 128 + ((((tt >> 1) ^ (k & 7)) << 4) | ((tt & 1) << 3)));
    ct_round_a_sm<VAR>(v, twa, bc);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        *reinterpret_cast<u64*>(rowp + k * 128 + ((((tt >> 1) ^ (k & 7)) << 4) | ((tt & 1) << 3))) = v[k];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c)
    {
        const ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(lineB + ((c ^ sw) << 4));
        v[2 * c] = t2.x;
        v[2 * c + 1] = t2.y;
    }
    ct_round_a_sm<VAR>(v, twb, bc);
#pragma unroll
    for (int c = 0; c < 8; ++c)
    {
        ulonglong2 t2;
        t2.x = ct_finish<VAR>(v[2 * c], bc, pc);
        t2.y = ct_finish<VAR>(v[2 * c + 1], bc, pc);
        *reinterpret_cast<ulonglong2*>(lineB + ((c ^ sw) << 4)) = t2;
    }
}

struct PipeTicket {
    long long z;
    int tile;
    bool is_row, valid, end;
};

// ticket -> work item (segment s: 0 -> col(0); odd s -> col((s+1)/2); even s >= 2 -> row(s/2 - 1))
__device__ __forceinline__ PipeTicket pipe_decode(long long t, long long n_polys, int G, long long n_groups)
{
    PipeTicket r;
    const long long seg = (long long) G * 16;
    r.end = t >= (2 * n_groups + 1) * seg;
    const long long sgm = t / seg;
    const int w = (int) (t % seg);
    r.is_row = sgm >= 2 && (sgm & 1) == 0;
    const long long g = sgm == 0 ? 0 : r.is_row ? sgm / 2 - 1 : (sgm + 1) / 2;
    r.z = g * G + w / 16;
    r.tile = w % 16;
    r.valid = !r.end && g < n_groups && r.z < n_polys;
    return r;
}

template <class Map>
__global__ void __launch_bounds__(kPipeThreads, 1)
    ntt16_fwd_pipe(Map map, const __grid_constant__ CUtensorMap tm_in_col,
                   const __grid_constant__ CUtensorMap tm_out_col,
                   const __grid_constant__ CUtensorMap tm_out_row, const u64* in_base, const u64* out_base,
                   const TwPair* __restrict__ tw_all, const TwPair* __restrict__ rowb_all,
                   const double* __restrict__ rowc_all, const PrimeConst* __restrict__ pcs, int variant,
                   long long n_polys, int G, int* done)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kPipeStages], comp[kPipeStages], empty[kPipeStages];
    __shared__ long long item[kPipeStages];
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int logn = 16, S = 8;
    const long long n_groups = (n_polys + G - 1) / G;
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < kPipeStages; ++s)
        {
            mbar_init(&full[s], 1);
            mbar_init(&comp[s], 256);
            mbar_init(&empty[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 0)
    {
        // ---------------- producer ----------------
        if (threadIdx.x != 0)
            return;
        long long t = blockIdx.x;
        int sentinels = 0;
        for (long long i = 0;; ++i)
        {
            const int s = (int) (i % kPipeStages);
            if (i >= kPipeStages)
                mbar_wait(&empty[s], (unsigned) ((i / kPipeStages - 1) & 1));
            PipeTicket k = pipe_decode(t, n_polys, G, n_groups);
            while (!k.end && !k.valid)
            {
                t += gridDim.x;
                k = pipe_decode(t, n_polys, G, n_groups);
            }
            if (k.end)
            {
                item[s] = -1;
                mbar_arrive(&full[s]);
                if (++sentinels == kPipeGroups)
                    break;
                continue;
            }
            t += gridDim.x;
            item[s] = (k.z << 8) | (k.tile << 1) | (k.is_row ? 1 : 0);
            const u64* in;
            u64* out;
            int prime, aux;
            map.get(k.z, in, out, prime, aux);
            unsigned char* stage = buf0 + s * kRowTileBytes;
            if (k.is_row)
            {
                while (ld_acquire_gpu(done + k.z) < 16)
                    __nanosleep(32);
                fence_proxy_async_all(); // other CTAs' tile stores -> this CTA's TMA read
                mbar_arrive_expect_tx(&full[s], kRowTileBytes);
                tma_load_2d(stage, &tm_out_row, &full[s], 0, (int) ((out - out_base) >> 4) + k.tile * 256);
            }
            else
            {
                mbar_arrive_expect_tx(&full[s], kRowTileBytes);
                tma_load_3d(stage, &tm_in_col, &full[s], 0, k.tile, (int) ((in - in_base) >> 8));
            }
        }
        return;
    }
    if (warp == 1)
    {
        // ---------------- storer ----------------
        if (threadIdx.x != 32)
            return;
        long long pend[3] = {-1, -1, -1};
        int sentinels = 0;
        long long n_store = 0;
        auto publish = [&](long long z) {
            if (z >= 0)
            {
                fence_proxy_async_all();
                __threadfence();
                atomicAdd(done + z, 1);
            }
        };
        for (long long i = 0;; ++i)
        {
            const int s = (int) (i % kPipeStages);
            if (!mbar_test(&comp[s], (unsigned) ((i / kPipeStages) & 1)))
            {
                // nothing to store right now: finish the stores in flight and publish their column
                // tiles (a row tile somewhere may be waiting for exactly these), then block
                tma_store_wait_all<0>();
                for (int j = 0; j < 3; ++j)
                {
                    publish(pend[j]);
                    pend[j] = -1;
                }
                mbar_wait(&comp[s], (unsigned) ((i / kPipeStages) & 1));
            }
            const long long it = item[s];
            if (it < 0)
            {
                if (++sentinels == kPipeGroups)
                    break;
                continue;
            }
            const long long z = it >> 8;
            const int tile = (int) ((it >> 1) & 127);
            const bool is_row = it & 1;
            const u64* in;
            u64* out;
            int prime, aux;
            map.get(z, in, out, prime, aux);
            unsigned char* stage = buf0 + s * kRowTileBytes;
            if (is_row)
                tma_store_2d(&tm_out_row, stage, 0, (int) ((out - out_base) >> 4) + tile * 256);
            else
                tma_store_3d(&tm_out_col, stage, 0, tile, (int) ((out - out_base) >> 8));
            tma_store_commit();
            tma_store_wait_read<0>();
            mbar_arrive(&empty[s]);
            // stores older than the two most recent ones are complete: publish their column tiles
            tma_store_wait_all<2>();
            publish(pend[n_store % 3]);
            pend[n_store % 3] = is_row ? -1 : z;
            ++n_store;
        }
        tma_store_wait_all<0>();
        for (int j = 0; j < 3; ++j)
            publish(pend[j]);
        return;
    }
    // ---------------- consumers ----------------
    const int grp = (threadIdx.x - 64) >> 8;
    const int tid = (threadIdx.x - 64) & 255;
    for (long long i = grp;; i += kPipeGroups)
    {
        const int s = (int) (i % kPipeStages);
        mbar_wait(&full[s], (unsigned) ((i / kPipeStages) & 1));
        const long long it = item[s];
        if (it < 0)
        {
            mbar_arrive(&comp[s]);
            break;
        }
        const long long z = it >> 8;
        const int tile = (int) ((it >> 1) & 127);
        const bool is_row = it & 1;
        const u64* in;
        u64* out;
        int prime, aux;
        map.get(z, in, out, prime, aux);
        const PrimeConst pc = pcs[prime];
        const TwPair* tw = tw_all + ((long long) prime << logn);
        unsigned char* stage = buf0 + s * kRowTileBytes;
        if (!is_row)
        {
            if (pc.fp_var == 3)
                pipe_col_tile<3>(stage, map, prime, pc, tw, aux, tid, 1 + grp);
            else if (pc.fp_var == 4)
                pipe_col_tile<4>(stage, map, prime, pc, tw, aux, tid, 1 + grp);
            else if (variant == 1 || !pc.nc_ok)
                pipe_col_tile<1>(stage, map, prime, pc, tw, aux, tid, 1 + grp);
            else
                pipe_col_tile<2>(stage, map, prime, pc, tw, aux, tid, 1 + grp);
        }
        else
        {
            const int tt = tid & 15, rl = tid >> 4;
            const int r = tile * 16 + rl;
            unsigned char* rowp = stage + rl * 2048;
            if (pc.fp_var == 3)
                pipe_row_tile_fp<3>(rowp, pc, rowc_all + ((((long long) prime << S) + r) << 8), tt);
            else if (pc.fp_var == 4)
                pipe_row_tile_fp<4>(rowp, pc, rowc_all + ((((long long) prime << S) + r) << 8), tt);
            else
            {
                const TwPair* blk = rowb_all + ((((long long) prime << S) + r) << 8);
                if (variant == 1 || !pc.nc_ok)
                    row_pass_tma_body<false, 1>(rowp, pc, tw, blk, S, r, tt, nullptr);
                else
                    row_pass_tma_body<false, 2>(rowp, pc, tw, blk, S, r, tt, nullptr);
            }
        }
        fence_proxy_async_smem();
        mbar_arrive(&comp[s]);
    }
}

// ---------------------------------------------------------------------------
// Column pass through TMA (N = 2^16): one CTA owns a tile of 16 columns x 256 rows (32 KiB).  A 3-D tensor
// map {16 words, 16 lines per row, rows} with box {16, 1, 256} brings the tile into shared memory with the
// 128-byte swizzle (one bulk copy instead of 16 strided 8-byte loads per thread through the LSU, which kept
// L1 80 % busy in the register-resident form), the eight stages run out of registers with one in-place
// transposition through the same buffer (pipe_col_tile), and the lazy words leave through a TMA store.
// The map's per-word transform (fused Method-I mod-up, divide-round stage one) is applied as the words are
// read from shared memory.
// ---------------------------------------------------------------------------
template <class Map>
__global__ void __launch_bounds__(256, 3)
    ntt_col_pass_tma(Map map, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                     const u64* in_base, const u64* out_base, const TwPair* __restrict__ tw_all,
                     const PrimeConst* __restrict__ pcs, int variant)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    unsigned char* buf = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const long long z = blockIdx.x >> 4;
    const int tile = blockIdx.x & 15;
    const u64* in;
    u64* out;
    int prime, aux;
    map.get(z, in, out, prime, aux);
    if (threadIdx.x == 0)
    {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_arrive_expect_tx(&bar, kRowTileBytes);
        tma_load_3d(buf, &tm_in, &bar, 0, tile, (int) ((in - in_base) >> 8));
    }
    const PrimeConst pc = pcs[prime];
    const TwPair* tw = tw_all + ((long long) prime << 16);
    mbar_wait(&bar, 0);
    if (pc.fp_var == 3)
        pipe_col_tile<3>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
    else if (pc.fp_var == 4)
        pipe_col_tile<4>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
    else if (variant == 1 || !pc.nc_ok)
        pipe_col_tile<1>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
    else
        pipe_col_tile<2>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        tma_store_3d(&tm_out, buf, 0, tile, (int) ((out - out_base) >> 8));
        tma_store_commit();
        tma_store_wait_read<0>();
    }
}

// The same tile, pipelined: one CTA walks G consecutive tiles of a polynomial through two 32 KiB buffers.  The
// load of tile i+1 is in flight while tile i is transformed, and the store of tile i leaves while tile i+1 is
// transformed, so the arithmetic of a CTA never waits for DRAM after its first tile (the single-tile form keeps
// the FP64 pipe 56-69 % busy: three resident CTAs cannot cover each other's load latency).
template <class Map, int G, int NB>
__global__ void __launch_bounds__(256, NB == 3 ? 2 : 3)
    ntt_col_pass_tma_pipe(Map map, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                          const u64* in_base, const u64* out_base, const TwPair* __restrict__ tw_all,
                          const PrimeConst* __restrict__ pcs, int variant)
{
    static_assert(16 % G == 0, "the tiles of a CTA belong to one polynomial");
    static_assert(NB == 2 || NB == 3, "two or three tile buffers");
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[NB];
    __shared__ __align__(8) uint64_t twbar;
    __shared__ __align__(16) TwPair twsm[256]; // the 255 twiddles of the eight column stages of this prime
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const long long z = blockIdx.x / (16 / G);
    const int tile0 = (blockIdx.x % (16 / G)) * G;
    const u64* in;
    u64* out;
    int prime, aux;
    map.get(z, in, out, prime, aux);
    const int row_in = (int) ((in - in_base) >> 8), row_out = (int) ((out - out_base) >> 8);
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int b = 0; b < NB; ++b)
            mbar_init(&bar[b], 1);
        mbar_init(&twbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_arrive_expect_tx(&bar[0], kRowTileBytes);
        tma_load_3d(buf0, &tm_in, &bar[0], 0, tile0, row_in);
        mbar_arrive_expect_tx(&twbar, sizeof(twsm));
        tma_load_1d(twsm, tw_all + ((long long) prime << 16), sizeof(twsm), &twbar);
    }
    const PrimeConst pc = pcs[prime];
    const TwPair* tw = twsm;
    mbar_wait(&twbar, 0);
    int b = 0, ph = 0; // buffer and mbarrier phase of tile i
#pragma unroll 1
    for (int i = 0; i < G; ++i)
    {
        unsigned char* buf = buf0 + b * kRowTileBytes;
        const int bn = (b + 1 == NB) ? 0 : b + 1;
        if (threadIdx.x == 0 && i + 1 < G)
        {
            // buffer bn was stored NB-1 tiles ago: with three buffers the store issued last may still be reading
            tma_store_wait_read<NB - 2>();
            mbar_arrive_expect_tx(&bar[bn], kRowTileBytes);
            tma_load_3d(buf0 + bn * kRowTileBytes, &tm_in, &bar[bn], 0, tile0 + i + 1, row_in);
        }
        mbar_wait(&bar[b], ph);
        if (pc.fp_var == 3)
            pipe_col_tile<3, Map, true>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
        else if (pc.fp_var == 4)
            pipe_col_tile<4, Map, true>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
        else if (variant == 1 || !pc.nc_ok)
            pipe_col_tile<1, Map, true>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
        else
            pipe_col_tile<2, Map, true>(buf, map, prime, pc, tw, aux, threadIdx.x, 1);
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            tma_store_3d(&tm_out, buf, 0, tile0 + i, row_out);
            tma_store_commit();
        }
        if (bn == 0)
            ph ^= 1;
        b = bn;
    }
    if (threadIdx.x == 0)
        tma_store_wait_read<0>();
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
    // measured on B200: +5.5 % for the in-place maps, -1.4 % for the fused mod-up map (64-byte chunks of a
    // source that 38 output limbs share) -> narrow CTAs only where the map does not transform its input
    if (S == 8 && (c.col_threads == 128 || (c.col_threads == 0 && (!Map::kXform || Map::kGather))))
    {
        // 8 columns per CTA: twice as many, half as large CTAs (finer-grained overlap of load / compute / store)
        ntt_col_pass<8, INV, Map, 128><<<grid * 2, 128, 0, st>>>(m, INV ? c.d_inv : c.d_fwd, c.d_pc, c.d_inv_last, c.logn,
                                                               first, c.ntt_variant);
        return;
    }
#define HEON_COL(SS)                                                                               \
    case SS:                                                                                       \
        ntt_col_pass<SS, INV, Map><<<grid, 256, 0, st>>>(m, INV ? c.d_inv : c.d_fwd, c.d_pc,       \
                                                         c.d_inv_last, c.logn, first, c.ntt_variant);             \
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
    ntt_row_pass<INV, Map><<<grid, 256, 0, st>>>(m, INV ? c.d_inv : c.d_fwd, c.d_pc, c.logn, first,
                                                 c.ntt_variant);
}

// ---- tensor maps -----------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p)
            throw std::runtime_error("cuTensorMapEncodeTiled is not available from this driver");
        return (EncodeTiledFn) p;
    }();
    return fn;
}

// Buffer viewed as `lines` rows of sixteen 64-bit words (128 B); box = 256 lines (one 16-row tile).
static CUtensorMap make_line_map(const u64* base, long long words, int box_lines = 256)
{
    CUtensorMap m;
    const cuuint64_t dims[2] = {16, (cuuint64_t) (words >> 4)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {16, (cuuint32_t) box_lines};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*) base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int) rc) + ")");
    return m;
}

// extent of the buffers a map touches (for the tensor-map bounds)
struct Extent {
    const u64* in_base;
    long long in_words;
    const u64* out_base;
    long long out_words;
    // buffer read by the FIRST pass of a forward transform (column tiles by TMA); nullptr when
    // its polynomials do not start on 2 KiB row boundaries relative to the base
    const u64* col_in_base = nullptr;
    long long col_in_words = 0;
};

template <bool INV, class Map, int ROWS>
static void launch_row_tma_rows(const Context& c, const Map& m, long long n_polys, bool first, const Extent& e,
                                cudaStream_t st)
{
    const int S = c.logn - 8;
    const long long n_tiles = n_polys * ((1 << S) / ROWS);
    // persistent (a few CTAs per SM walking tiles) or one tile per CTA
    const unsigned grid = c.ntt_persistent
                              ? (unsigned) std::min<long long>(n_tiles, (long long) c.num_sms * HEON_NTT_MINBLOCKS)
                              : (unsigned) n_tiles;
    const CUtensorMap tm_out = make_line_map(e.out_base, e.out_words, ROWS * 16);
    const CUtensorMap tm_in = first ? make_line_map(e.in_base, e.in_words, ROWS * 16) : tm_out;
    auto kfn = ntt_row_pass_tma<INV, Map, ROWS>;
    const int smem = 2 * ROWS * 2048 + 1024;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    LaunchScope scope(INV ? KC_NTT_INV_ROW : KC_NTT_FWD_ROW, st);
    kfn<<<grid, ROWS * 16, smem, st>>>(m, tm_in, tm_out, e.in_base, e.out_base, INV ? c.d_inv : c.d_fwd,
                                       INV ? c.d_inv_rowb : c.d_fwd_rowb, c.d_pc, c.logn, first, c.ntt_variant,
                                       n_tiles, (!INV && !c.ntt_persistent && c.use_fp64) ? c.d_fwd_rowc : nullptr);
}

// forward row pass walking same-prime polynomials (MapContig, second pass); false: not applicable
template <bool INV, class Map>
static bool launch_row_walk(const Context& c, const Map& m, long long n_polys, bool first, const Extent& e, cudaStream_t st)
{
    if constexpr (INV || !std::is_same<Map, MapContig>::value)
        return false;
    else
    {
        const int period = m.pl.count;
        if (first || c.row_walk == 0 || !c.use_fp64 || c.ntt_persistent || c.logn < 11 || period < 1)
            return false;
        const long long per_prime = (n_polys + period - 1) / period;
        if (per_prime < 4)
            return false;
        const int S = c.logn - 8;
        const int tiles = (1 << S) / 8;
        // walking serialises tiles that would otherwise run side by side: only when the GPU stays full
        // (four resident CTAs per SM, a few waves deep); small transforms are latency-bound and keep one tile per CTA
        const long long all_tiles = per_prime * period * tiles;
        if (c.row_walk < 0 && all_tiles < 16ll * c.num_sms)
            return false;
        const long long want = c.row_walk > 0 ? c.row_walk : all_tiles / (8ll * c.num_sms);
        const int G = want >= 8 ? 8 : want >= 4 ? 4 : 2; // the instantiated walks
        const long long groups = (per_prime + G - 1) / G;
        const long long grid = groups * period * tiles;
        if (grid > 0x7fffffffll)
            return false;
        const CUtensorMap tm_out = make_line_map(e.out_base, e.out_words, 8 * 16);
        const int smem = 3 * 8 * 2048 + 1024;
        LaunchScope scope(KC_NTT_FWD_ROW, st);
        auto go = [&](auto kfn) {
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            kfn<<<(unsigned) grid, 128, smem, st>>>(m, tm_out, e.out_base, c.d_fwd, c.d_fwd_rowb, c.d_pc, c.logn, c.ntt_variant,
                                                   n_polys, c.d_fwd_rowc);
        };
        if (G >= 8)
            go(ntt_row_pass_tma_walk<Map, 8>);
        else if (G >= 4)
            go(ntt_row_pass_tma_walk<Map, 4>);
        else
            go(ntt_row_pass_tma_walk<Map, 2>);
        return true;
    }
}

template <bool INV, class Map>
static void launch_row_tma(const Context& c, const Map& m, long long n_polys, bool first, const Extent& e,
                           cudaStream_t st)
{
    if (launch_row_walk<INV>(c, m, n_polys, first, e, st))
        return;
    if (c.row_tile == 4 && !c.ntt_persistent)
        launch_row_tma_rows<INV, Map, 4>(c, m, n_polys, first, e, st);
    else if (c.row_tile == 8 && !c.ntt_persistent)
        launch_row_tma_rows<INV, Map, 8>(c, m, n_polys, first, e, st);
    else
        launch_row_tma_rows<INV, Map, 16>(c, m, n_polys, first, e, st);
}


// rows-of-2-KiB view for the column tiles: {16 words, 16 lines per row, rows}, box {16, 1, 256}
static CUtensorMap make_col_map(const u64* base, long long words)
{
    CUtensorMap m;
    const cuuint64_t dims[3] = {16, 16, (cuuint64_t) (words >> 8)};
    const cuuint64_t strides[2] = {128, 2048};
    const cuuint32_t box[3] = {16, 1, 256};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*) base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled (column map) failed (" + std::to_string((int) rc) + ")");
    return m;
}

// `ein`: extent of the buffer the FIRST pass reads (may differ from e.in_base, which describes
// what the row pass reads); every polynomial must start a multiple of 256 words from the bases.
template <class Map>
static bool launch_fwd_pipe(const Context& c, const Map& m, long long n_polys, const u64* in_base,
                            long long in_words, const Extent& e, cudaStream_t st)
{
    if (c.logn != 16 || !c.ntt_pipe || !c.use_fp64 || n_polys < 1)
        return false;
    if ((reinterpret_cast<uintptr_t>(in_base) | reinterpret_cast<uintptr_t>(e.out_base)) & 127)
        return false;
    const int G = std::max(1, c.ntt_group);
    int* done = nullptr;
    const size_t sync_bytes = (size_t) n_polys * sizeof(int);
    if (cudaMallocAsync(&done, sync_bytes, st) != cudaSuccess)
        return false;
    cudaMemsetAsync(done, 0, sync_bytes, st);
    const CUtensorMap tm_in_col = make_col_map(in_base, in_words);
    const CUtensorMap tm_out_col = make_col_map(e.out_base, e.out_words);
    const CUtensorMap tm_out_row = make_line_map(e.out_base, e.out_words);
    const int smem = kPipeStages * kRowTileBytes + 1024;
    const long long n_tiles = n_polys * 32;
    const unsigned grid = (unsigned) std::min<long long>((n_tiles + 3) / 4, c.num_sms);
    auto kfn = ntt16_fwd_pipe<Map>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    {
        LaunchScope scope(KC_NTT_FWD_COL, st);
        kfn<<<grid, kPipeThreads, smem, st>>>(m, tm_in_col, tm_out_col, tm_out_row, in_base, e.out_base, c.d_fwd,
                                              c.d_fwd_rowb, c.d_fwd_rowc, c.d_pc, c.ntt_variant, n_polys, G, done);
    }
    cudaFreeAsync(done, st);
    return true;
}

template <class Map>
static bool launch_fwd_fused(const Context& c, const Map& m, long long n_polys, const Extent& e, cudaStream_t st)
{
    const int S = c.logn - 8;
    if (S < 5 || !c.ntt_fused)
        return false;
    const int tiles = (1 << S) / 16;
    // group = polynomials whose column-pass output waits in L2 for its row tiles (~8 MiB)
    int G = (int) std::max<long long>(1, (8ll << 20) / (8ll << c.logn));
    const long long n_groups = (n_polys + G - 1) / G;
    const long long n_tickets = (2 * n_groups + 1) * (long long) G * tiles;
    if (n_tickets + (long long) c.num_sms * 8 > 0x7fffffffll)
        return false;
    int* sync = nullptr;
    const size_t sync_bytes = (size_t) (n_polys + 1) * sizeof(int);
    if (cudaMallocAsync(&sync, sync_bytes, st) != cudaSuccess)
        return false;
    cudaMemsetAsync(sync, 0, sync_bytes, st);
    const CUtensorMap tm_out = make_line_map(e.out_base, e.out_words);
    const int smem = 2 * kRowTileBytes + 1024;
    const unsigned grid = (unsigned) std::min<long long>(n_tickets, (long long) c.num_sms * HEON_NTT_MINBLOCKS);
    {
        LaunchScope scope(KC_NTT_FWD_COL, st);
#define HEON_FUSED(SS)                                                                                      \
    case SS: {                                                                                              \
        auto kfn = ntt_fwd_fused<SS, Map>;                                                                  \
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                       \
        kfn<<<grid, 256, smem, st>>>(m, tm_out, e.out_base, c.d_fwd, c.d_fwd_rowb,                          \
                                     c.use_fp64 ? c.d_fwd_rowc : nullptr, c.d_pc, c.d_inv_last,             \
                                     c.ntt_variant, n_polys, G, sync);                                      \
        break;                                                                                              \
    }
        switch (S)
        {
            HEON_FUSED(5)
            HEON_FUSED(6)
            HEON_FUSED(7)
            HEON_FUSED(8)
        }
#undef HEON_FUSED
    }
    cudaFreeAsync(sync, st);
    return true;
}

// forward column pass through TMA tiles (N = 2^16, polynomials on 2 KiB row boundaries); false: not applicable
template <class Map>
static bool launch_col_tma(const Context& c, const Map& m, long long n_polys, const Extent& e, cudaStream_t st)
{
    if constexpr (Map::kGather)
        return false;
    else
    {
        // measured on B200 (C3-II, per op): register-resident LSU form 91.9 us, one TMA tile per CTA 94.7 us,
        // pipelined walk over 8 tiles with the stage twiddles in shared memory 77.5 us (two buffers, 3 CTAs/SM;
        // three buffers at 2 CTAs/SM: 82.4 us)
        if (c.col_tma == 0 || !c.use_tma || c.logn != 16 || !e.col_in_base || n_polys < 1)
            return false;
        if ((reinterpret_cast<uintptr_t>(e.col_in_base) | reinterpret_cast<uintptr_t>(e.out_base)) & 127)
            return false;
        if (n_polys * 16 > 0x7fffffffll)
            return false;
        const CUtensorMap tm_in = make_col_map(e.col_in_base, e.col_in_words);
        const CUtensorMap tm_out = make_col_map(e.out_base, e.out_words);
        LaunchScope scope(KC_NTT_FWD_COL, st);
        auto go = [&](auto kfn, int G, int nb) {
            const int smem = (G > 1 ? nb : 1) * kRowTileBytes + 1024;
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            kfn<<<(unsigned) (n_polys * 16 / G), 256, smem, st>>>(m, tm_in, tm_out, e.col_in_base, e.out_base, c.d_fwd, c.d_pc,
                                                                c.ntt_variant);
        };
        // tiles per CTA: enough CTAs to fill the GPU three deep before the walk is lengthened
        const long long tiles = n_polys * 16;
        const int want = c.col_tma_tiles > 0 ? c.col_tma_tiles : (tiles >= 16 * 444 ? 8 : tiles >= 4 * 444 ? 4 : 1);
        const bool three = c.col_tma_bufs == 3;
        if (want >= 16)
            three ? go(ntt_col_pass_tma_pipe<Map, 16, 3>, 16, 3) : go(ntt_col_pass_tma_pipe<Map, 16, 2>, 16, 2);
        else if (want >= 8)
            three ? go(ntt_col_pass_tma_pipe<Map, 8, 3>, 8, 3) : go(ntt_col_pass_tma_pipe<Map, 8, 2>, 8, 2);
        else if (want >= 4)
            three ? go(ntt_col_pass_tma_pipe<Map, 4, 3>, 4, 3) : go(ntt_col_pass_tma_pipe<Map, 4, 2>, 4, 2);
        else if (want >= 2)
            go(ntt_col_pass_tma_pipe<Map, 2, 2>, 2, 2);
        else
            go(ntt_col_pass_tma<Map>, 1, 1);
        return true;
    }
}

template <class Map>
static void run_ntt(const Context& c, const Map& m, long long n_polys, bool inverse, const Extent& e,
                    cudaStream_t st, bool col_only = false)
{
    if (n_polys <= 0)
        return;
    if (col_only)
    {
        // first n-8 stages only: the row stages run inside the fused inner-product kernel (k_row_mac)
        if (!launch_col_tma(c, m, n_polys, e, st))
            launch_col<false>(c, m, n_polys, true, st);
        return;
    }
    // TMA needs 16-byte aligned bases; fall back to the LSU row pass otherwise
    const bool tma = c.use_tma && ((reinterpret_cast<uintptr_t>(e.in_base) | reinterpret_cast<uintptr_t>(e.out_base)) & 15) == 0;
    if (!inverse)
    {
        if constexpr (!Map::kGather)
        {
            if (tma && e.col_in_base && launch_fwd_pipe(c, m, n_polys, e.col_in_base, e.col_in_words, e, st))
                return;
            if (tma && launch_fwd_fused(c, m, n_polys, e, st))
                return;
        }
        if (!launch_col_tma(c, m, n_polys, e, st))
            launch_col<false>(c, m, n_polys, true, st);
        if (tma)
            launch_row_tma<false>(c, m, n_polys, false, e, st);
        else
            launch_row<false>(c, m, n_polys, false, st);
    }
    else
    {
        if (tma)
            launch_row_tma<true>(c, m, n_polys, true, e, st);
        else
            launch_row<true>(c, m, n_polys, true, st);
        launch_col<true>(c, m, n_polys, false, st);
    }
}

} // namespace heon
