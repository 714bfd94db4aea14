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

#ifndef HEON_NTT_MINBLOCKS
#define HEON_NTT_MINBLOCKS 3
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
    static constexpr bool kLazyIn = false;
    __device__ __forceinline__ u64 xform(u64 x, int prime, const PrimeConst& pc, int) const
    {
        x = mod_add(x, half, plast);
        x = reduce_u64(x, pc);
        return mod_sub(x, half_mod[prime], pc.p);
    }
};

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

// Column pass body: S stages on columns (stride 256 words).  T = 2^S/16 threads
// cooperate on one column, C = 256/T adjacent columns per CTA.
template <int S, bool INV, int VAR, class Map>
__device__ __forceinline__ void col_pass_body(const Map& map, const u64* in, u64* out, int prime,
                                              const PrimeConst& pc, const TwPair* __restrict__ tw,
                                              const TwPair* __restrict__ inv_last, int tile,
                                              bool first_pass, int aux, u64* sm)
{
    constexpr int T = (1 << S) / 16;
    constexpr int C = 256 / T;
    const BflyConst bc = make_bc(pc);
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
                x = map.xform(x, prime, pc, aux);
            v[k] = ct_prep<VAR>(x, bc, Map::kLazyIn);
        }
        ct_round_a<VAR>(v, tw, 0, 0, bc);
        if constexpr (S > 4)
        {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                sm[(tt + T * k) * C + c] = v[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k)
                v[k] = sm[(16 * tt + k) * C + c];
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
        gs_round_a_final<VAR>(v, tw, bc, ninv, wninv);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            out[(long long) (tt + T * k) * 256 + col] = v[k];
    }
}

template <int S, bool INV, class Map>
__global__ void __launch_bounds__(256, HEON_NTT_MINBLOCKS) ntt_col_pass(Map map, const TwPair* __restrict__ tw_all,
                                                    const PrimeConst* __restrict__ pcs,
                                                    const TwPair* __restrict__ inv_last, int logn,
                                                    bool first_pass, int variant)
{
    constexpr int T = (1 << S) / 16;
    constexpr int C = 256 / T;
    __shared__ u64 sm[(S > 4) ? (1 << S) * C : 1];
    const int tiles = T; // 256 / C
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
    if (!INV && pc.fp_var == 3)
        col_pass_body<S, INV, INV ? 1 : 3>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm);
    else if (!INV && pc.fp_var == 4)
        col_pass_body<S, INV, INV ? 1 : 4>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm);
    else if (INV || variant == 1 || !pc.nc_ok)
        col_pass_body<S, INV, 1>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm);
    else
        col_pass_body<S, INV, 2>(map, in, out, prime, pc, tw, inv_last, tile, first_pass, aux, sm);
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
            v[k] = t2.x;
            v[k + 1] = t2.y;
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
    if (!INV && pc.fp_var == 3)
        row_pass_body<INV, INV ? 1 : 3>(rin, rout, pc, tw, S1, r, tt, srow);
    else if (!INV && pc.fp_var == 4)
        row_pass_body<INV, INV ? 1 : 4>(rin, rout, pc, tw, S1, r, tt, srow);
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
            v[2 * c] = t2.x;
            v[2 * c + 1] = t2.y;
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
template <bool INV, class Map>
__global__ void __launch_bounds__(256, HEON_NTT_MINBLOCKS)
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
    const int tiles = (1 << S1) / 16;
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
        line_in = (int) ((in - in_base) >> 4) + tile_idx * 256;
        line_out = (int) ((out - out_base) >> 4) + tile_idx * 256;
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
        mbar_arrive_expect_tx(&bar[0], twsm ? 2 * kRowTileBytes : kRowTileBytes);
        tma_load_2d(buf0, tmi, &bar[0], 0, li);
        if (twsm)
            tma_load_1d(buf0 + kRowTileBytes, rowc_all + ((((long long) pr << S1) + ti * 16) << 8), kRowTileBytes,
                        &bar[0]);
    }
    for (int it = 0; t < n_tiles; ++it, t += gridDim.x)
    {
        const int b = it & 1;
        unsigned char* tile = buf0 + b * kRowTileBytes;
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
                mbar_arrive_expect_tx(&bar[b ^ 1], kRowTileBytes);
                tma_load_2d(buf0 + (b ^ 1) * kRowTileBytes, tmi, &bar[b ^ 1], 0, li);
            }
        }
        int line_in, line_out, prime, tile_idx;
        tile_lines(t, line_in, line_out, prime, tile_idx);
        const PrimeConst pc = pcs[prime];
        const TwPair* tw = tw_all + ((long long) prime << logn);
        const int r = tile_idx * 16 + rl;
        const TwPair* blk = rowb_all + ((((long long) prime << S1) + r) << 8);
        unsigned char* rowp = tile + rl * 2048;
        const double* rowtw =
            rowc_all ? reinterpret_cast<const double*>(buf0 + kRowTileBytes) + rl * 256 : nullptr;
        mbar_wait(&bar[b], (it >> 1) & 1);

        if (!INV && pc.fp_var == 3)
            row_pass_tma_body<INV, INV ? 1 : 3>(rowp, pc, tw, blk, S1, r, tt, rowtw);
        else if (!INV && pc.fp_var == 4)
            row_pass_tma_body<INV, INV ? 1 : 4>(rowp, pc, tw, blk, S1, r, tt, rowtw);
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
static CUtensorMap make_line_map(const u64* base, long long words)
{
    CUtensorMap m;
    const cuuint64_t dims[2] = {16, (cuuint64_t) (words >> 4)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {16, 256};
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
};

template <bool INV, class Map>
static void launch_row_tma(const Context& c, const Map& m, long long n_polys, bool first, const Extent& e,
                           cudaStream_t st)
{
    const int S = c.logn - 8;
    const long long n_tiles = n_polys * ((1 << S) / 16);
    // persistent (a few CTAs per SM walking tiles) or one tile per CTA
    const unsigned grid = c.ntt_persistent
                              ? (unsigned) std::min<long long>(n_tiles, (long long) c.num_sms * HEON_NTT_MINBLOCKS)
                              : (unsigned) n_tiles;
    const CUtensorMap tm_out = make_line_map(e.out_base, e.out_words);
    const CUtensorMap tm_in = first ? make_line_map(e.in_base, e.in_words) : tm_out;
    static bool attr_set[2] = {false, false};
    auto kfn = ntt_row_pass_tma<INV, Map>;
    const int smem = 2 * kRowTileBytes + 1024;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    (void) attr_set;
    LaunchScope scope(INV ? KC_NTT_INV_ROW : KC_NTT_FWD_ROW, st);
    kfn<<<grid, 256, smem, st>>>(m, tm_in, tm_out, e.in_base, e.out_base, INV ? c.d_inv : c.d_fwd,
                                 INV ? c.d_inv_rowb : c.d_fwd_rowb, c.d_pc, c.logn, first, c.ntt_variant,
                                 n_tiles, (!INV && !c.ntt_persistent && c.use_fp64) ? c.d_fwd_rowc : nullptr);
}

template <class Map>
static void run_ntt(const Context& c, const Map& m, long long n_polys, bool inverse, const Extent& e,
                    cudaStream_t st)
{
    if (n_polys <= 0)
        return;
    // TMA needs 16-byte aligned bases; fall back to the LSU row pass otherwise
    const bool tma = c.use_tma && ((reinterpret_cast<uintptr_t>(e.in_base) | reinterpret_cast<uintptr_t>(e.out_base)) & 15) == 0;
    if (!inverse)
    {
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

void launch_ntt(const Context& c, const u64* src, u64* dst, long long n_polys, const PrimeList& pl,
                bool inverse, cudaStream_t st)
{
    MapContig m{src, dst, pl, c.logn};
    const long long w = n_polys << c.logn;
    run_ntt(c, m, n_polys, inverse, Extent{src, w, dst, w}, st);
}

void launch_ntt_scattered(const Context& c, u64* base, const long long* d_offsets, int n_polys,
                          int prime, bool inverse, long long extent_words, bool aligned, cudaStream_t st)
{
    MapScatter m{base, d_offsets, prime};
    // offsets that are not multiples of 16 words cannot be addressed in 128-byte lines
    Extent e{aligned ? base : base + 1, extent_words, aligned ? base : base + 1, extent_words};
    run_ntt(c, m, n_polys, inverse, e, st);
}

void launch_ntt_strided(const Context& c, u64* base, long long bstride, int per_batch, int first,
                        long long batch, const PrimeList& pl, bool inverse, cudaStream_t st)
{
    MapStrided m{base, bstride, per_batch, first, pl, c.logn};
    const long long w = (batch - 1) * bstride + ((long long) (first + per_batch) << c.logn);
    const u64* b0 = (bstride & 15) ? base + 1 : base; // odd strides: no line addressing -> LSU path
    run_ntt(c, m, batch * per_batch, inverse, Extent{b0, w, b0, w}, st);
}

void launch_ntt_strided_copy(const Context& c, const u64* src, long long src_bstride, u64* dst,
                             int per_batch, long long batch, const PrimeList& pl, bool inverse,
                             cudaStream_t st)
{
    MapStridedCopy m{src, dst, src_bstride, per_batch, pl, c.logn};
    const long long wi = (batch - 1) * src_bstride + ((long long) per_batch << c.logn);
    const long long wo = (batch * per_batch) << c.logn;
    const u64* s0 = (src_bstride & 15) ? src + 1 : src;
    run_ntt(c, m, batch * per_batch, inverse, Extent{s0, wi, dst, wo}, st);
}

void launch_modup1_ntt(const Context& c, const u64* coef, long long coef_bstride, u64* out, int L,
                       int depth, long long batch, cudaStream_t st)
{
    const int Qpl = L + c.P_size;
    MapModUpI m{coef, out, coef_bstride, L, Qpl, depth, c.logn, c.d_pc};
    const long long wo = (batch * L * Qpl) << c.logn;
    run_ntt(c, m, batch * L * Qpl, false, Extent{out, wo, out, wo}, st);
}

void launch_divround1_ntt(const Context& c, const u64* src, long long bstride, long long cstride,
                          u64* out, int Lout, u64 half, u64 plast, const u64* d_half_mod,
                          long long batch, cudaStream_t st)
{
    MapDivRoundOne m{src, out, bstride, cstride, Lout, c.logn, half, plast, d_half_mod};
    const long long wo = (batch * 2 * Lout) << c.logn;
    run_ntt(c, m, batch * 2 * Lout, false, Extent{out, wo, out, wo}, st);
}

} // namespace heon
