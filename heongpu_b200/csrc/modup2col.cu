// Method-II mod-up fused with the column pass of the forward NTT, source tiles staged once per CTA
// (k_modup2_col).
//
// reference: base_conversion_DtoQtilde_relin_leveled_kernel (src/lib/kernel/switchkey.cu:985-1046) followed by
// the first kernel(s) of GPU_NTT_Modulus_Ordered_Inplace (thirdparty/GPU-NTT ntt.cu, ForwardCore).
//
// The separate kernels write the converted digits tmp[b][d][Q'_l][N] (the largest buffer of the key switch)
// and the column pass reads them back.  MapModUpII (ntt_modup2.cu) removed that round trip by converting
// inside the column-pass load, but every output word then gathers its I_j source words from L2 ahead of
// the butterflies, and the latency cost more than the traffic saved.  Here one CTA owns
// (ciphertext b, digit i, column tile) and walks ALL targets of the digit:
//   * the I_j source tiles (16 columns x 256 rows of the digit's limbs, coefficient domain) arrive once
//     by TMA (3-D tensor map, 128-byte swizzle) and are turned in place into the partial words
//     x_j * Mi_inv_j mod q_j; the fp32 correction r (the reference's float sequence, operation for
//     operation) is kept per coefficient;
//   * for each target prime t_y the thread rebuilds its 16 input words  sum_j partial_j * M_{j,y} - r*prod_y
//     from shared memory (FP64 products for primes below 2^50, Shoup products otherwise), runs the eight
//     column stages out of registers with one in-place transposition, and the lazy words leave through a
//     TMA store;
//   * two 256-thread groups take alternate targets, each with its own output tile and twiddle buffers, so
//     one group's barrier / store drain overlaps the other's arithmetic.
// The converted digits never exist in memory and the source words cross HBM once per digit instead of once
// per target.  Every step is exact, so the words equal those of the separate kernels (and the reference's).
#include "ntt_impl.cuh"

namespace heon {

struct Mu2ColDigits {
    short I_loc[65], I_j[65];
};

constexpr int kMcTile = kRowTileBytes; // 32 KiB

// per-target constants of the CTA's digit, staged in shared memory once (no global loads between targets)
struct __align__(16) McRec {
    TwPair m[4];   // conversion factors {M, companion}: doubles {M, RN(M/t)} or Shoup pairs
    u64 rp[5];     // r * prod mod t for r = 0..I_j
    int y, prime;  // limb slot in Q'_l, prime index
    PrimeConst pc;
};

// the thread's 16 input words of target y in ct_prep<VAR> form; src = the thread's slot in source tile 0
template <int VAR>
__device__ __forceinline__ void mc_gather(const unsigned char* src, int ij, bool fp_src, bool fp, const McRec& rc,
                                          unsigned long long rpack, u64 (&v)[16], const BflyConst& c)
{
    u64 rp[5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
    {
        const u64 t = rc.rp[r];
        rp[r] = (VAR >= 3 && fp) ? d2u(fp_from_u64(t)) : t;
    }
    if (VAR >= 3 && fp)
    {
        double acc[16]; // |acc| <= I_j * 0.57 t
#pragma unroll
        for (int k = 0; k < 16; ++k)
            acc[k] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < ij)
            {
                const TwPair t = rc.m[j];
                const double w = u2d(t.w), wi = u2d(t.ws);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    acc[k] = __dadd_rn(acc[k], fp_mulmod(u2d(*reinterpret_cast<const u64*>(src + j * kMcTile + k * 2048)), w, wi, c.dnp));
            }
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            const unsigned r = (unsigned) (rpack >> (4 * k)) & 15u;
            const u64 x = r == 0 ? rp[0] : r == 1 ? rp[1] : r == 2 ? rp[2] : r == 3 ? rp[3] : rp[4];
            // |acc - rp| < 3.3 t; balanced residue |v| <= t/2 (+1), as in MapModUpII::gather16
            v[k] = d2u(fp_reduce(__dsub_rn(acc[k], u2d(x)), c.dpinv, c.dnp));
        }
        return;
    }
    const u64 p4 = 4 * c.p, np = c.np;
    u64 a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
        a[k] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (j < ij)
        {
            const TwPair t = rc.m[j];
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                const u64 raw = *reinterpret_cast<const u64*>(src + j * kMcTile + k * 2048);
                const u64 pj = fp_src ? (u64) __double2ll_rn(u2d(raw)) : raw;
                a[k] = csub(a[k] + shoup_lazy_ptx(pj, t.w, t.ws, np), p4);
            }
        }
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
        const unsigned r = (unsigned) (rpack >> (4 * k)) & 15u;
        const u64 rv = r == 0 ? rp[0] : r == 1 ? rp[1] : r == 2 ? rp[2] : r == 3 ? rp[3] : rp[4];
        u64 x = csub(csub(a[k], 2 * c.p), c.p);
        x = mod_sub(x, rv, c.p);
        v[k] = ct_prep<VAR>(x, c, false);
    }
}

// one target: conversion, eight column stages, TMA store of the lazy words
template <int VAR>
__device__ __forceinline__ void mc_target(const unsigned char* src, unsigned char* otile, const TwPair* twsm, int ij,
                                          bool fp_src, const McRec& rc, unsigned long long rpack, int tid, int bar_id)
{
    const PrimeConst& pc = rc.pc;
    const BflyConst bc = make_bc(pc);
    const int c = tid & 15, tt = tid >> 4;
    u64 v[16];
    mc_gather<VAR>(src, ij, fp_src, fp_src && pc.fp_var != 0, rc, rpack, v, bc);
    ct_round_a<VAR, 0, true>(v, twsm, 0, 0, bc);
    // the output tile: its previous store must have finished reading it
    if (tid == 0)
        tma_store_wait_read<0>();
    named_bar_sync(bar_id, 256);
    unsigned char* pa = otile + tt * 128 + ((((c >> 1) ^ (tt & 7)) << 4) | ((c & 1) << 3)); // rows tt + 16k
#pragma unroll
    for (int k = 0; k < 16; ++k)
        *reinterpret_cast<u64*>(pa + k * 2048) = v[k];
    named_bar_sync(bar_id, 256);
    unsigned char* pb = otile + tt * 2048 + ((c & 1) << 3); // rows 16*tt + k
#pragma unroll
    for (int k = 0; k < 16; ++k)
        v[k] = *reinterpret_cast<const u64*>(pb + k * 128 + (((c >> 1) ^ (k & 7)) << 4));
    ct_round_b<8, VAR, 0, true>(v, twsm, 0, 0, tt, bc);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        *reinterpret_cast<u64*>(pb + k * 128 + (((c >> 1) ^ (k & 7)) << 4)) = v[k]; // lazy, finished by the row stages
}

__global__ void __launch_bounds__(512, 1)
    k_modup2_col(const __grid_constant__ CUtensorMap tm_coef, const __grid_constant__ CUtensorMap tm_out,
                 const PrimeConst* __restrict__ pcs, const TwPair* __restrict__ mi_inv,
                 const TwPair* __restrict__ bc_pair, const u64* __restrict__ rprod, const TwPair* __restrict__ tw_all,
                 const Mu2ColDigits dig, int d, int Qpl, int L, int depth, int coef_rbs, int ij_max,
                 unsigned long long dfp_mask, int variant)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t srcbar;
    __shared__ __align__(8) uint64_t twbar[2][2];
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ssrc = buf0;                           // ij_max source tiles
    unsigned char* sout = buf0 + ij_max * kMcTile;        // one output tile per group
    TwPair* stw = reinterpret_cast<TwPair*>(sout + 2 * kMcTile); // [group][buffer][256]
    unsigned char* srs = reinterpret_cast<unsigned char*>(stw + 4 * 256); // r per coefficient [16][256]
    McRec* rec = reinterpret_cast<McRec*>(srs + 4096);                    // [n_targets]
    const int tile = blockIdx.x;
    const int dg = blockIdx.y;
    const long long b = blockIdx.z;
    const int ij = dig.I_j[dg], iloc = dig.I_loc[dg];
    const int grp = threadIdx.x >> 8, tid = threadIdx.x & 255;
    const bool fp_src = (dfp_mask >> dg) & 1;
    const int n_targets = Qpl - ij;
    // t-th target of the digit: the special limbs first (their 60-bit primes cost 2.75 x an FP64 target: spread
    // over both groups), then the Q limbs with the digit's own limbs skipped
    const int Ksp = Qpl - L;
    auto target = [&](int t) {
        if (t < Ksp)
            return L + t;
        t -= Ksp;
        return t < iloc ? t : t + ij;
    };

    if (threadIdx.x == 0)
    {
        mbar_init(&srcbar, 1);
        mbar_init(&twbar[0][0], 1);
        mbar_init(&twbar[0][1], 1);
        mbar_init(&twbar[1][0], 1);
        mbar_init(&twbar[1][1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_arrive_expect_tx(&srcbar, ij * kMcTile);
        for (int j = 0; j < ij; ++j)
            tma_load_3d(ssrc + j * kMcTile, &tm_coef, &srcbar, 0, tile, (int) (b * coef_rbs) + ((iloc + j) << 8));
    }
    if (tid == 0 && grp < n_targets)
    {
        const int prime = level_prime(target(grp), L, depth);
        mbar_arrive_expect_tx(&twbar[grp][0], 256 * sizeof(TwPair));
        tma_load_1d(stw + (grp * 2 + 0) * 256, tw_all + ((long long) prime << 16), 256 * sizeof(TwPair), &twbar[grp][0]);
    }
    // per-target constants (overlaps the source tiles' flight)
    for (int t = threadIdx.x; t < n_targets; t += 512)
    {
        McRec rc;
        rc.y = target(t);
        rc.prime = level_prime(rc.y, L, depth);
        rc.pc = pcs[rc.prime];
        const TwPair* m = bc_pair + (long long) iloc * Qpl + (long long) rc.y * ij;
        for (int j = 0; j < 4; ++j)
            rc.m[j] = j < ij ? m[j] : TwPair{0, 0};
        for (int r = 0; r < 5; ++r)
            rc.rp[r] = r <= ij ? rprod[((long long) r * d + dg) * Qpl + rc.y] : 0;
        rec[t] = rc;
    }
    const int c = tid & 15, tt = tid >> 4;
    const unsigned slot = tt * 128 + ((((c >> 1) ^ (tt & 7)) << 4) | ((c & 1) << 3)); // rows tt + 16k: + k*2048
    mbar_wait(&srcbar, 0);
    // partial words in place, fp32 correction per coefficient (k_modup2_prep's sequence); group g takes k = 8g..8g+7
    {
        float r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            r[k] = 0.f;
        for (int j = 0; j < ij; ++j)
        {
            const PrimeConst pj = pcs[iloc + j];
            const TwPair mi = mi_inv[iloc + j];
            const float mod = __ull2float_rn(pj.p);
            unsigned char* sp = ssrc + j * kMcTile + slot + grp * 8 * 2048;
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                const u64 x = *reinterpret_cast<const u64*>(sp + k * 2048);
                const u64 pw = csub(shoup_mul_lazy(x, mi.w, mi.ws, pj.p), pj.p);
                r[k] = __fadd_rn(r[k], __fdiv_rn(__ull2float_rn(pw), mod));
                *reinterpret_cast<u64*>(sp + k * 2048) = fp_src ? d2u(fp_from_u64(pw)) : pw;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            srs[(grp * 8 + k) * 256 + tid] = (unsigned char) (unsigned) roundf(r[k]);
    }
    __syncthreads();
    unsigned long long rpack = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        rpack |= (unsigned long long) (srs[k * 256 + tid] & 15u) << (4 * k);

    unsigned char* otile = sout + grp * kMcTile;
    const int bar_id = 1 + grp;
    const int out_row0 = (int) (((b * d + dg) * Qpl) << 8);
    int it = 0;
#pragma unroll 1
    for (int t = grp; t < n_targets; t += 2, ++it)
    {
        const McRec& rc = rec[t];
        const int y = rc.y;
        const unsigned fpv = rc.pc.fp_var, ncok = rc.pc.nc_ok;
        const int tb = it & 1;
        if (tid == 0 && t + 2 < n_targets)
        {
            // the other twiddle buffer served the previous target, finished behind a group barrier
            mbar_arrive_expect_tx(&twbar[grp][tb ^ 1], 256 * sizeof(TwPair));
            tma_load_1d(stw + (grp * 2 + (tb ^ 1)) * 256, tw_all + ((long long) rec[t + 2].prime << 16), 256 * sizeof(TwPair),
                        &twbar[grp][tb ^ 1]);
        }
        const TwPair* twsm = stw + (grp * 2 + tb) * 256;
        mbar_wait(&twbar[grp][tb], (it >> 1) & 1);
        if (fpv == 3)
            mc_target<3>(ssrc + slot, otile, twsm, ij, fp_src, rc, rpack, tid, bar_id);
        else if (fpv == 4)
            mc_target<4>(ssrc + slot, otile, twsm, ij, fp_src, rc, rpack, tid, bar_id);
        else if (variant == 1 || !ncok)
            mc_target<1>(ssrc + slot, otile, twsm, ij, fp_src, rc, rpack, tid, bar_id);
        else
            mc_target<2>(ssrc + slot, otile, twsm, ij, fp_src, rc, rpack, tid, bar_id);
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 256);
        if (tid == 0)
        {
            tma_store_3d(&tm_out, otile, 0, tile, out_row0 + (y << 8));
            tma_store_commit();
        }
    }
    if (tid == 0)
        tma_store_wait_read<0>();
}

static int mc_smem_bytes(int ij_max, int Qpl)
{
    return (ij_max + 2) * kMcTile + 4 * 256 * (int) sizeof(TwPair) + 4096 + Qpl * (int) sizeof(McRec) + 1024;
}

// true when the fused mod-up + column pass can serve this key switch (column stages only: the row stages
// follow inside k_row_mac)
bool modup2_col_available(const Context& c, int depth, const u64* coef, long long coef_bs, const u64* tmp, int batch,
                          bool own_stashed, bool col_only)
{
    if (c.method != 2 || !c.modup_col || !c.use_tma || c.logn != 16 || !own_stashed || !col_only || c.P_size > 4)
        return false;
    const LevelTablesII& t = c.lvl2[depth];
    if (t.d > 64)
        return false;
    int ij_max = 1;
    for (int i = 0; i < t.d; ++i)
    {
        if (t.I_j[i] > 4)
            return false;
        ij_max = std::max(ij_max, t.I_j[i]);
    }
    if (mc_smem_bytes(ij_max, c.Q_size - depth + c.P_size) > 227 * 1024)
        return false;
    if ((coef_bs & 255) != 0 || ((reinterpret_cast<uintptr_t>(coef) | reinterpret_cast<uintptr_t>(tmp)) & 127) != 0)
        return false;
    const int L = c.Q_size - depth, Qpl = L + c.P_size;
    // one CTA per SM: not for grids that leave half the GPU idle
    if (c.modup_col == 1 && 32ll * t.d * batch < c.num_sms)
        return false;
    const long long rows_out = ((long long) batch * t.d * Qpl) << 8;
    const long long rows_in = ((long long) (batch - 1) * coef_bs + ((long long) L << 16)) >> 8;
    return rows_out < 0x7fffffffll && rows_in < 0x7fffffffll;
}

// tmp[b][i][y] = column stages of NTT(mod-up of digit i to prime y) for every target y outside digit i
void launch_modup2_col(const Context& c, const u64* coef, long long coef_bs, u64* tmp, int depth, long long batch,
                       cudaStream_t st)
{
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    const LevelTablesII& t = c.lvl2[depth];
    Mu2ColDigits dig;
    unsigned long long dfp_mask = 0;
    int ij_max = 1;
    for (int i = 0; i < t.d; ++i)
    {
        dig.I_loc[i] = (short) t.I_loc[i];
        dig.I_j[i] = (short) t.I_j[i];
        ij_max = std::max(ij_max, t.I_j[i]);
        bool dfp = c.use_fp64 && t.I_j[i] <= 4; // same rule as the table upload (context.cu)
        for (int j = 0; j < t.I_j[i]; ++j)
            dfp = dfp && c.mod[t.I_loc[i] + j].bit <= 50;
        if (dfp)
            dfp_mask |= 1ull << i;
    }
    const long long wi = (batch - 1) * coef_bs + ((long long) L << c.logn);
    const long long wo = (batch * t.d * Qpl) << c.logn;
    const CUtensorMap tm_coef = make_col_map(coef, wi);
    const CUtensorMap tm_out = make_col_map(tmp, wo);
    const int smem = mc_smem_bytes(ij_max, Qpl);
    cudaFuncSetAttribute(k_modup2_col, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    LaunchScope scope(KC_MODUP2, st);
    k_modup2_col<<<dim3(16, t.d, (unsigned) batch), 512, smem, st>>>(tm_coef, tm_out, c.d_pc, t.d_mi_inv_pair,
                                                                    t.d_base_change_pair, t.d_rprod, c.d_fwd, dig, t.d, Qpl,
                                                                    L, depth, (int) (coef_bs >> 8), ij_max, dfp_mask,
                                                                    c.ntt_variant);
}

} // namespace heon
