// Key-switch inner product fused behind the forward row pass (k_row_mac).
#include "ntt_impl.cuh"

namespace heon {

// ---------------------------------------------------------------------------
// Key-switch inner product fused behind the forward row pass.
//
//   acc[b][c][y] = sum_i NTT(tmp[b][i][y]) (.) key[i][c][prime(y)]        (c = 0, 1)
//
// replaces  ntt_row_pass_tma (last eight stages of d*Q' transforms)  +  k_keyswitch_mac
// (reference: the trailing kernel of GPU_NTT_Modulus_Ordered_Inplace, ntt.cu:3106-3255, followed by
// keyswitch_multiply_accumulate_leveled[_method_II]_kernel, switchkey.cu:164-398).  One CTA owns
// (ciphertext b, limb y, a tile of ROWS rows) and walks the d digits: the column-pass words of digit i
// arrive by TMA (2-D tensor map, 128-byte swizzle), the eight row stages run out of registers, and the
// finished words are multiplied into the two key tiles (brought in by TMA with the same swizzle, so the
// register layout of the transform is also the conflict-free layout of the key read) and accumulated
// in registers.  Only the two accumulator tiles are stored.  The transformed digits -- d*Q'*N words,
// the largest buffer of the operator -- are never written back and never read again.
//
// Arithmetic (FP = true, primes below 2^50): the transform leaves integer-valued doubles x, |x| < 2^51;
// a key word k < p becomes a double exactly, T = x*k mod p comes from fp_mulmod with the quotient
// multiplier RN(k * RN(1/p)) (|T| <= p), and the terms are summed in one double per (coefficient,
// component): exact while |sum| < 2^53, so the sum is reduced every `red_period` digits.  The final
// word is canonical and equals the reference's per-term Barrett sum.  FP = false (58..61-bit primes):
// canonical words, 128-bit lazy integer accumulation, one reduction per output.
// The batch index is the fastest block coordinate: the CTAs that need the same key tiles run together
// and share them through L2 (the key crosses HBM once per batch).
// Digit-own limbs (Method II, see MapDigitSkip) hold canonical NTT-domain words already and skip the
// stages.
// ---------------------------------------------------------------------------
struct OwnLimbs {
    int d, own;
    short I_loc[65], I_j[65];
};
struct LimbList {
    unsigned char y[128];
};

template <bool FP, int ROWS>
__global__ void __launch_bounds__(ROWS * 16, FP ? 24 / ROWS : 16 / ROWS)
    k_row_mac(const __grid_constant__ CUtensorMap tm_tmp, const __grid_constant__ CUtensorMap tm_key,
              const __grid_constant__ CUtensorMap tm_out, const TwPair* __restrict__ tw_all,
              const TwPair* __restrict__ rowb_all, const double* __restrict__ rowc_all,
              const PrimeConst* __restrict__ pcs, const LimbList limb_list, int logn, int L,
              int Qpl, int Qp0, int depth, int variant, int red_period_lo, int red_period_hi, OwnLimbs own)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    constexpr int T = ROWS * 2048;
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sdata = buf0;
    unsigned char* skey0 = buf0 + T;
    unsigned char* skey1 = buf0 + 2 * T;
    unsigned char* stw = buf0 + 3 * T;
    const int S1 = logn - 8;
    const int lpp = 1 << (logn - 4); // 128-byte lines per polynomial
    const long long b = blockIdx.x;
    const int tile_idx = blockIdx.y;
    const int y = limb_list.y[blockIdx.z];
    const int prime = level_prime(y, L, depth);
    const PrimeConst pc = pcs[prime];
    const BflyConst bc = make_bc(pc);
    const int d = own.d;
    const int tt = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const int r = tile_idx * ROWS + rl;
    const int line0 = tile_idx * ROWS * 16;
    auto dline = [&](int i) { return (int) (((b * d + i) * Qpl + y) * lpp) + line0; };
    auto kline = [&](int i, int c) { return (int) ((((long long) i * 2 + c) * Qp0 + prime) * lpp) + line0; };

    if (threadIdx.x == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_arrive_expect_tx(&bar[0], FP ? 2 * T : T);
        tma_load_2d(sdata, &tm_tmp, &bar[0], 0, dline(0));
        if (FP)
            tma_load_1d(stw, rowc_all + ((((long long) prime << S1) + tile_idx * ROWS) << 8), T, &bar[0]);
        mbar_arrive_expect_tx(&bar[1], 2 * T);
        tma_load_2d(skey0, &tm_key, &bar[1], 0, kline(0, 0));
        tma_load_2d(skey1, &tm_key, &bar[1], 0, kline(0, 1));
    }
    const TwPair* tw = tw_all + ((long long) prime << logn);
    const TwPair* blk = rowb_all + ((((long long) prime << S1) + r) << 8);
    const double* rowtw = reinterpret_cast<const double*>(stw) + rl * 256;
    unsigned char* rowp = sdata + rl * 2048;
    const int sw = tt & 7;
    const unsigned lineoff = rl * 2048 + tt * 128;

    // accumulators: FP -> one double per (coefficient, component); integer -> 128 bits each
    double fa0[FP ? 16 : 1], fa1[FP ? 16 : 1];
    u64 il0[FP ? 1 : 16], ih0[FP ? 1 : 16], il1[FP ? 1 : 16], ih1[FP ? 1 : 16];
    if constexpr (FP)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            fa0[k] = fa1[k] = 0.0;
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            il0[k] = ih0[k] = il1[k] = ih1[k] = 0;
    }
    const int red_period = pc.fp_var == 3 ? red_period_lo : red_period_hi;
    int since_red = 0;

    for (int i = 0; i < d; ++i)
    {
        u64 v[16];
        mbar_wait(&bar[0], i & 1);
        const bool own_i = own.own && y < L && y >= own.I_loc[i] && y < own.I_loc[i] + own.I_j[i];
        if (own_i)
        {
            // canonical NTT-domain words (stashed from the input): no stages
#pragma unroll
            for (int c = 0; c < 8; ++c)
            {
                const ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(sdata + lineoff + ((c ^ sw) << 4));
                v[2 * c] = FP ? d2u(fp_from_u64(t2.x)) : t2.x;
                v[2 * c + 1] = FP ? d2u(fp_from_u64(t2.y)) : t2.y;
            }
        }
        else if constexpr (FP)
        {
            if (pc.fp_var == 3)
                row_fwd_stages<3>(rowp, bc, tw, blk, S1, r, tt, rowtw, v);
            else
                row_fwd_stages<4>(rowp, bc, tw, blk, S1, r, tt, rowtw, v);
        }
        else
        {
            if (variant == 1 || !pc.nc_ok)
            {
                row_fwd_stages<1>(rowp, bc, tw, blk, S1, r, tt, nullptr, v);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    v[k] = ct_finish<1>(v[k], bc, pc);
            }
            else
            {
                row_fwd_stages<2>(rowp, bc, tw, blk, S1, r, tt, nullptr, v);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    v[k] = ct_finish<2>(v[k], bc, pc);
            }
        }
        // every warp holds its words in registers: the data tile can take the next digit
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0 && i + 1 < d)
        {
            mbar_arrive_expect_tx(&bar[0], T);
            tma_load_2d(sdata, &tm_tmp, &bar[0], 0, dline(i + 1));
        }
        mbar_wait(&bar[1], i & 1);
        if constexpr (FP)
        {
            const double pinv = bc.dpinv, dnp = bc.dnp;
#pragma unroll
            for (int c = 0; c < 8; ++c)
            {
                const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(skey0 + lineoff + ((c ^ sw) << 4));
                const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(skey1 + lineoff + ((c ^ sw) << 4));
                const double x0 = u2d(v[2 * c]), x1 = u2d(v[2 * c + 1]);
                const double a0 = fp_from_u64(k0.x), a1 = fp_from_u64(k0.y);
                const double b0 = fp_from_u64(k1.x), b1 = fp_from_u64(k1.y);
                fa0[2 * c] = __dadd_rn(fa0[2 * c], fp_mulmod(x0, a0, __dmul_rn(a0, pinv), dnp));
                fa0[2 * c + 1] = __dadd_rn(fa0[2 * c + 1], fp_mulmod(x1, a1, __dmul_rn(a1, pinv), dnp));
                fa1[2 * c] = __dadd_rn(fa1[2 * c], fp_mulmod(x0, b0, __dmul_rn(b0, pinv), dnp));
                fa1[2 * c + 1] = __dadd_rn(fa1[2 * c + 1], fp_mulmod(x1, b1, __dmul_rn(b1, pinv), dnp));
            }
            if (++since_red >= red_period && i + 1 < d)
            {
                since_red = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                {
                    fa0[k] = fp_reduce(fa0[k], pinv, dnp);
                    fa1[k] = fp_reduce(fa1[k], pinv, dnp);
                }
            }
        }
        else
        {
#pragma unroll
            for (int c = 0; c < 8; ++c)
            {
                const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(skey0 + lineoff + ((c ^ sw) << 4));
                const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(skey1 + lineoff + ((c ^ sw) << 4));
                mac128(il0[2 * c], ih0[2 * c], v[2 * c], k0.x);
                mac128(il0[2 * c + 1], ih0[2 * c + 1], v[2 * c + 1], k0.y);
                mac128(il1[2 * c], ih1[2 * c], v[2 * c], k1.x);
                mac128(il1[2 * c + 1], ih1[2 * c + 1], v[2 * c + 1], k1.y);
            }
        }
        __syncthreads(); // the key tiles have been consumed
        if (threadIdx.x == 0 && i + 1 < d)
        {
            mbar_arrive_expect_tx(&bar[1], 2 * T);
            tma_load_2d(skey0, &tm_key, &bar[1], 0, kline(i + 1, 0));
            tma_load_2d(skey1, &tm_key, &bar[1], 0, kline(i + 1, 1));
        }
    }
    // canonical results leave through the two (now idle) key buffers
#pragma unroll
    for (int c = 0; c < 8; ++c)
    {
        ulonglong2 r0, r1;
        if constexpr (FP)
        {
            r0.x = fp_canon(fa0[2 * c], bc.dpinv, bc.dnp, bc.dp);
            r0.y = fp_canon(fa0[2 * c + 1], bc.dpinv, bc.dnp, bc.dp);
            r1.x = fp_canon(fa1[2 * c], bc.dpinv, bc.dnp, bc.dp);
            r1.y = fp_canon(fa1[2 * c + 1], bc.dpinv, bc.dnp, bc.dp);
        }
        else
        {
            r0.x = reduce_u128(il0[2 * c], ih0[2 * c], pc);
            r0.y = reduce_u128(il0[2 * c + 1], ih0[2 * c + 1], pc);
            r1.x = reduce_u128(il1[2 * c], ih1[2 * c], pc);
            r1.y = reduce_u128(il1[2 * c + 1], ih1[2 * c + 1], pc);
        }
        *reinterpret_cast<ulonglong2*>(skey0 + lineoff + ((c ^ sw) << 4)) = r0;
        *reinterpret_cast<ulonglong2*>(skey1 + lineoff + ((c ^ sw) << 4)) = r1;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        tma_store_2d(&tm_out, skey0, 0, (int) (((b * 2 + 0) * Qpl + y) * lpp) + line0);
        tma_store_2d(&tm_out, skey1, 0, (int) (((b * 2 + 1) * Qpl + y) * lpp) + line0);
        tma_store_commit();
        tma_store_wait_read<0>();
    }
}

// true when the fused row-pass + inner-product kernel can serve this key switch
bool row_mac_available(const Context& c, const u64* tmp, const u64* key, const u64* acc, int d)
{
    return c.use_tma && c.row_mac && c.logn >= 12 && d >= 1 && d <= 64 &&
           ((reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(key) | reinterpret_cast<uintptr_t>(acc)) & 15) == 0;
}

// acc[b][2][Qpl][N] = sum_i rowpass(tmp[b][i][y]) (.) key[i][c][prime(y)]; tmp holds column-pass output
// (lazy words), digit-own limbs (own_stashed) hold canonical NTT-domain words.
void launch_row_mac(const Context& c, const u64* tmp, const u64* key, u64* acc, int d, int depth, int batch,
                    bool own_stashed, const int* I_loc, const int* I_j, cudaStream_t st)
{
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    OwnLimbs own;
    own.d = d;
    own.own = own_stashed ? 1 : 0;
    for (int i = 0; i < d && i < 65; ++i)
    {
        own.I_loc[i] = (short) (own_stashed ? I_loc[i] : 0);
        own.I_j[i] = (short) (own_stashed ? I_j[i] : 0);
    }
    // limb slots by arithmetic: FP64 primes / integer primes (two launches, different register budgets)
    LimbList lfp, lint;
    int nfp = 0, nint = 0;
    for (int y = 0; y < Qpl; ++y)
    {
        const bool fp = c.use_fp64 && c.mod[level_prime(y, L, depth)].bit <= 50;
        if (fp)
            lfp.y[nfp++] = (unsigned char) y;
        else
            lint.y[nint++] = (unsigned char) y;
    }
    const long long wt = ((long long) batch * d * Qpl) << c.logn;
    const long long wk = ((long long) d * 2 * c.Qp) << c.logn;
    const long long wa = ((long long) batch * 2 * Qpl) << c.logn;
    const int rows = (c.row_mac_rows == 4) ? 4 : 8;
    const CUtensorMap tm_tmp = make_line_map(tmp, wt, rows * 16);
    const CUtensorMap tm_key = make_line_map(key, wk, rows * 16);
    const CUtensorMap tm_out = make_line_map(acc, wa, rows * 16);
    const int tiles = (1 << (c.logn - 8)) / rows;
    // |T| <= 1.3 p per term (lazy x up to 2.25 p, quotient rounded by FRND): the double accumulator stays exact
    // while (1.3 * terms + 1/2) * p < 2^53 -> every 5 terms for p < 2^50, every 40 for p < 2^47
    const int red_lo = 40, red_hi = 5;
    auto go = [&](auto kfn, int nl, const LimbList& list, int smem) {
        if (nl == 0)
            return;
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        LaunchScope scope(KC_ROW_MAC, st);
        kfn<<<dim3(batch, tiles, nl), rows * 16, smem, st>>>(tm_tmp, tm_key, tm_out, c.d_fwd, c.d_fwd_rowb, c.d_fwd_rowc, c.d_pc,
                                                        list, c.logn, L, Qpl, c.Qp, depth, c.ntt_variant, red_lo, red_hi, own);
    };
    if (rows == 8)
    {
        go(k_row_mac<true, 8>, nfp, lfp, 4 * 8 * 2048 + 1024);
        go(k_row_mac<false, 8>, nint, lint, 3 * 8 * 2048 + 1024);
    }
    else
    {
        go(k_row_mac<true, 4>, nfp, lfp, 4 * 4 * 2048 + 1024);
        go(k_row_mac<false, 4>, nint, lint, 3 * 4 * 2048 + 1024);
    }
}

// ---------------------------------------------------------------------------
// Method-II mod-down: forward row pass of the correction polynomials fused with the final combination
// (k_row_final).
//
//   Method II:  out[b][c][y] = (ct[b][c][y] if selected) + acc[b][c][y] * M_y + NTT(corr[b][c][y])        (y < L)
//   Method I :  out[b][c][y] = (ct[b][c][y] if selected) + (acc[b][c][y] - NTT(corr[b][c][y])) * P^-1     (M1)
//
// replaces  ntt_row_pass_tma_walk over the 2*L correction polynomials  +  k_moddown2_final / k_moddown1_stage2
// (reference: the tail of divide_round_lastq_extended_leveled_kernel, switchkey.cu:1255-1349, and the addition
// kernel that follows it in relinearize / rotate, ckks/operator.cu:1025-1154; Method I:
// divide_round_lastq_leveled_stage_two[_switchkey]_kernel, switchkey.cu:707-771).  One CTA owns (limb y, a tile of 4 rows) and walks G
// polynomials z = 2*b + c of that limb: the column-pass words of the correction arrive by TMA through two
// buffers, the tile's FP64 twiddles are staged once per CTA, the eight row stages run out of registers, and the
// finished words meet the accumulator tile and the ciphertext tile (TMA, same swizzle: the transform's
// register layout is their conflict-free layout).  The transformed corrections are never written back:
// 2*L*N words less to store and to read again per ciphertext.
// Every step is exact modular arithmetic on canonical inputs, so the stored words equal k_moddown2_final's.
// ---------------------------------------------------------------------------
template <bool FP, bool M1>
__global__ void __launch_bounds__(64, FP ? 5 : 4)
    k_row_final(const __grid_constant__ CUtensorMap tm_tmp, const __grid_constant__ CUtensorMap tm_acc,
                const __grid_constant__ CUtensorMap tm_ct, const __grid_constant__ CUtensorMap tm_out,
                const TwPair* __restrict__ tw_all, const TwPair* __restrict__ rowb_all,
                const double* __restrict__ rowc_all, const PrimeConst* __restrict__ pcs,
                const TwPair* __restrict__ mprod, const LimbList limb_list, int logn, int L, int Qpl, int G,
                int n_polys, int ct_lbs, int out_lbs, int variant, int add_mask)
{
    constexpr int ROWS = 4;
    constexpr int T = ROWS * 2048;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[3];
    unsigned char* buf0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sacc = buf0 + 2 * T;
    unsigned char* sct = buf0 + 3 * T;
    unsigned char* stw = buf0 + 4 * T;
    const int S1 = logn - 8;
    const int lpp = 1 << (logn - 4); // 128-byte lines per polynomial
    const int tile_idx = blockIdx.y;
    const int y = limb_list.y[blockIdx.z]; // limb slot = prime index (the Q primes of the level)
    const int z0 = blockIdx.x * G;
    const int cnt = min(G, n_polys - z0);
    if (cnt <= 0)
        return;
    const PrimeConst pc = pcs[y];
    const BflyConst bc = make_bc(pc);
    const TwPair m = mprod[y];
    const int tt = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const int r = tile_idx * ROWS + rl;
    const int line0 = tile_idx * ROWS * 16;
    auto tline = [&](int z) { return (z * L + y) * lpp + line0; };
    auto aline = [&](int z) { return (z * Qpl + y) * lpp + line0; };
    auto cline = [&](int z, int lbs) { return (z >> 1) * lbs + ((z & 1) * L + y) * lpp + line0; };
    auto adds = [&](int z) { return ((add_mask >> (z & 1)) & 1) != 0; };

    if (threadIdx.x == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_arrive_expect_tx(&bar[0], FP ? 2 * T : T);
        tma_load_2d(buf0, &tm_tmp, &bar[0], 0, tline(z0));
        if (FP)
            tma_load_1d(stw, rowc_all + ((((long long) y << S1) + tile_idx * ROWS) << 8), T, &bar[0]);
        mbar_arrive_expect_tx(&bar[2], adds(z0) ? 2 * T : T);
        tma_load_2d(sacc, &tm_acc, &bar[2], 0, aline(z0));
        if (adds(z0))
            tma_load_2d(sct, &tm_ct, &bar[2], 0, cline(z0, ct_lbs));
    }
    const TwPair* tw = tw_all + ((long long) y << logn);
    const TwPair* blk = rowb_all + ((((long long) y << S1) + r) << 8);
    const double* rowtw = reinterpret_cast<const double*>(stw) + rl * 256;
    const int sw = tt & 7;
    const unsigned lineoff = rl * 2048 + tt * 128;
    const double pinv = bc.dpinv, dnp = bc.dnp;
    const double Md = FP ? fp_from_u64(m.w) : 0.0;
    const double Mi = FP ? __dmul_rn(Md, pinv) : 0.0;

#pragma unroll 1
    for (int k = 0; k < cnt; ++k)
    {
        const int b = k & 1;
        const int z = z0 + k;
        unsigned char* sdata = buf0 + b * T;
        if (threadIdx.x == 0 && k + 1 < cnt)
        {
            // the other data buffer went into registers one iteration ago (a barrier has passed since)
            mbar_arrive_expect_tx(&bar[b ^ 1], T);
            tma_load_2d(buf0 + (b ^ 1) * T, &tm_tmp, &bar[b ^ 1], 0, tline(z + 1));
        }
        u64 v[16];
        mbar_wait(&bar[b], (k >> 1) & 1);
        unsigned char* rowp = sdata + rl * 2048;
        if constexpr (FP)
        {
            if (pc.fp_var == 3)
                row_fwd_stages<3>(rowp, bc, tw, blk, S1, r, tt, rowtw, v);
            else
                row_fwd_stages<4>(rowp, bc, tw, blk, S1, r, tt, rowtw, v);
        }
        else
        {
            if (variant == 1 || !pc.nc_ok)
            {
                row_fwd_stages<1>(rowp, bc, tw, blk, S1, r, tt, nullptr, v);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    v[q] = ct_finish<1>(v[q], bc, pc);
            }
            else
            {
                row_fwd_stages<2>(rowp, bc, tw, blk, S1, r, tt, nullptr, v);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    v[q] = ct_finish<2>(v[q], bc, pc);
            }
        }
        mbar_wait(&bar[2], k & 1);
        const bool add = adds(z);
#pragma unroll
        for (int c = 0; c < 8; ++c)
        {
            const unsigned off = lineoff + ((c ^ sw) << 4);
            const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(sacc + off);
            ulonglong2 t = {0, 0}, o;
            if (add)
                t = *reinterpret_cast<const ulonglong2*>(sct + off);
            if constexpr (FP)
            {
                double s0, s1;
                if constexpr (M1)
                {
                    // (acc - corr) * P^-1: |acc - x| <= 3.25 p; fp_mulmod wants |Y| < 2^51
                    s0 = __dsub_rn(fp_from_u64(a.x), u2d(v[2 * c]));
                    s1 = __dsub_rn(fp_from_u64(a.y), u2d(v[2 * c + 1]));
                    if (pc.fp_var != 3)
                    {
                        s0 = fp_reduce(s0, pinv, dnp);
                        s1 = fp_reduce(s1, pinv, dnp);
                    }
                    s0 = fp_mulmod(s0, Md, Mi, dnp);
                    s1 = fp_mulmod(s1, Md, Mi, dnp);
                }
                else
                {
                    // |x| <= 2.25 p (lazy transform output), |a*M mod p| <= 0.6 p, ct < p: below 2^52
                    s0 = __dadd_rn(u2d(v[2 * c]), fp_mulmod(fp_from_u64(a.x), Md, Mi, dnp));
                    s1 = __dadd_rn(u2d(v[2 * c + 1]), fp_mulmod(fp_from_u64(a.y), Md, Mi, dnp));
                }
                if (add)
                {
                    s0 = __dadd_rn(s0, fp_from_u64(t.x));
                    s1 = __dadd_rn(s1, fp_from_u64(t.y));
                }
                o.x = fp_canon(s0, pinv, dnp, bc.dp);
                o.y = fp_canon(s1, pinv, dnp, bc.dp);
            }
            else
            {
                const u64 p = pc.p;
                if constexpr (M1)
                {
                    o.x = csub(shoup_mul_lazy(mod_sub(a.x, v[2 * c], p), m.w, m.ws, p), p);
                    o.y = csub(shoup_mul_lazy(mod_sub(a.y, v[2 * c + 1], p), m.w, m.ws, p), p);
                }
                else
                {
                    o.x = mod_add(csub(shoup_mul_lazy(a.x, m.w, m.ws, p), p), v[2 * c], p);
                    o.y = mod_add(csub(shoup_mul_lazy(a.y, m.w, m.ws, p), p), v[2 * c + 1], p);
                }
                if (add)
                {
                    o.x = mod_add(t.x, o.x, p);
                    o.y = mod_add(t.y, o.y, p);
                }
            }
            *reinterpret_cast<ulonglong2*>(sacc + off) = o;
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            tma_store_2d(&tm_out, sacc, 0, cline(z, out_lbs));
            tma_store_commit();
            if (k + 1 < cnt)
            {
                tma_store_wait_read<0>(); // the accumulator buffer is the store's source
                mbar_arrive_expect_tx(&bar[2], adds(z + 1) ? 2 * T : T);
                tma_load_2d(sacc, &tm_acc, &bar[2], 0, aline(z + 1));
                if (adds(z + 1))
                    tma_load_2d(sct, &tm_ct, &bar[2], 0, cline(z + 1, ct_lbs));
            }
        }
    }
    if (threadIdx.x == 0)
        tma_store_wait_read<0>();
}

// true when the fused row-pass + final-combination kernel can serve this Method-II mod-down
bool row_final_available(const Context& c, const u64* tmp, const u64* acc, const u64* ct_in, long long ct_bs,
                         const u64* out, long long out_bs, int L, int batch)
{
    if (!(c.use_tma && c.row_final && c.logn >= 12))
        return false;
    uintptr_t a = reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(out);
    if (ct_in)
        a |= reinterpret_cast<uintptr_t>(ct_in);
    if ((a & 15) != 0 || ((ct_bs | out_bs) & 15) != 0 || L > 128)
        return false;
    // line indices are 32-bit
    const long long span = std::max(ct_bs, out_bs) * (long long) (batch - 1) + 2ll * L * c.n;
    const long long wacc = (long long) batch * 2 * (L + c.P_size) * c.n;
    return (span >> 4) < 0x7fffffffll && (wacc >> 4) < 0x7fffffffll;
}

// out[b][c][y] = (ct[b][c][y] if add_mask bit c) + acc[b][c][y]*M_y + rowpass(tmp[b][c][y]); tmp holds the
// column-pass output of the corrections (lazy words).
void launch_row_final(const Context& c, const u64* tmp, const u64* acc, const u64* ct_in, long long ct_bs, u64* out,
                      long long out_bs, int depth, int batch, int add_mask, cudaStream_t st)
{
    const bool m1 = c.method == 1;
    const int L = c.Q_size - depth, K = c.P_size, Qpl = L + K;
    LimbList lfp, lint;
    int nfp = 0, nint = 0;
    for (int y = 0; y < L; ++y)
    {
        const bool fp = c.use_fp64 && c.mod[y].bit <= 50;
        if (fp)
            lfp.y[nfp++] = (unsigned char) y;
        else
            lint.y[nint++] = (unsigned char) y;
    }
    if (!ct_in)
        add_mask = 0;
    const long long wt = ((long long) batch * 2 * L) << c.logn;
    const long long wa = ((long long) batch * 2 * Qpl) << c.logn;
    const long long wc = (long long) (batch - 1) * ct_bs + (2ll * L << c.logn);
    const long long wo = (long long) (batch - 1) * out_bs + (2ll * L << c.logn);
    constexpr int rows = 4;
    const CUtensorMap tm_tmp = make_line_map(tmp, wt, rows * 16);
    const CUtensorMap tm_acc = make_line_map(acc, wa, rows * 16);
    const CUtensorMap tm_out = make_line_map(out, wo, rows * 16);
    const CUtensorMap tm_ct = add_mask ? make_line_map(ct_in, wc, rows * 16) : tm_out;
    const int tiles = (1 << (c.logn - 8)) / rows;
    const int n_polys = 2 * batch;
    // walk up to 8 polynomials per CTA while the grid stays a few waves deep
    int G = 8;
    while (G > 2 && (long long) ((n_polys + G - 1) / G) * tiles * L < 8ll * 5 * c.num_sms)
        G >>= 1;
    if (c.row_final > 1)
        G = std::min(c.row_final, 64); // HEON_ROW_FINAL=G forces the walk length (tests: ragged last group)
    const int groups = (n_polys + G - 1) / G;
    const int smem = 5 * rows * 2048 + 1024;
    auto go = [&](auto kfn, int nl, const LimbList& list) {
        if (nl == 0)
            return;
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        LaunchScope scope(KC_MODDOWN, st);
        kfn<<<dim3(groups, tiles, nl), rows * 16, smem, st>>>(tm_tmp, tm_acc, tm_ct, tm_out, c.d_fwd, c.d_fwd_rowb, c.d_fwd_rowc,
                                                         c.d_pc, m1 ? c.d_lqm_pair : c.d_md2_M, list, c.logn, L, Qpl, G, n_polys,
                                                         (int) (ct_bs >> 4), (int) (out_bs >> 4), c.ntt_variant, add_mask);
    };
    if (m1)
    {
        go(k_row_final<true, true>, nfp, lfp);
        go(k_row_final<false, true>, nint, lint);
    }
    else
    {
        go(k_row_final<true, false>, nfp, lfp);
        go(k_row_final<false, false>, nint, lint);
    }
}

} // namespace heon
