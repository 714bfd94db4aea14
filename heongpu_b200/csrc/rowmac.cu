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

} // namespace heon
