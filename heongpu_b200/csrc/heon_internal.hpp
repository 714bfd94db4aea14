// Internal context of the B200 RNS engine (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "host_math.hpp"

namespace heon {

enum Scheme { SCHEME_BFV = 1, SCHEME_CKKS = 2 };

// Shoup pair: multiplier w (< p) and floor(w * 2^64 / p).
struct alignas(16) TwPair {
    u64 w;
    u64 ws;
};

// Per-prime constants used by the fused kernels.
struct PrimeConst {
    u64 p;
    u64 inv64; // floor(2^64 / p): Shoup word of the multiplier 1
    u64 r64; // 2^64 mod p
    u64 r64s; // Shoup word of r64
    // small-quotient reduction of a lazy word x < 128p:
    //   q = ((x >> fin_shift) * fin_m) >> 56,  x - q*p in [0,2p)
    unsigned fin_m; // floor(2^(bits+31) / p), fits 32 bits
    unsigned fin_shift; // bits - 25
    unsigned bits; // bit length of p
    unsigned nc_ok; // 1 if p <= 57 bits: butterflies may skip every per-stage correction
    unsigned fp_var; // 0: integer butterflies; 3 / 4: forward twiddles are doubles {w, RN(w/p)}, use VAR 3 / 4
    unsigned pad_;
    double pinv; // RN(1/p) (FP64 variants)
    double pinv_lo; // RN(1/p - pinv): pinv + pinv_lo = 1/p to ~106 bits
};

// Method-II (hybrid, K > 1) level tables; one entry per depth.
// Layouts follow src/lib/kernel/contextpool.cpp:193-438 of the reference.
struct LevelTablesII {
    int d = 0; // digit count at this depth
    std::vector<u64> base_change; // [digit][k over Q'_l][i in digit]
    std::vector<u64> mi_inv; // [I_location + i]
    std::vector<u64> prod; // [digit][k over Q'_l]
    std::vector<int> I_j; // digit sizes
    std::vector<int> I_loc; // prefix sums
    // device copies
    u64* d_base_change = nullptr;
    u64* d_mi_inv = nullptr;
    u64* d_prod = nullptr;
    TwPair* d_mi_inv_pair = nullptr; // mi_inv with Shoup words
    TwPair* d_base_change_pair = nullptr; // base_change with Shoup words (for the target prime)
    u64* d_rprod = nullptr; // [r = 0..K][digit][k]: r * prod mod t_k
    int* d_I_j = nullptr;
    int* d_I_loc = nullptr;
};

// BFV BEHZ multiplication tables (reference member names, bfv/context.cu:543-660)
struct BfvTables {
    std::vector<u64> base_change_matrix_Bsk, inv_punctured_prod_mod_base_array, base_change_matrix_m_tilde,
        inv_m_tilde_mod_Bsk, prod_q_mod_Bsk, inv_prod_q_mod_Bsk, base_change_matrix_q, base_change_matrix_msk,
        inv_punctured_prod_mod_B_array, prod_B_mod_q;
    u64 inv_prod_q_mod_m_tilde = 0, inv_prod_B_mod_m_sk = 0;
    // plaintext operands (bfv/context.cu:501-516, 936-984): Q mod t, floor(Q/t) mod q_i, (t+1)>>1, q_i - t
    u64 Q_mod_t = 0, upper_threshold = 0;
    std::vector<u64> coeff_div_plainmod, upper_halfincrement;
    u64 *d_coeff_div_plainmod = nullptr, *d_upper_halfincrement = nullptr;
    u64 *d_base_change_matrix_Bsk = nullptr, *d_inv_punctured_prod_mod_base_array = nullptr,
        *d_base_change_matrix_m_tilde = nullptr, *d_inv_m_tilde_mod_Bsk = nullptr, *d_prod_q_mod_Bsk = nullptr,
        *d_inv_prod_q_mod_Bsk = nullptr, *d_base_change_matrix_q = nullptr, *d_base_change_matrix_msk = nullptr,
        *d_inv_punctured_prod_mod_B_array = nullptr, *d_prod_B_mod_q = nullptr;
};

struct Context {
    int device = 0;
    int scheme = SCHEME_CKKS;
    int n = 0, logn = 0;
    int Q_size = 0, P_size = 0, Qp = 0;
    int method = 1; // key-switching method: 1 (K == 1) or 2 (K > 1)
    int ntt_variant = 2; // butterfly variant (ntt_core.cuh); HEON_NTT_VARIANT overrides
    std::vector<Mod64> mod; // [q_0..q_{Q-1}, p_0..p_{K-1}]
    std::vector<u64> psi; // minimal primitive 2N-th roots

    // Reference-layout host tables (kept for introspection and parity tests).
    std::vector<u64> ntt_table, intt_table, n_inverse;
    std::vector<u64> last_q_modinv, half, half_mod, factor;
    std::vector<u64> rescaled_last_q_modinv, rescaled_half_mod, rescaled_half;
    std::vector<LevelTablesII> lvl2; // Method II only

    // BFV only: plain modulus, size of the auxiliary base Bsk (its primes follow the Q' chain in `mod`)
    u64 plain_modulus = 0;
    int bsk = 0;
    BfvTables bfv;

    // Device tables
    Mod64* d_mod = nullptr; // [Qp]
    PrimeConst* d_pc = nullptr; // [Qp]
    TwPair* d_fwd = nullptr; // [Qp][N]  psi^bitrev(i) with Shoup word
    TwPair* d_inv = nullptr; // [Qp][N]  psi^-bitrev(i) with Shoup word
    TwPair* d_inv_last = nullptr; // [Qp][2] {n^-1, W_inv[1]*n^-1}
    // lane-major copies of the last-four-stage twiddles of the row pass:
    // [Qp][rows][16 entries][16 lanes] (ntt_core.cuh: ct_round_b_lm)
    TwPair* d_fwd_rowb = nullptr;
    TwPair* d_inv_rowb = nullptr;
    // compact FP64 row-pass twiddles, bare doubles: [Qp][rows][256] (ntt_core.cuh: stage16_sm);
    // a 16-row tile's twiddles are 32 KiB contiguous and travel by TMA next to the data tile
    double* d_fwd_rowc = nullptr;
    int use_tma = 1; // row pass through TMA tensor maps (HEON_NTT_TMA=0 disables)
    int num_sms = 148;
    int ntt_persistent = 0; // HEON_NTT_PERSISTENT=1: row-pass CTAs walk several tiles (double-buffered TMA)
    int ntt_pipe = 0; // N = 2^16: warp-specialised pipelined fused forward transform (HEON_NTT_PIPE=1 enables; needs all CTAs co-resident, i.e. an otherwise idle GPU)
    int ntt_group = 48; // polynomials per L2-resident group of the fused transform (HEON_NTT_GROUP)
    int col_threads = 0; // threads per CTA of the column pass at N = 2^16: 256, 128, or 0 = per map (HEON_COL_THREADS)
    int row_tile = 8; // rows per CTA of the TMA row pass: 16 (32 KiB tiles, 3 CTAs/SM), 8 (6 CTAs/SM, default: +3..7 % measured) or 4 (HEON_ROW_TILE)
    int skip_own = 1; // Method II: the digits' own limbs skip mod-up and forward NTT (HEON_SKIP_OWN=0 disables)
    int galois_ntt = 1; // CKKS automorphisms as NTT-domain permutations after an NTT-domain key switch (HEON_GALOIS_NTT=0: coefficient-domain path)
    int ntt_fused = 0; // forward transform as one ticket-ordered kernel, pass-to-pass data in L2 (HEON_NTT_FUSED=1 enables; superseded by the pipelined kernel)
    int use_fp64 = 1; // FP64-pipe quotient for primes < 2^50 (HEON_NTT_FP64=0 disables)
    int col_tma = 1;             // HEON_COL_TMA: forward column pass at N = 2^16 through pipelined TMA tiles (0: register-resident LSU form)
    int modup_doubles = 1;       // HEON_MODUP_DOUBLES: the fast Method-II mod-up leaves FP64-prime words as doubles for the column pass
    int row_walk = -1;           // HEON_ROW_WALK: forward row pass walks this many same-prime polynomials per CTA (-1 = 8, 0 = one tile per CTA)
    int col_tma_bufs = 2;        // HEON_COL_TMA_BUFS: tile buffers per CTA of the pipelined TMA column pass (2: 3 CTAs/SM, 3: 2 CTAs/SM)
    int col_tma_tiles = 0;       // HEON_COL_TMA_TILES: tiles one CTA of the TMA column pass walks (0 = by grid size)
    int row_mac = 1; // key switch: forward row pass fused with the inner product (HEON_ROW_MAC=0: separate kernels)
    int modup_fused = 0; // HEON_MODUP_FUSED=1: Method-II mod-up computed inside the column-pass load (no converted-digit buffer; measured slower than the separate FP64 kernel on B200, kept opt-in)
    int row_final = 1; // Method-II mod-down: forward row pass of the corrections fused with the final combination (HEON_ROW_FINAL=0: separate kernels; > 1: polynomials one CTA walks)
    int modup_col = 0; // HEON_MODUP_COL=1: Method-II mod-up fused with the column pass, source tiles staged once per CTA (N = 2^16; 2: also for tiny grids).  Bit-exact, measured slower than the separate kernels on B200 (153 against 117 us/op at C3-II: one 185 KB CTA per SM keeps the FP64 pipe 42 % busy), kept opt-in
    int modup_cw = 2; // HEON_MODUP_CW: coefficients per thread of the fast Method-II mod-up (2; 4 measured slower: 54.8 against 51.2 us/op at C3-II, 119 registers halve the resident warps)
    int row_mac_rows = 4; // rows per CTA of the fused kernel: 4 (default, 6 CTAs/SM: +2 % measured) or 8 (HEON_ROW_MAC_ROWS)
    u64* d_last_q_modinv = nullptr;
    TwPair* d_lqm_pair = nullptr; // last_q_modinv with Shoup words
    // FP64 form of the correction chain for Q primes below 2^50 (k_moddown2_corr):
    //   c_y = Cst_y - sum_i (hi_i * (2^30 B_{i,y}) + lo_i * B_{i,y}),  B_{i,y} = prod_{j>=i} m_{j,y}
    // [Q][K][2] double pairs {w, RN(w/q_y)} for {2^30*B mod q, B}, and Cst_y = sum_i half_mod_i * B_i mod q_y
    TwPair* d_md2_B = nullptr;
    u64* d_md2_cst = nullptr;
    TwPair* d_md2_M = nullptr; // [Q]: prod_i last_q_modinv[block i][y] = (p_0..p_{K-1})^-1 mod q_y, Shoup pair
    u64* d_half = nullptr;
    u64* d_half_mod = nullptr;
    u64* d_rescaled_last_q_modinv = nullptr;
    u64* d_rescaled_half_mod = nullptr;
    u64* d_rescaled_half = nullptr;

    std::string last_error;

    ~Context();
};

// Prime index used by slot y of a levelled limb set {q_0..q_{L-1}, p_0..p_{K-1}}:
// y < L -> y, else y + depth   (reference: ckks/operator.cu:24-39).
__host__ __device__ inline int level_prime(int y, int L, int depth) { return y < L ? y : y + depth; }

// Small by-value list of prime indices (one per limb slot of a limb set).
struct PrimeList {
    int count;
    unsigned char idx[128];
};

inline PrimeList level_primes(int L, int K, int depth)
{
    PrimeList pl;
    pl.count = L + K;
    for (int y = 0; y < L + K; ++y)
        pl.idx[y] = (unsigned char) level_prime(y, L, depth);
    return pl;
}
inline PrimeList range_primes(int first, int count)
{
    PrimeList pl;
    pl.count = count;
    for (int y = 0; y < count; ++y)
        pl.idx[y] = (unsigned char) (first + y);
    return pl;
}

void build_host_tables(Context& c);
void upload_tables(Context& c);
void build_bfv_tables(Context& c);
void upload_bfv_tables(Context& c);

// ---- NTT launchers (ntt.cu); all asynchronous on `st` ----
// col_only: run only the first n-8 (column) stages; the row stages follow inside launch_row_mac
void launch_ntt(const Context& c, const u64* src, u64* dst, long long n_polys, const PrimeList& pl,
                bool inverse, cudaStream_t st, bool col_only = false);
void launch_ntt_digit_skip(const Context& c, u64* tmp, int d, const int* I_loc, const int* I_j, int L, int depth,
                           long long batch, cudaStream_t st, bool col_only = false, unsigned long long dbl_mask = 0);
bool modup2_fused_available(const Context& c, int depth, const u64* coef, long long coef_bs);
void launch_modup2_ntt(const Context& c, const u64* coef, long long coef_bs, u64* tmp, u64* part, unsigned char* rq,
                       int depth, long long batch, bool own_stashed, bool col_only, cudaStream_t st);
bool row_mac_available(const Context& c, const u64* tmp, const u64* key, const u64* acc, int d);
void launch_row_mac(const Context& c, const u64* tmp, const u64* key, u64* acc, int d, int depth, int batch,
                    bool own_stashed, const int* I_loc, const int* I_j, cudaStream_t st);
bool modup2_col_available(const Context& c, int depth, const u64* coef, long long coef_bs, const u64* tmp, int batch,
                          bool own_stashed, bool col_only);
void launch_modup2_col(const Context& c, const u64* coef, long long coef_bs, u64* tmp, int depth, long long batch,
                       cudaStream_t st);
bool row_final_available(const Context& c, const u64* tmp, const u64* acc, const u64* ct_in, long long ct_bs,
                         const u64* out, long long out_bs, int L, int batch);
void launch_row_final(const Context& c, const u64* tmp, const u64* acc, const u64* ct_in, long long ct_bs, u64* out,
                      long long out_bs, int depth, int batch, int add_mask, cudaStream_t st);
void launch_ntt_scattered(const Context& c, u64* base, const long long* d_offsets, int n_polys,
                          int prime, bool inverse, long long extent_words, bool aligned, cudaStream_t st);
void launch_ntt_strided(const Context& c, u64* base, long long bstride, int per_batch, int first,
                        long long batch, const PrimeList& pl, bool inverse, cudaStream_t st);
void launch_ntt_strided_copy(const Context& c, const u64* src, long long src_bstride, u64* dst,
                             int per_batch, long long batch, const PrimeList& pl, bool inverse,
                             cudaStream_t st);
void launch_modup1_ntt(const Context& c, const u64* coef, long long coef_bstride, u64* out, int L,
                       int depth, long long batch, cudaStream_t st, bool col_only = false);
void launch_divround1_ntt(const Context& c, const u64* src, long long bstride, long long cstride,
                          u64* out, int Lout, u64 half, u64 plast, const u64* d_half_mod,
                          long long batch, cudaStream_t st, bool col_only = false);

} // namespace heon
