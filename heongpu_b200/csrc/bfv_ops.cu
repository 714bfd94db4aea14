// BFV hot path: BEHZ ciphertext multiplication (fast base conversion q -> Bsk U
// {m_tilde}, tensor product over q U Bsk, fast floor back to q) and the
// un-levelled key switch behind relinearize.
//
// Replaces src/lib/host/bfv/operator.cu:336-430 (multiply_bfv), :505-671
// (relinearize_seal_method_inplace / relinearize_external_product_method2_inplace)
// and the kernels fast_convertion / fast_floor (src/lib/kernel/multiplication.cu:
// 10-100, 128-272), cipher_broadcast_kernel, keyswitch_multiply_accumulate_kernel,
// divide_round_lastq_kernel / _extended_kernel, base_conversion_DtoQtilde_relin_kernel
// (src/lib/kernel/switchkey.cu:11-27, 61-162, 400-437, 480-543, 872-927).
// Every step is exact modular arithmetic on canonical residues, so stored
// words equal the reference's Barrett results.
#include "modarith.cuh"
#include "ops.hpp"

namespace heon {

constexpr int kMaxBase = 64; // MAX_BSK_SIZE (src/include/heongpu/kernel/defines.h)

__device__ __forceinline__ u64 mulmod_pc(u64 a, u64 b, const PrimeConst& pc)
{
    return reduce_u128(a * b, __umul64hi(a, b), pc);
}

// q -> Bsk with the Montgomery-style m_tilde correction (SmMRq).
// in: two ciphertexts of two polys each ([2][Q][N], coefficient domain);
// out[b][poly 0..3][Q + m][N]: the q limbs copied, then the m Bsk limbs.
__global__ void __launch_bounds__(256)
    k_bfv_fast_convertion(const u64* __restrict__ in1, long long in1_bs, const u64* __restrict__ in2,
                          long long in2_bs, u64* __restrict__ out, const PrimeConst* __restrict__ pcs,
                          const u64* __restrict__ inv_punct, const u64* __restrict__ bcm_bsk,
                          const u64* __restrict__ bcm_mt, unsigned inv_prod_q_mod_mt,
                          const u64* __restrict__ inv_mt_mod_bsk, const u64* __restrict__ prod_q_mod_bsk,
                          int logn, int Q, int Qp, int m)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int idy = blockIdx.y; // 0..3
    const long long bz = blockIdx.z;
    const u64* input = (idy >> 1) == 0 ? in1 + bz * in1_bs : in2 + bz * in2_bs;
    input += idx + ((long long) ((idy & 1) * Q) << logn);
    u64* po = out + (((bz * 4 + idy) * (Q + m)) << logn) + idx;

    u64 temp[kMaxBase];
    unsigned acc_mt = 0; // arithmetic modulo m_tilde = 2^32 is plain 32-bit wrap-around
    for (int i = 0; i < Q; ++i)
    {
        const PrimeConst pc = pcs[i];
        const u64 x = input[(long long) i << logn];
        po[(long long) i << logn] = x;
        u64 t = mulmod_pc(x, 1ull << 32, pc);
        t = mulmod_pc(t, inv_punct[i], pc);
        temp[i] = t;
        acc_mt += (unsigned) t * (unsigned) bcm_mt[i];
    }
    // r_m_tilde = m_tilde - (acc * inv_prod_q mod m_tilde), a value in [1, 2^32]
    const u64 r_mt = (1ull << 32) - (u64) (unsigned) (acc_mt * inv_prod_q_mod_mt);
    for (int k = 0; k < m; ++k)
    {
        const PrimeConst pb = pcs[Qp + k];
        u64 lo = 0, hi = 0;
        for (int j = 0; j < Q; ++j)
            mac128(lo, hi, temp[j], bcm_bsk[j + k * Q]);
        const u64 s = reduce_u128(lo, hi, pb);
        u64 t3 = r_mt;
        if (t3 >= (1ull << 31))
            t3 = mod_add(pb.p - (1ull << 32), r_mt, pb.p);
        t3 = mulmod_pc(t3, prod_q_mod_bsk[k], pb);
        t3 = mod_add(s, t3, pb.p);
        po[(long long) (Q + k) << logn] = mulmod_pc(t3, inv_mt_mod_bsk[k], pb);
    }
}

// tensor product over the merged base q U Bsk (cross_multiplication with a prime list)
__global__ void __launch_bounds__(256)
    k_bfv_cross_multiply(const u64* __restrict__ t1, u64* __restrict__ t2, const PrimeConst* __restrict__ pcs,
                         PrimeList pl, int logn)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z;
    const int W = pl.count;
    const PrimeConst pc = pcs[pl.idx[y]];
    const u64* a = t1 + ((bz * 4 * W + y) << logn) + idx;
    const long long comp = (long long) W << logn;
    const u64 a0 = a[0], a1 = a[comp], b0 = a[2 * comp], b1 = a[3 * comp];
    u64* o = t2 + ((bz * 3 * W + y) << logn) + idx;
    o[0] = mulmod_pc(a0, b0, pc);
    o[comp] = mod_add(mulmod_pc(a0, b1, pc), mulmod_pc(a1, b0, pc), pc.p);
    o[2 * comp] = mulmod_pc(a1, b1, pc);
}

// floor(t/q * x): q U Bsk -> Bsk (exact division by q) -> Shenoy-Kumaresan back to q.
__global__ void __launch_bounds__(256)
    k_bfv_fast_floor(const u64* __restrict__ in, u64* __restrict__ out, long long out_bs,
                     const PrimeConst* __restrict__ pcs, u64 plain_modulus, const u64* __restrict__ inv_punct,
                     const u64* __restrict__ bcm_bsk, const u64* __restrict__ inv_prod_q_mod_bsk,
                     const u64* __restrict__ inv_punct_B, const u64* __restrict__ bcm_q,
                     const u64* __restrict__ bcm_msk, u64 inv_prod_B_mod_msk, const u64* __restrict__ prod_B_mod_q,
                     int logn, int Q, int Qp, int m)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int idy = blockIdx.y; // 0..2
    const long long bz = blockIdx.z;
    const u64* pq = in + (((bz * 3 + idy) * (Q + m)) << logn) + idx;
    const u64* pB = pq + ((long long) Q << logn);
    u64* po = out + bz * out_bs + ((long long) (idy * Q) << logn) + idx;

    u64 reg_q[kMaxBase], reg_B[kMaxBase], temp3[kMaxBase];
    for (int i = 0; i < Q; ++i)
    {
        const PrimeConst pc = pcs[i];
        u64 t = mulmod_pc(pq[(long long) i << logn], plain_modulus, pc);
        reg_q[i] = mulmod_pc(t, inv_punct[i], pc);
    }
    for (int k = 0; k < m; ++k)
    {
        const PrimeConst pb = pcs[Qp + k];
        const u64 rB = mulmod_pc(pB[(long long) k << logn], plain_modulus, pb);
        u64 lo = 0, hi = 0;
        for (int j = 0; j < Q; ++j)
            mac128(lo, hi, reg_q[j], bcm_bsk[j + k * Q]);
        const u64 t = reduce_u128(lo, hi, pb);
        // (reg_Bsk - t) * (prod q)^-1   [the reference computes sub(p, t) which is p - t, or p for t = 0;
        //  the following add folds both cases to the same canonical value]
        const u64 d = mod_sub(rB, t, pb.p);
        reg_B[k] = mulmod_pc(d, inv_prod_q_mod_bsk[k], pb);
    }
    for (int k = 0; k < m - 1; ++k)
        temp3[k] = mulmod_pc(reg_B[k], inv_punct_B[k], pcs[Qp + k]);

    const PrimeConst psk = pcs[Qp + m - 1];
    u64 lo = 0, hi = 0;
    for (int j = 0; j < m - 1; ++j)
        mac128(lo, hi, temp3[j], bcm_msk[j]);
    u64 alpha = mod_sub(reduce_u128(lo, hi, psk), reg_B[m - 1], psk.p);
    alpha = mulmod_pc(alpha, inv_prod_B_mod_msk, psk);
    const bool neg = alpha > (psk.p >> 1);

    for (int i = 0; i < Q; ++i)
    {
        const PrimeConst pc = pcs[i];
        u64 l2 = 0, h2 = 0;
        for (int j = 0; j < m - 1; ++j)
            mac128(l2, h2, reduce_u64(temp3[j], pc), bcm_q[j + i * (m - 1)]);
        const u64 t4 = reduce_u128(l2, h2, pc);
        const u64 msk_q = reduce_u64(psk.p, pc);
        const u64 alpha_q = reduce_u64(alpha, pc);
        u64 inner;
        if (neg)
            inner = mulmod_pc(mod_sub(msk_q, alpha_q, pc.p), prod_B_mod_q[i], pc);
        else
            inner = mulmod_pc(mod_sub(0, prod_B_mod_q[i], pc.p), alpha_q, pc);
        po[(long long) i << logn] = mod_add(t4, inner, pc.p);
    }
}

static void check_launch_bfv()
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("kernel launch: ") + cudaGetErrorString(e));
}

struct ScratchB {
    void* p = nullptr;
    cudaStream_t st;
    ScratchB(size_t bytes, cudaStream_t s) : st(s)
    {
        if (cudaMallocAsync(&p, bytes, s) != cudaSuccess)
            throw std::runtime_error("cudaMallocAsync failed");
    }
    ~ScratchB()
    {
        if (p)
            cudaFreeAsync(p, st);
    }
    u64* w() const { return (u64*) p; }
};

// a, b: [2][Q][N] coefficient domain; out: [3][Q][N] coefficient domain.
void op_bfv_multiply(const Context& c, const u64* a, long long a_bs, const u64* b, long long b_bs, u64* out,
                     long long o_bs, int batch, cudaStream_t st)
{
    if (c.scheme != SCHEME_BFV)
        throw std::invalid_argument("not a BFV context");
    const int Q = c.Q_size, m = c.bsk, W = Q + m;
    const long long N = c.n;
    if (W > kMaxBase)
        throw std::invalid_argument("too many RNS primes for the BEHZ kernels");
    const BfvTables& t = c.bfv;
    ScratchB t1((size_t) batch * 4 * W * N * 8, st), t2((size_t) batch * 3 * W * N * 8, st);
    PrimeList pl;
    pl.count = W;
    for (int i = 0; i < Q; ++i)
        pl.idx[i] = (unsigned char) i;
    for (int k = 0; k < m; ++k)
        pl.idx[Q + k] = (unsigned char) (c.Qp + k);
    {
        LaunchScope scope(KC_MODUP2, st);
        k_bfv_fast_convertion<<<dim3(c.n >> 8, 4, batch), 256, 0, st>>>(
            a, a_bs, b, b_bs, t1.w(), c.d_pc, t.d_inv_punctured_prod_mod_base_array, t.d_base_change_matrix_Bsk,
            t.d_base_change_matrix_m_tilde, (unsigned) t.inv_prod_q_mod_m_tilde, t.d_inv_m_tilde_mod_Bsk,
            t.d_prod_q_mod_Bsk, c.logn, Q, c.Qp, m);
    }
    check_launch_bfv();
    launch_ntt(c, t1.w(), t1.w(), (long long) batch * 4 * W, pl, false, st);
    {
        LaunchScope scope(KC_CROSS_MULTIPLY, st);
        k_bfv_cross_multiply<<<dim3(c.n >> 8, W, batch), 256, 0, st>>>(t1.w(), t2.w(), c.d_pc, pl, c.logn);
    }
    check_launch_bfv();
    launch_ntt(c, t2.w(), t2.w(), (long long) batch * 3 * W, pl, true, st);
    {
        LaunchScope scope(KC_MODDOWN, st);
        k_bfv_fast_floor<<<dim3(c.n >> 8, 3, batch), 256, 0, st>>>(
            t2.w(), out, o_bs, c.d_pc, c.plain_modulus, t.d_inv_punctured_prod_mod_base_array,
            t.d_base_change_matrix_Bsk, t.d_inv_prod_q_mod_Bsk, t.d_inv_punctured_prod_mod_B_array,
            t.d_base_change_matrix_q, t.d_base_change_matrix_msk, t.inv_prod_B_mod_m_sk, t.d_prod_B_mod_q, c.logn,
            Q, c.Qp, m);
    }
    check_launch_bfv();
}

// ---------------------------------------------------------------------------
// BFV plaintext operands (coefficient-domain ciphertexts, plaintext = [N] values below t)
// ---------------------------------------------------------------------------

// component 0 +/- (m * floor(Q/t) + round-fix), the other components copied.
// reference: src/lib/kernel/addition.cu:50-173 (addition_plain_bfv_poly, substraction_plain_bfv_poly)
template <int OP>
__global__ void __launch_bounds__(256)
    k_bfv_addsub_plain(const u64* __restrict__ ct, long long ct_bs, const u64* __restrict__ pt, long long pt_bs,
                       u64* __restrict__ out, long long o_bs, const Mod64* __restrict__ mods, u64 plain_mod,
                       u64 Q_mod_t, u64 upper_threshold, const u64* __restrict__ coeff_div, int logn, int Q, int comps)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z / comps;
    const int c = blockIdx.z % comps;
    const long long loc = idx + ((long long) (c * Q + y) << logn);
    u64 x = ct[bz * ct_bs + loc];
    if (c == 0)
    {
        const Mod64 m = mods[y];
        const u64 message = pt[bz * pt_bs + idx];
        u64 fix = message * Q_mod_t + upper_threshold;
        fix = (u64) (long long) (int) (fix / plain_mod); // `int(fix / plain_mod.value)` in the reference kernel
        u64 r = barrett_mul(message, coeff_div[y], m);
        r = mod_add(r, fix, m.value);
        x = OP == 0 ? mod_add(r, x, m.value) : mod_sub(x, r, m.value);
    }
    out[bz * o_bs + loc] = x;
}

// centred lift of the plaintext into every q_i: m >= (t+1)/2 -> m + (q_i - t)
// reference: src/lib/kernel/multiplication.cu:274-296 (threshold_kernel)
__global__ void __launch_bounds__(256)
    k_bfv_threshold(const u64* __restrict__ pt, long long pt_bs, u64* __restrict__ out, const Mod64* __restrict__ mods,
                    const u64* __restrict__ upper_inc, u64 upper_threshold, int logn, int Q)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z;
    const u64 v = pt[bz * pt_bs + idx];
    out[((bz * Q + y) << logn) + idx] = v >= upper_threshold ? mod_add(v, upper_inc[y], mods[y].value) : v;
}

// reference: src/lib/kernel/multiplication.cu:298-311 (cipherplain_kernel), in place on the NTT-domain copy
__global__ void __launch_bounds__(256)
    k_bfv_cipherplain(u64* __restrict__ ct, const u64* __restrict__ pt, const Mod64* __restrict__ mods, int logn, int Q)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long bz = blockIdx.z >> 1;
    const int c = blockIdx.z & 1;
    const long long o = (((bz * 2 + c) * Q + y) << logn) + idx;
    ct[o] = barrett_mul(ct[o], pt[((bz * Q + y) << logn) + idx], mods[y]);
}

// add_plain_bfv / sub_plain_bfv (bfv/operator.cu:216-340): op 1 add, 2 subtract
void op_bfv_addsub_plain(const Context& c, const u64* ct, long long ct_bs, const u64* pt, long long pt_bs, u64* out,
                         long long o_bs, int comps, int batch, int op, cudaStream_t st)
{
    if (c.scheme != SCHEME_BFV)
        throw std::invalid_argument("not a BFV context");
    if (comps < 2 || comps > 3)
        throw std::invalid_argument("Invalid Ciphertexts size!");
    const BfvTables& t = c.bfv;
    dim3 g(c.n >> 8, c.Q_size, batch * comps);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        if (op == 1)
            k_bfv_addsub_plain<0><<<g, 256, 0, st>>>(ct, ct_bs, pt, pt_bs, out, o_bs, c.d_mod, c.plain_modulus, t.Q_mod_t,
                                                   t.upper_threshold, t.d_coeff_div_plainmod, c.logn, c.Q_size, comps);
        else
            k_bfv_addsub_plain<1><<<g, 256, 0, st>>>(ct, ct_bs, pt, pt_bs, out, o_bs, c.d_mod, c.plain_modulus, t.Q_mod_t,
                                                   t.upper_threshold, t.d_coeff_div_plainmod, c.logn, c.Q_size, comps);
    }
    check_launch_bfv();
}

// multiply_plain_bfv (bfv/operator.cu:432-503), coefficient-domain ciphertext: lift the plaintext,
// NTT both, multiply, INTT.  out: [b][2][Q][N] contiguous.
void op_bfv_multiply_plain(const Context& c, const u64* ct, long long ct_bs, const u64* pt, long long pt_bs, u64* out,
                           long long o_bs, int batch, cudaStream_t st)
{
    if (c.scheme != SCHEME_BFV)
        throw std::invalid_argument("not a BFV context");
    const int Q = c.Q_size;
    const long long N = c.n;
    if (o_bs != 2 * Q * N && batch > 1)
        throw std::invalid_argument("multiply_plain needs a contiguous output batch");
    const BfvTables& t = c.bfv;
    ScratchB tp((size_t) batch * Q * N * 8, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_bfv_threshold<<<dim3(c.n >> 8, Q, batch), 256, 0, st>>>(pt, pt_bs, tp.w(), c.d_mod, t.d_upper_halfincrement,
                                                               t.upper_threshold, c.logn, Q);
    }
    check_launch_bfv();
    launch_ntt(c, tp.w(), tp.w(), (long long) batch * Q, range_primes(0, Q), false, st);
    launch_ntt_strided_copy(c, ct, ct_bs, out, 2 * Q, batch, range_primes(0, Q), false, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_bfv_cipherplain<<<dim3(c.n >> 8, Q, batch * 2), 256, 0, st>>>(out, tp.w(), c.d_mod, c.logn, Q);
    }
    check_launch_bfv();
    launch_ntt(c, out, out, (long long) batch * 2 * Q, range_primes(0, Q), true, st);
}

} // namespace heon
