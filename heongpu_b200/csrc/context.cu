// Context construction: prime chain, roots, NTT tables and every key-switch /
// rescale constant the kernels consume.
//
// Table contents follow the reference bit for bit (all entries are exact
// canonical residues, which is what the reference's host Barrett code yields):
//   primes / psi / ntt tables     src/lib/util/util.cu:219-276,356-464
//   last_q_modinv, half, half_mod src/lib/util/util.cu:700-767
//   rescale tables                src/lib/host/ckks/context.cu:342-368
//   Method-II level tables        src/lib/kernel/contextpool.cpp:11-66,193-438
#include <cstdlib>
#include <cstring>
#include "heon_internal.hpp"
#include "modarith.cuh"

namespace heon {

#define HEON_CUDA(x)                                                                               \
    do                                                                                             \
    {                                                                                              \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess)                                                                     \
            throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_));              \
    } while (0)

template <class T> static T* upload(const std::vector<T>& h)
{
    if (h.empty())
        return nullptr;
    T* d = nullptr;
    HEON_CUDA(cudaMalloc(&d, h.size() * sizeof(T)));
    HEON_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

static std::vector<int> digit_sizes(int l, int m)
{
    std::vector<int> r;
    while (l > 0)
    {
        r.push_back(l > m ? m : l);
        l -= m;
    }
    return r;
}

void build_host_tables(Context& c)
{
    const int N = c.n, Qp = c.Qp, Q = c.Q_size, K = c.P_size;
    const int T = (int) c.mod.size(); // Q' chain followed by the BFV auxiliary base Bsk (if any)
    c.psi.resize(T);
    c.ntt_table.assign((size_t) T * N, 0);
    c.intt_table.assign((size_t) T * N, 0);
    c.n_inverse.resize(T);
    std::vector<u64> pw(N);
    for (int i = 0; i < T; ++i)
    {
        const u64 p = c.mod[i].value;
        c.psi[i] = minimal_primitive_root(2 * (u64) N, p);
        for (int pass = 0; pass < 2; ++pass)
        {
            const u64 root = pass ? invmod(c.psi[i], p) : c.psi[i];
            pw[0] = 1;
            for (int j = 1; j < N; ++j)
                pw[j] = mulmod(pw[j - 1], root, p);
            u64* dst = (pass ? c.intt_table.data() : c.ntt_table.data()) + (size_t) i * N;
            for (int j = 0; j < N; ++j)
                dst[j] = pw[bitrev(j, c.logn)];
        }
        c.n_inverse[i] = invmod(N, p);
    }

    // mod-down constants, one block per dropped P prime (last P first)
    c.last_q_modinv.clear();
    c.half.clear();
    c.half_mod.clear();
    c.factor.clear();
    for (int i = 0; i < K; ++i)
    {
        const u64 last = c.mod[Qp - 1 - i].value;
        c.half.push_back(last >> 1);
        for (int j = 0; j < Qp - 1 - i; ++j)
        {
            const u64 pj = c.mod[j].value;
            c.last_q_modinv.push_back(invmod(last % pj, pj));
            c.half_mod.push_back((last >> 1) % pj);
        }
        for (int j = 0; j < Q; ++j)
            c.factor.push_back(last % c.mod[j].value);
    }

    // rescale constants for every depth
    c.rescaled_half.clear();
    c.rescaled_half_mod.clear();
    c.rescaled_last_q_modinv.clear();
    for (int j = 0; j < Q - 1; ++j)
    {
        const int inner = Q - 1 - j;
        const u64 ql = c.mod[inner].value;
        c.rescaled_half.push_back(ql >> 1);
        for (int i = 0; i < inner; ++i)
        {
            const u64 pi = c.mod[i].value;
            c.rescaled_last_q_modinv.push_back(invmod(ql % pi, pi));
            c.rescaled_half_mod.push_back((ql >> 1) % pi);
        }
    }

    // Method II (hybrid key switching with K > 1): per-depth digit tables
    c.lvl2.clear();
    if (c.method == 2)
    {
        // BFV has no levels: only the depth-0 tables (contextpool.cpp:160-191, 242-264, 361-394)
        for (int depth = 0; depth < (c.scheme == SCHEME_CKKS ? Q : 1); ++depth)
        {
            const int L = Q - depth;
            // limb set at this depth: q_0..q_{L-1}, p_0..p_{K-1}
            std::vector<u64> base;
            for (int y = 0; y < L + K; ++y)
                base.push_back(c.mod[level_prime(y, L, depth)].value);
            LevelTablesII t;
            // digit size: |P| for CKKS (contextpool.cpp:104); the reference's BFV branch never sets it and
            // keeps the member default m = 2 (contextpool.hpp:29, contextpool.cpp:78-92)
            t.I_j = digit_sizes(L, c.scheme == SCHEME_CKKS ? K : 2);
            t.d = (int) t.I_j.size();
            t.I_loc.assign(t.d, 0);
            for (int l = 1; l < t.d; ++l)
                t.I_loc[l] = t.I_loc[l - 1] + t.I_j[l - 1];
            // Generators follow contextpool.cpp:160-191 / 193-236 (base change), 242-264 / 266-308 (Mi_inv),
            // 361-394 / 396-438 (prod) with the reference's host Barrett product and its UNREDUCED prime
            // operand: the words equal the reference's even where that leaves a non-canonical representative.
            std::vector<Mod64> bm;
            for (u64 b : base)
                bm.push_back(make_mod(b));
            for (int l = 0; l < t.d; ++l)
            {
                const int lo = t.I_loc[l], sz = t.I_j[l];
                for (int k = 0; k < L + K; ++k)
                    for (int i = 0; i < sz; ++i)
                    {
                        u64 prod = 1;
                        for (int j = 0; j < sz; ++j)
                            if (j != i)
                                prod = barrett_mult_host(prod, base[lo + j], bm[k]);
                        t.base_change.push_back(prod);
                    }
            }
            for (int l = 0; l < t.d; ++l)
            {
                const int lo = t.I_loc[l], sz = t.I_j[l];
                for (int i = 0; i < sz; ++i)
                {
                    u64 prod = 1;
                    for (int j = 0; j < sz; ++j)
                        if (j != i)
                            prod = barrett_mult_host(prod, base[lo + j], bm[lo + i]);
                    t.mi_inv.push_back(barrett_modinv_host(prod, bm[lo + i]));
                }
            }
            for (int l = 0; l < t.d; ++l)
            {
                const int lo = t.I_loc[l], sz = t.I_j[l];
                for (int k = 0; k < L + K; ++k)
                {
                    u64 prod = 1;
                    for (int j = 0; j < sz; ++j)
                        prod = barrett_mult_host(prod, base[lo + j], bm[k]);
                    t.prod.push_back(prod);
                }
            }
            c.lvl2.push_back(std::move(t));
        }
    }
}

// BFV (BEHZ) multiplication constants.  Every entry is the exact residue the
// reference's host code produces (src/lib/host/bfv/context.cu:990-1290).
// Prime chain layout: mod = [q_0..q_{Q-1}, p_0..p_{K-1}, B_0..B_{m-1}], the
// auxiliary base Bsk = {B_0..B_{m-2}} U {m_sk = B_{m-1}}; m_tilde = 2^32.
static u64 inv_mod_pow2_32(u64 a)
{
    // a odd; Newton iteration modulo 2^32
    u64 x = a;
    for (int i = 0; i < 6; ++i)
        x = (x * (2 - a * x)) & 0xffffffffull;
    return x & 0xffffffffull;
}

// plaintext-operand constants (bfv/context.cu:501-516, 936-984); floor(Q/t) mod q_i = -(Q mod t) * t^-1 mod q_i
// because Q - (Q mod t) is divisible by t and Q = 0 (mod q_i)  (the reference divides the big integer with GMP)
static void build_bfv_plain_constants(Context& c)
{
    BfvTables& t = c.bfv;
    const u64 tt = c.plain_modulus;
    u64 r = 1;
    for (int i = 0; i < c.Q_size; ++i)
        r = mulmod(r, c.mod[i].value % tt, tt);
    t.Q_mod_t = r;
    t.upper_threshold = (tt + 1) >> 1;
    t.coeff_div_plainmod.clear();
    t.upper_halfincrement.clear();
    for (int i = 0; i < c.Q_size; ++i)
    {
        const u64 q = c.mod[i].value;
        const u64 rq = r % q;
        t.coeff_div_plainmod.push_back(rq == 0 ? 0 : mulmod(q - rq, invmod(tt % q, q), q));
        t.upper_halfincrement.push_back(q - tt);
    }
}

void build_bfv_tables(Context& c)
{
    const int Q = c.Q_size, m = c.bsk;
    const u64 mt = 1ull << 32;
    auto q = [&](int i) { return c.mod[i].value; };
    auto B = [&](int i) { return c.mod[c.Qp + i].value; };
    BfvTables& t = c.bfv;
    t = BfvTables();
    for (int k = 0; k < m; ++k) // generate_base_matrix_q_Bsk
        for (int i = 0; i < Q; ++i)
        {
            u64 v = 1;
            for (int j = 0; j < Q; ++j)
                if (j != i)
                    v = mulmod(v, q(j) % B(k), B(k));
            t.base_change_matrix_Bsk.push_back(v);
        }
    for (int i = 0; i < Q; ++i) // calculate_Mi_inv(prime_vector_, Q_size)
    {
        u64 v = 1;
        for (int j = 0; j < Q; ++j)
            if (j != i)
                v = mulmod(v, q(j) % q(i), q(i));
        t.inv_punctured_prod_mod_base_array.push_back(invmod(v, q(i)));
    }
    u64 prod_mt = 1;
    for (int i = 0; i < Q; ++i) // generate_base_change_matrix_m_tilde / inv_prod_q_mod_m_tilde
    {
        u64 v = 1;
        for (int j = 0; j < Q; ++j)
            if (j != i)
                v = (v * (q(j) % mt)) % mt;
        t.base_change_matrix_m_tilde.push_back(v);
        prod_mt = (prod_mt * (q(i) % mt)) % mt;
    }
    t.inv_prod_q_mod_m_tilde = inv_mod_pow2_32(prod_mt);
    for (int i = 0; i < m; ++i)
    {
        t.inv_m_tilde_mod_Bsk.push_back(invmod(mt % B(i), B(i)));
        u64 v = 1;
        for (int j = 0; j < Q; ++j)
            v = mulmod(v, q(j) % B(i), B(i));
        t.prod_q_mod_Bsk.push_back(v);
        t.inv_prod_q_mod_Bsk.push_back(invmod(v, B(i)));
    }
    for (int k = 0; k < Q; ++k) // generate_base_matrix_Bsk_q
        for (int i = 0; i < m - 1; ++i)
        {
            u64 v = 1;
            for (int j = 0; j < m - 1; ++j)
                if (j != i)
                    v = mulmod(v, B(j) % q(k), q(k));
            t.base_change_matrix_q.push_back(v);
        }
    const u64 msk = B(m - 1);
    u64 prodB_msk = 1;
    for (int i = 0; i < m - 1; ++i)
    {
        u64 v = 1, w = 1;
        for (int j = 0; j < m - 1; ++j)
            if (j != i)
            {
                v = mulmod(v, B(j) % msk, msk);
                w = mulmod(w, B(j) % B(i), B(i));
            }
        t.base_change_matrix_msk.push_back(v);
        t.inv_punctured_prod_mod_B_array.push_back(invmod(w, B(i)));
        prodB_msk = mulmod(prodB_msk, B(i) % msk, msk);
    }
    t.inv_prod_B_mod_m_sk = invmod(prodB_msk, msk);
    for (int i = 0; i < Q; ++i)
    {
        u64 v = 1;
        for (int j = 0; j < m - 1; ++j)
            v = mulmod(v, B(j) % q(i), q(i));
        t.prod_B_mod_q.push_back(v);
    }
    build_bfv_plain_constants(c);
}

void upload_bfv_tables(Context& c)
{
    BfvTables& t = c.bfv;
    {
        t.d_coeff_div_plainmod = upload(t.coeff_div_plainmod);
        t.d_upper_halfincrement = upload(t.upper_halfincrement);
    }
    t.d_base_change_matrix_Bsk = upload(t.base_change_matrix_Bsk);
    t.d_inv_punctured_prod_mod_base_array = upload(t.inv_punctured_prod_mod_base_array);
    t.d_base_change_matrix_m_tilde = upload(t.base_change_matrix_m_tilde);
    t.d_inv_m_tilde_mod_Bsk = upload(t.inv_m_tilde_mod_Bsk);
    t.d_prod_q_mod_Bsk = upload(t.prod_q_mod_Bsk);
    t.d_inv_prod_q_mod_Bsk = upload(t.inv_prod_q_mod_Bsk);
    t.d_base_change_matrix_q = upload(t.base_change_matrix_q);
    t.d_base_change_matrix_msk = upload(t.base_change_matrix_msk);
    t.d_inv_punctured_prod_mod_B_array = upload(t.inv_punctured_prod_mod_B_array);
    t.d_prod_B_mod_q = upload(t.prod_B_mod_q);
}

void upload_tables(Context& c)
{
    const int N = c.n;
    const int Qp = (int) c.mod.size(); // device NTT tables cover every prime (Q' chain + Bsk)
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    struct Restore {
        int prev, dev;
        ~Restore()
        {
            if (prev >= 0 && prev != dev)
                cudaSetDevice(prev);
        }
    } restore{prev_device, c.device};
    HEON_CUDA(cudaSetDevice(c.device));
    HEON_CUDA(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, c.device));
    {
        // keep stream-ordered scratch cached in the pool across synchronisation points
        // (the default release threshold of 0 hands it back to the OS at every sync)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, c.device) == cudaSuccess)
        {
            // ... unless the caller has configured the pool already (MemoryPoolConfig of the class layer)
            unsigned long long thr = 0;
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            if (thr == 0 && !getenv("HEON_POOL_KEEP_DEFAULT"))
            {
                thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
    }
    c.d_mod = upload(c.mod);
    std::vector<PrimeConst> pcs(Qp);
    std::vector<TwPair> fwd((size_t) Qp * N), inv((size_t) Qp * N), last(2 * (size_t) Qp);
    for (int i = 0; i < Qp; ++i)
    {
        const u64 p = c.mod[i].value;
        pcs[i].p = p;
        pcs[i].inv64 = shoup(1, p);
        pcs[i].r64 = (u64) ((((u128) 1) << 64) % p);
        pcs[i].r64s = shoup(pcs[i].r64, p);
        pcs[i].bits = (unsigned) c.mod[i].bit;
        pcs[i].fin_shift = pcs[i].bits - 25;
        pcs[i].fin_m = (unsigned) ((((u128) 1) << (pcs[i].bits + 31)) / p);
        pcs[i].nc_ok = pcs[i].bits <= 57 ? 1u : 0u;
        // measured on B200 (tools/microbench2.cu, microbench3.cu; SMSP cycles per warp-butterfly):
        // integer VAR 2 30.9, all-FP64 VAR 3 16.3, VAR 4 (X reduced every other stage) 19.4
        pcs[i].fp_var = !c.use_fp64 ? 0u : pcs[i].bits <= 47 ? 3u : pcs[i].bits <= 50 ? 4u : 0u;
        pcs[i].pinv = 1.0 / (double) p;
        {
            // 1/p - pinv = (1 - p*pinv)/p; the residual 1 - p*pinv is exact in one fma
            const double res = std::fma(-(double) p, pcs[i].pinv, 1.0);
            pcs[i].pinv_lo = res / (double) p;
        }
        pcs[i].pad_ = 0;
        for (int j = 0; j < N; ++j)
        {
            const size_t o = (size_t) i * N + j;
            fwd[o].w = c.ntt_table[o];
            if (pcs[i].fp_var)
            {
                // {w, RN(w/p)} as doubles: both operands are exact (< 2^50), IEEE division rounds once
                const double wd = (double) c.ntt_table[o];
                const double winv = wd / (double) p;
                std::memcpy(&fwd[o].w, &wd, 8);
                std::memcpy(&fwd[o].ws, &winv, 8);
            }
            else
                fwd[o].ws = shoup(c.ntt_table[o], p);
            inv[o].w = c.intt_table[o];
            if (pcs[i].fp_var)
            {
                const double wd = (double) c.intt_table[o];
                const double winv = wd / (double) p;
                std::memcpy(&inv[o].w, &wd, 8);
                std::memcpy(&inv[o].ws, &winv, 8);
            }
            else
                inv[o].ws = shoup(c.intt_table[o], p);
        }
        const u64 ninv = c.n_inverse[i];
        const u64 wn = mulmod(c.intt_table[(size_t) i * N + 1], ninv, p);
        last[2 * i] = TwPair{ninv, shoup(ninv, p)};
        last[2 * i + 1] = TwPair{wn, shoup(wn, p)};
        if (pcs[i].fp_var)
            for (int e = 0; e < 2; ++e)
            {
                const double wd = (double) last[2 * i + e].w, winv = wd / (double) p;
                std::memcpy(&last[2 * i + e].w, &wd, 8);
                std::memcpy(&last[2 * i + e].ws, &winv, 8);
            }
    }
    {
        const int S1 = c.logn - 8, R = 1 << S1;
        std::vector<TwPair> fb((size_t) Qp * R * 256), ib((size_t) Qp * R * 256);
        for (int i = 0; i < Qp; ++i)
            for (int r = 0; r < R; ++r)
                for (int u = 4; u < 8; ++u)
                    for (int g = 0; g < (1 << (u - 4)); ++g)
                        for (int tt = 0; tt < 16; ++tt)
                        {
                            const size_t src = (size_t) i * N + (1u << (S1 + u)) + ((size_t) r << u) + (tt << (u - 4)) + g;
                            const size_t dst = (((size_t) i * R + r) * 16 + ((1 << (u - 4)) - 1 + g)) * 16 + tt;
                            fb[dst] = fwd[src];
                            ib[dst] = inv[src];
                        }
        c.d_fwd_rowb = upload(fb);
        c.d_inv_rowb = upload(ib);
        // compact copy for the FP64 primes: bare doubles, round-A entries then the lane-major block
        std::vector<double> rc((size_t) Qp * R * 256, 0.0);
        for (int i = 0; i < Qp; ++i)
        {
            if (!pcs[i].fp_var)
                continue;
            for (int r = 0; r < R; ++r)
            {
                double* row = rc.data() + ((size_t) i * R + r) * 256;
                for (int u = 0; u < 4; ++u)
                    for (int g = 0; g < (1 << u); ++g)
                        row[(1 << u) - 1 + g] = (double) c.ntt_table[(size_t) i * N + (1u << (S1 + u)) + ((size_t) r << u) + g];
                for (int u = 4; u < 8; ++u)
                    for (int g = 0; g < (1 << (u - 4)); ++g)
                        for (int tt = 0; tt < 16; ++tt)
                            row[16 + ((1 << (u - 4)) - 1 + g) * 16 + tt] =
                                (double) c.ntt_table[(size_t) i * N + (1u << (S1 + u)) + ((size_t) r << u) + (tt << (u - 4)) + g];
            }
        }
        c.d_fwd_rowc = upload(rc);
    }
    c.d_pc = upload(pcs);
    c.d_fwd = upload(fwd);
    c.d_inv = upload(inv);
    c.d_inv_last = upload(last);
    c.d_last_q_modinv = upload(c.last_q_modinv);
    {
        // block i of last_q_modinv holds entries for primes j = 0..Qp-2-i
        std::vector<TwPair> pairs;
        size_t o = 0;
        for (int i = 0; i < c.P_size; ++i)
            for (int j = 0; j < c.Qp - 1 - i; ++j, ++o)
                pairs.push_back(TwPair{c.last_q_modinv[o], shoup(c.last_q_modinv[o], c.mod[j].value)});
        c.d_lqm_pair = upload(pairs);
        // product over the K blocks, per Q prime: the factor the NTT-domain mod-down multiplies by
        std::vector<TwPair> mprod;
        for (int y = 0; y < c.Q_size; ++y)
        {
            const u64 q = c.mod[y].value;
            u64 m = 1;
            size_t l2 = 0;
            for (int i = 0; i < c.P_size; ++i)
            {
                m = mulmod(m, c.last_q_modinv[l2 + y], q);
                l2 += c.Qp - 1 - i;
            }
            mprod.push_back(TwPair{m, shoup(m, q)});
        }
        c.d_md2_M = upload(mprod);
        // FP64 tables of the correction chain (only meaningful for primes the FP64 path handles)
        const int K = c.P_size;
        std::vector<TwPair> btab((size_t) c.Q_size * K * 2);
        std::vector<u64> cst(c.Q_size);
        for (int y = 0; y < c.Q_size; ++y)
        {
            const u64 q = c.mod[y].value;
            std::vector<u64> m(K), hm(K), B(K);
            size_t l2 = 0;
            for (int i = 0; i < K; ++i)
            {
                m[i] = c.last_q_modinv[l2 + y];
                hm[i] = c.half_mod[l2 + y];
                l2 += c.Qp - 1 - i;
            }
            u64 suffix = 1, acc = 0;
            for (int i = K - 1; i >= 0; --i)
            {
                suffix = mulmod(suffix, m[i], q);
                B[i] = suffix;
            }
            for (int i = 0; i < K; ++i)
                acc = (u64) (((u128) acc + mulmod(hm[i] % q, B[i], q)) % q);
            cst[y] = acc;
            for (int i = 0; i < K; ++i)
            {
                const u64 b30 = mulmod(B[i], (u64) ((1ull << 30) % q), q);
                const double w0 = (double) b30, w1 = (double) B[i];
                const double i0 = w0 / (double) q, i1 = w1 / (double) q;
                TwPair t0, t1;
                std::memcpy(&t0.w, &w0, 8);
                std::memcpy(&t0.ws, &i0, 8);
                std::memcpy(&t1.w, &w1, 8);
                std::memcpy(&t1.ws, &i1, 8);
                btab[((size_t) y * K + i) * 2] = t0;
                btab[((size_t) y * K + i) * 2 + 1] = t1;
            }
        }
        c.d_md2_B = upload(btab);
        c.d_md2_cst = upload(cst);
    }
    c.d_half = upload(c.half);
    c.d_half_mod = upload(c.half_mod);
    c.d_rescaled_last_q_modinv = upload(c.rescaled_last_q_modinv);
    c.d_rescaled_half_mod = upload(c.rescaled_half_mod);
    c.d_rescaled_half = upload(c.rescaled_half);
    for (auto& t : c.lvl2)
    {
        t.d_base_change = upload(t.base_change);
        t.d_mi_inv = upload(t.mi_inv);
        t.d_prod = upload(t.prod);
        {
            const int depth = (int) (&t - &c.lvl2[0]);
            const int L = c.Q_size - depth, K = c.P_size, Ql = L + K;
            std::vector<TwPair> mp;
            for (size_t i = 0; i < t.mi_inv.size(); ++i) // digit primes are q_i (same index at every depth)
                mp.push_back(TwPair{t.mi_inv[i] % c.mod[i].value, shoup(t.mi_inv[i] % c.mod[i].value, c.mod[i].value)});
            t.d_mi_inv_pair = upload(mp);
            // base_change layout [digit][k over Q'_l][i in digit]: Shoup word for target prime t_k
            std::vector<TwPair> bp;
            {
                size_t o = 0;
                auto fp_ok = [&](int prime) { return c.use_fp64 && c.mod[prime].bit <= 50; };
                for (int l = 0; l < t.d; ++l)
                {
                    bool dfp = t.I_j[l] <= 4; // k_modup2 uses the FP64 pipe when digit and target primes allow it
                    for (int i = 0; i < t.I_j[l]; ++i)
                        dfp = dfp && fp_ok(t.I_loc[l] + i);
                    for (int k = 0; k < Ql; ++k)
                    {
                        const int pk = level_prime(k, L, depth);
                        const u64 tk = c.mod[pk].value;
                        for (int i = 0; i < t.I_j[l]; ++i, ++o)
                        {
                            const u64 bcw = t.base_change[o] % tk; // canonical residue of the exported word
                            if (dfp && fp_ok(pk))
                            {
                                const double md = (double) bcw, minv = md / (double) tk;
                                TwPair tp;
                                std::memcpy(&tp.w, &md, 8);
                                std::memcpy(&tp.ws, &minv, 8);
                                bp.push_back(tp);
                            }
                            else
                                bp.push_back(TwPair{bcw, shoup(bcw, tk)});
                        }
                    }
                }
            }
            t.d_base_change_pair = upload(bp);
            std::vector<u64> rp((size_t) (K + 1) * t.d * Ql);
            for (int r = 0; r <= K; ++r)
                for (int dg = 0; dg < t.d; ++dg)
                    for (int k = 0; k < Ql; ++k)
                    {
                        const u64 tk = c.mod[level_prime(k, L, depth)].value;
                        rp[((size_t) r * t.d + dg) * Ql + k] = mulmod((u64) r, t.prod[(size_t) dg * Ql + k] % tk, tk);
                    }
            t.d_rprod = upload(rp);
        }
        t.d_I_j = upload(t.I_j);
        t.d_I_loc = upload(t.I_loc);
    }
}

Context::~Context()
{
    if (device < 0)
        return; // host-only context: nothing was uploaded, and no CUDA call may touch this thread's error state
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    struct Restore {
        int prev, dev;
        ~Restore()
        {
            if (prev >= 0 && prev != dev)
                cudaSetDevice(prev);
        }
    } restore{prev_device, device};
    if (prev_device != device)
        cudaSetDevice(device);
    cudaFree(d_mod);
    cudaFree(d_pc);
    cudaFree(d_fwd);
    cudaFree(d_inv);
    cudaFree(d_inv_last);
    cudaFree(d_fwd_rowb);
    cudaFree(d_inv_rowb);
    cudaFree(d_fwd_rowc);
    cudaFree(d_last_q_modinv);
    cudaFree(d_lqm_pair);
    cudaFree(d_md2_M);
    cudaFree(d_md2_B);
    cudaFree(d_md2_cst);
    cudaFree(d_half);
    cudaFree(d_half_mod);
    cudaFree(d_rescaled_last_q_modinv);
    cudaFree(d_rescaled_half_mod);
    cudaFree(d_rescaled_half);
    cudaFree(bfv.d_base_change_matrix_Bsk);
    cudaFree(bfv.d_inv_punctured_prod_mod_base_array);
    cudaFree(bfv.d_base_change_matrix_m_tilde);
    cudaFree(bfv.d_inv_m_tilde_mod_Bsk);
    cudaFree(bfv.d_prod_q_mod_Bsk);
    cudaFree(bfv.d_inv_prod_q_mod_Bsk);
    cudaFree(bfv.d_base_change_matrix_q);
    cudaFree(bfv.d_base_change_matrix_msk);
    cudaFree(bfv.d_inv_punctured_prod_mod_B_array);
    cudaFree(bfv.d_prod_B_mod_q);
    cudaFree(bfv.d_coeff_div_plainmod);
    cudaFree(bfv.d_upper_halfincrement);
    for (auto& t : lvl2)
    {
        cudaFree(t.d_base_change);
        cudaFree(t.d_mi_inv);
        cudaFree(t.d_prod);
        cudaFree(t.d_mi_inv_pair);
        cudaFree(t.d_base_change_pair);
        cudaFree(t.d_rprod);
        cudaFree(t.d_I_j);
        cudaFree(t.d_I_loc);
    }
}

} // namespace heon
