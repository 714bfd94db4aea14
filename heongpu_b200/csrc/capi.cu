// extern "C" boundary: argument validation, exception -> status translation.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <nvtx3/nvToolsExt.h>
#include <zlib.h>
#include "../../include/heon_b200.h"
#include "ops.hpp"

using namespace heon;

struct heon_context_s {
    Context c;
};

namespace heon {
std::atomic<long long> g_launches{0};

static std::atomic<bool> g_profiling{false}; // operators may run on several host threads (one stream each)
struct ProfRec {
    int cls;
    cudaEvent_t e0, e1;
};
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

LaunchScope::LaunchScope(int c, cudaStream_t s) : cls(c), st(s)
{
    g_launches++;
    if (g_profiling)
    {
        cudaEventCreate(&e0);
        cudaEventRecord(e0, st);
    }
}
LaunchScope::~LaunchScope()
{
    if (e0)
    {
        cudaEvent_t e1;
        cudaEventCreate(&e1);
        cudaEventRecord(e1, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(ProfRec{cls, e0, e1});
    }
}
void profile_begin()
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.clear();
    g_profiling = true;
}
void profile_end(double* ms, long long* launches)
{
    g_profiling = false;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < KC_COUNT; ++i)
    {
        ms[i] = 0;
        launches[i] = 0;
    }
    for (auto& r : g_prof)
    {
        float t = 0;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        ms[r.cls] += t;
        launches[r.cls]++;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
}
} // namespace heon

static thread_local std::string g_err;

// One NVTX range per ABI call (header-only NVTX3: a no-op unless a profiler is attached), named after the
// entry point, so Nsight timelines show the operator structure the reference's tracing shows.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

template <class F> static int guarded_named(const char* name, F&& f)
{
    NvtxRange range(name);
    try
    {
        f();
        return HEON_OK;
    }
    catch (const std::invalid_argument& e)
    {
        g_err = e.what();
        return HEON_ERR_INVALID;
    }
    catch (const std::logic_error& e)
    {
        g_err = e.what();
        return HEON_ERR_LOGIC;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return HEON_ERR_RUNTIME;
    }
}
#define guarded(...) guarded_named(__func__, __VA_ARGS__)

// Makes the context's device current for the duration of one ABI call and restores the caller's
// device afterwards; also drops any stale error a previous, unrelated CUDA call left on this thread, so
// that the launch checks of this call (cudaGetLastError) only ever report this call's own launches.
struct DeviceGuard {
    int prev = -1, dev = -1;
    explicit DeviceGuard(const Context& c)
    {
        if (c.device < 0)
            throw std::runtime_error("context was created host-only (device < 0): no CUDA path");
        dev = c.device;
        cudaGetDevice(&prev);
        if (prev != dev)
        {
            cudaError_t e = cudaSetDevice(dev);
            if (e != cudaSuccess)
                throw std::runtime_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        }
        (void) cudaGetLastError();
    }
    ~DeviceGuard()
    {
        if (prev >= 0 && prev != dev)
            cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

static void finish_context(Context& c, int device, int log_n, int n_q, int n_p, u64 plain_modulus = 0)
{
    if (log_n < 12 || log_n > 16)
        throw std::logic_error("Poly modulus degree is not supported");
    if (n_p < 1)
        throw std::logic_error("log_P_bases_bit_sizes cannot be empty!");
    if (n_q < 1 || n_q + n_p > 128)
        throw std::logic_error("invalid modulus count");
    c.device = device;
    c.scheme = plain_modulus ? SCHEME_BFV : SCHEME_CKKS;
    c.plain_modulus = plain_modulus;
    c.logn = log_n;
    c.n = 1 << log_n;
    c.Q_size = n_q;
    c.P_size = n_p;
    c.Qp = n_q + n_p;
    c.method = (n_p == 1) ? 1 : 2;
    if (c.method == 2 && n_p > 15)
        throw std::logic_error("P size above 15 is not supported");
    if (const char* v = getenv("HEON_NTT_VARIANT"))
        c.ntt_variant = atoi(v);
    if (const char* v = getenv("HEON_NTT_PERSISTENT"))
        c.ntt_persistent = atoi(v);
    if (const char* v = getenv("HEON_NTT_FP64"))
        c.use_fp64 = atoi(v);
    if (const char* v = getenv("HEON_COL_THREADS"))
        c.col_threads = atoi(v);
    if (const char* v = getenv("HEON_ROW_TILE"))
        c.row_tile = atoi(v);
    if (const char* v = getenv("HEON_SKIP_OWN"))
        c.skip_own = atoi(v);
    if (const char* v = getenv("HEON_COL_TMA"))
        c.col_tma = atoi(v);
    if (const char* v = getenv("HEON_COL_TMA_TILES"))
        c.col_tma_tiles = atoi(v);
    if (const char* v = getenv("HEON_MODUP_DOUBLES"))
        c.modup_doubles = atoi(v);
    if (const char* v = getenv("HEON_ROW_WALK"))
        c.row_walk = atoi(v);
    if (const char* v = getenv("HEON_COL_TMA_BUFS"))
        c.col_tma_bufs = atoi(v);
    if (const char* v = getenv("HEON_ROW_MAC"))
        c.row_mac = atoi(v);
    if (const char* v = getenv("HEON_MODUP_FUSED"))
        c.modup_fused = atoi(v);
    if (const char* v = getenv("HEON_MODUP_CW"))
        c.modup_cw = atoi(v);
    if (const char* v = getenv("HEON_MODUP_COL"))
        c.modup_col = atoi(v);
    if (const char* v = getenv("HEON_ROW_FINAL"))
        c.row_final = atoi(v);
    if (const char* v = getenv("HEON_ROW_MAC_ROWS"))
        c.row_mac_rows = atoi(v);
    if (const char* v = getenv("HEON_GALOIS_NTT"))
        c.galois_ntt = atoi(v);
    if (const char* v = getenv("HEON_NTT_PIPE"))
        c.ntt_pipe = atoi(v);
    if (const char* v = getenv("HEON_NTT_GROUP"))
        c.ntt_group = atoi(v);
    if (const char* v = getenv("HEON_NTT_FUSED"))
        c.ntt_fused = atoi(v);
    if (const char* v = getenv("HEON_NTT_TMA"))
        c.use_tma = atoi(v);
    if (c.scheme == SCHEME_BFV)
    {
        // auxiliary base Bsk: the largest 61-bit primes, one more than needed (the largest is the
        // gamma prime of the decryptor and is not part of Bsk); bfv/context.cu:518-531, util.cu:278-310
        int total_bits = 0;
        for (int i = 0; i < n_q + n_p; ++i)
            total_bits += (int) c.mod[i].bit;
        c.bsk = n_q + n_p;
        if (bit_length(plain_modulus) + total_bits + 32 >= 61 * n_q + 61)
            c.bsk++;
        std::vector<u64> big = largest_ntt_primes(2ull << log_n, 61, (size_t) c.bsk + 1);
        for (int i = 0; i < c.bsk; ++i)
            c.mod.push_back(make_mod(big[c.bsk - i])); // ascending; big[0] (largest) is gamma
        if ((int) c.mod.size() > 128)
            throw std::logic_error("invalid modulus count");
    }
    build_host_tables(c);
    if (c.scheme == SCHEME_BFV)
        build_bfv_tables(c);
    if (device >= 0)
    {
        upload_tables(c);
        if (c.scheme == SCHEME_BFV)
            upload_bfv_tables(c);
    }
}

// ---- client side (client.cu) ----
namespace heon {
void client_keygen_secret(const Context& c, u64 seed, int hamming_weight, u64* sk, cudaStream_t st);
void client_keygen_public(const Context& c, const u64* sk, u64 seed, u64* pk, cudaStream_t st);
void client_keygen_evk(const Context& c, const u64* under, const u64* target, u64 seed, u64* key, cudaStream_t st);
void client_keygen_relin(const Context& c, const u64* sk, u64 seed, u64* key, cudaStream_t st);
void client_keygen_galois(const Context& c, const u64* sk, unsigned galois_elt, u64 seed, u64* key, cudaStream_t st);
void client_encrypt(const Context& c, const u64* pk, const u64* pt, u64 seed, u64* ct, cudaStream_t st);
void client_decrypt_ckks(const Context& c, const u64* sk, const u64* ct, int comps, int depth, u64* pt, cudaStream_t st);
void client_decrypt_bfv(const Context& c, const u64* sk, const u64* ct, int comps, u64* pt, cudaStream_t st, int* budget = nullptr);
void client_noise_budget_bfv(const Context& c, const u64* sk, const u64* ct, int comps, int* bits, cudaStream_t st);
void client_ckks_encode(const Context& c, const double* values, int count, double scale, int depth, u64* pt, cudaStream_t st);
void client_ckks_decode(const Context& c, const u64* pt, int depth, double scale, double* out, int count, cudaStream_t st);
void client_bfv_encode(const Context& c, const u64* msg, int count, u64* pt, cudaStream_t st);
void client_bfv_decode(const Context& c, const u64* pt, u64* msg, int count, cudaStream_t st);
} // namespace heon

extern "C" {

const char* heon_last_error(void) { return g_err.c_str(); }
const char* heon_version(void) { return "heon-b200 0.1 (sm_100a)"; }

int heon_ckks_context_create(int device, int log_n, const int* q_bits, int n_q, const int* p_bits,
                             int n_p, heon_context_t* out)
{
    return guarded([&] {
        if (!out || !q_bits || !p_bits)
            throw std::invalid_argument("null argument");
        auto h = std::make_unique<heon_context_s>();
        std::vector<int> bits(q_bits, q_bits + n_q);
        bits.insert(bits.end(), p_bits, p_bits + n_p);
        if (log_n < 12 || log_n > 16)
            throw std::logic_error("Poly modulus degree is not supported");
        for (u64 p : primes_for_bit_sizes(1ull << log_n, bits))
            h->c.mod.push_back(make_mod(p));
        finish_context(h->c, device, log_n, n_q, n_p);
        *out = h.release();
    });
}

int heon_ckks_context_create_values(int device, int log_n, const uint64_t* q, int n_q,
                                    const uint64_t* p, int n_p, heon_context_t* out)
{
    return guarded([&] {
        if (!out || !q || !p)
            throw std::invalid_argument("null argument");
        if (log_n < 12 || log_n > 16)
            throw std::logic_error("Poly modulus degree is not supported");
        auto h = std::make_unique<heon_context_s>();
        for (int i = 0; i < n_q; ++i)
            h->c.mod.push_back(make_mod(q[i]));
        for (int i = 0; i < n_p; ++i)
            h->c.mod.push_back(make_mod(p[i]));
        for (auto& m : h->c.mod)
            if (m.bit > 61 || m.bit < 30 || !is_prime_u64(m.value) ||
                (m.value - 1) % (2ull << log_n))
                throw std::logic_error("invalid modulus");
        finish_context(h->c, device, log_n, n_q, n_p);
        *out = h.release();
    });
}

int heon_bfv_context_create(int device, int log_n, const int* q_bits, int n_q, const int* p_bits, int n_p,
                            uint64_t plain_modulus, heon_context_t* out)
{
    return guarded([&] {
        if (!out || !q_bits || !p_bits)
            throw std::invalid_argument("null argument");
        if (plain_modulus < 2)
            throw std::logic_error("invalid plain modulus");
        auto h = std::make_unique<heon_context_s>();
        std::vector<int> bits(q_bits, q_bits + n_q);
        bits.insert(bits.end(), p_bits, p_bits + n_p);
        if (log_n < 12 || log_n > 16)
            throw std::logic_error("Poly modulus degree is not supported");
        for (u64 p : primes_for_bit_sizes(1ull << log_n, bits))
            h->c.mod.push_back(make_mod(p));
        finish_context(h->c, device, log_n, n_q, n_p, plain_modulus);
        *out = h.release();
    });
}

int heon_bfv_context_create_values(int device, int log_n, const uint64_t* q, int n_q, const uint64_t* p, int n_p,
                                   uint64_t plain_modulus, heon_context_t* out)
{
    return guarded([&] {
        if (!out || !q || !p)
            throw std::invalid_argument("null argument");
        if (plain_modulus < 2)
            throw std::logic_error("invalid plain modulus");
        if (log_n < 12 || log_n > 16)
            throw std::logic_error("Poly modulus degree is not supported");
        auto h = std::make_unique<heon_context_s>();
        for (int i = 0; i < n_q; ++i)
            h->c.mod.push_back(make_mod(q[i]));
        for (int i = 0; i < n_p; ++i)
            h->c.mod.push_back(make_mod(p[i]));
        for (auto& m : h->c.mod)
            if (m.bit > 61 || m.bit < 30 || !is_prime_u64(m.value) || (m.value - 1) % (2ull << log_n))
                throw std::logic_error("invalid modulus");
        finish_context(h->c, device, log_n, n_q, n_p, plain_modulus);
        *out = h.release();
    });
}

int heon_bfv_multiply(heon_context_t ctx, const uint64_t* a, long long as, const uint64_t* b, long long bs,
                      uint64_t* out, long long os, int batch, void* stream)
{
    return guarded([&] {
        if (!ctx || !a || !b || !out)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (batch < 1)
            throw std::invalid_argument("batch must be positive");
        op_bfv_multiply(c, a, as, b, bs, out, os, batch, (cudaStream_t) stream);
    });
}

int heon_bfv_relinearize(heon_context_t ctx, uint64_t* ct, long long cs, const uint64_t* relin_key, int batch,
                         void* stream)
{
    return guarded([&] {
        if (!ctx || !ct || !relin_key)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (batch < 1)
            throw std::invalid_argument("batch must be positive");
        op_bfv_relinearize(c, ct, cs, relin_key, batch, (cudaStream_t) stream);
    });
}

int heon_bfv_apply_galois(heon_context_t ctx, const uint64_t* in, long long is, uint64_t* out, long long os,
                          const uint64_t* galois_key, uint32_t galois_elt, int batch, void* stream)
{
    return guarded([&] {
        if (!ctx || !in || !out || !galois_key || in == out)
            throw std::invalid_argument("invalid buffers");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (c.scheme != SCHEME_BFV)
            throw std::invalid_argument("not a BFV context");
        if (batch < 1)
            throw std::invalid_argument("batch must be positive");
        op_apply_galois(c, in, is, out, os, galois_key, galois_elt, 0, batch, (cudaStream_t) stream);
    });
}

int heon_bfv_add_plain(heon_context_t ctx, const uint64_t* ct, long long cs, const uint64_t* pt, long long ps,
                       uint64_t* out, long long os, int comps, int batch, void* stream)
{
    return guarded([&] {
        if (!ctx || !ct || !pt || !out)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (batch < 1)
            throw std::invalid_argument("batch must be positive");
        op_bfv_addsub_plain(c, ct, cs, pt, ps, out, os, comps, batch, 1, (cudaStream_t) stream);
    });
}
int heon_bfv_sub_plain(heon_context_t ctx, const uint64_t* ct, long long cs, const uint64_t* pt, long long ps,
                       uint64_t* out, long long os, int comps, int batch, void* stream)
{
    return guarded([&] {
        if (!ctx || !ct || !pt || !out)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (batch < 1)
            throw std::invalid_argument("batch must be positive");
        op_bfv_addsub_plain(c, ct, cs, pt, ps, out, os, comps, batch, 2, (cudaStream_t) stream);
    });
}
int heon_bfv_multiply_plain(heon_context_t ctx, const uint64_t* ct, long long cs, const uint64_t* pt, long long ps,
                            uint64_t* out, long long os, int batch, void* stream)
{
    return guarded([&] {
        if (!ctx || !ct || !pt || !out || ct == out)
            throw std::invalid_argument("invalid buffers");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (batch < 1)
            throw std::invalid_argument("batch must be positive");
        op_bfv_multiply_plain(c, ct, cs, pt, ps, out, os, batch, (cudaStream_t) stream);
    });
}

int heon_bfv_keyswitch(heon_context_t ctx, const uint64_t* in, long long is, uint64_t* out, long long os,
                       const uint64_t* switch_key, int batch, void* stream)
{
    return guarded([&] {
        if (!ctx || !in || !out || !switch_key || in == out)
            throw std::invalid_argument("invalid buffers");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (c.scheme != SCHEME_BFV)
            throw std::invalid_argument("not a BFV context");
        if (batch < 1)
            throw std::invalid_argument("batch must be positive");
        // the automorphism pipeline with the identity element: (c0, 0) + KeySwitch(c1), coefficient domain
        op_apply_galois(c, in, is, out, os, switch_key, 1u, 0, batch, (cudaStream_t) stream);
    });
}

void heon_context_destroy(heon_context_t ctx) { delete ctx; }

int heon_context_info(heon_context_t ctx, heon_info* o)
{
    return guarded([&] {
        if (!ctx || !o)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        *o = heon_info{c.scheme, c.n, c.logn, c.Q_size, c.P_size, c.method, c.device};
    });
}

int heon_context_table(heon_context_t ctx, int which, int depth, uint64_t* h_out, size_t cap,
                       size_t* count)
{
    return guarded([&] {
        if (!ctx || !count)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        std::vector<u64> tmp;
        const std::vector<u64>* src = nullptr;
        auto lvl = [&]() -> const LevelTablesII& {
            if (c.method != 2 || depth < 0 || depth >= (int) c.lvl2.size())
                throw std::invalid_argument("no Method II table at this depth");
            return c.lvl2[depth];
        };
        switch (which)
        {
            case HEON_TBL_MODULUS:
                for (auto& m : c.mod)
                {
                    tmp.push_back(m.value);
                    tmp.push_back(m.bit);
                    tmp.push_back(m.mu);
                }
                src = &tmp;
                break;
            case HEON_TBL_PSI: src = &c.psi; break;
            case HEON_TBL_NTT: src = &c.ntt_table; break;
            case HEON_TBL_INTT: src = &c.intt_table; break;
            case HEON_TBL_N_INVERSE: src = &c.n_inverse; break;
            case HEON_TBL_LAST_Q_MODINV: src = &c.last_q_modinv; break;
            case HEON_TBL_HALF: src = &c.half; break;
            case HEON_TBL_HALF_MOD: src = &c.half_mod; break;
            case HEON_TBL_FACTOR: src = &c.factor; break;
            case HEON_TBL_RESCALED_LAST_Q_MODINV: src = &c.rescaled_last_q_modinv; break;
            case HEON_TBL_RESCALED_HALF_MOD: src = &c.rescaled_half_mod; break;
            case HEON_TBL_RESCALED_HALF: src = &c.rescaled_half; break;
            case HEON_TBL_II_BASE_CHANGE: src = &lvl().base_change; break;
            case HEON_TBL_II_MI_INV: src = &lvl().mi_inv; break;
            case HEON_TBL_II_PROD: src = &lvl().prod; break;
            case HEON_TBL_II_I_J:
                for (int v : lvl().I_j)
                    tmp.push_back((u64) v);
                src = &tmp;
                break;
            case HEON_TBL_II_I_LOCATION:
                for (int v : lvl().I_loc)
                    tmp.push_back((u64) v);
                src = &tmp;
                break;
            case HEON_TBL_BFV_BASE_CHANGE_BSK: src = &c.bfv.base_change_matrix_Bsk; break;
            case HEON_TBL_BFV_INV_PUNCT_Q: src = &c.bfv.inv_punctured_prod_mod_base_array; break;
            case HEON_TBL_BFV_BASE_CHANGE_MTILDE: src = &c.bfv.base_change_matrix_m_tilde; break;
            case HEON_TBL_BFV_INV_MTILDE_MOD_BSK: src = &c.bfv.inv_m_tilde_mod_Bsk; break;
            case HEON_TBL_BFV_PROD_Q_MOD_BSK: src = &c.bfv.prod_q_mod_Bsk; break;
            case HEON_TBL_BFV_INV_PROD_Q_MOD_BSK: src = &c.bfv.inv_prod_q_mod_Bsk; break;
            case HEON_TBL_BFV_BASE_CHANGE_Q: src = &c.bfv.base_change_matrix_q; break;
            case HEON_TBL_BFV_BASE_CHANGE_MSK: src = &c.bfv.base_change_matrix_msk; break;
            case HEON_TBL_BFV_INV_PUNCT_B: src = &c.bfv.inv_punctured_prod_mod_B_array; break;
            case HEON_TBL_BFV_PROD_B_MOD_Q: src = &c.bfv.prod_B_mod_q; break;
            case HEON_TBL_BFV_SCALARS:
                tmp = {c.bfv.inv_prod_q_mod_m_tilde, c.bfv.inv_prod_B_mod_m_sk, (u64) c.bsk, c.plain_modulus};
                src = &tmp;
                break;
            case HEON_TBL_BFV_PLAIN:
                tmp = c.bfv.coeff_div_plainmod;
                tmp.insert(tmp.end(), c.bfv.upper_halfincrement.begin(), c.bfv.upper_halfincrement.end());
                tmp.push_back(c.bfv.Q_mod_t);
                tmp.push_back(c.bfv.upper_threshold);
                src = &tmp;
                break;
            default: throw std::invalid_argument("unknown table");
        }
        *count = src->size();
        if (h_out)
        {
            if (cap < src->size())
                throw std::invalid_argument("output buffer too small");
            std::memcpy(h_out, src->data(), src->size() * sizeof(u64));
        }
    });
}

int heon_steps_to_galois_elt(int steps, int n, int group_order)
{
    const int m = 2 * n;
    if (steps == 0)
        return m - 1;
    const int pos = steps < 0 ? -steps : steps;
    if (pos >= (n >> 1))
        return 0;
    int s = steps < 0 ? (n >> 1) - pos : pos;
    int g = 1;
    while (s-- > 0)
        g = (int) (((long long) g * group_order) & (m - 1));
    return g;
}

int heon_ntt(heon_context_t ctx, const uint64_t* in, uint64_t* out, long long n_polys,
             const int* h_prime_index, int mod_count, int inverse, void* stream)
{
    return guarded([&] {
        if (!ctx || !in || !out)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (mod_count < 1 || mod_count > 128)
            throw std::invalid_argument("invalid mod_count");
        PrimeList pl;
        pl.count = mod_count;
        for (int i = 0; i < mod_count; ++i)
        {
            int v = h_prime_index ? h_prime_index[i] : i;
            if (v < 0 || v >= c.Qp)
                throw std::invalid_argument("prime index out of range");
            pl.idx[i] = (unsigned char) v;
        }
        launch_ntt(c, in, out, n_polys, pl, inverse != 0, (cudaStream_t) stream);
    });
}

int heon_ntt_poly_ordered(heon_context_t ctx, uint64_t* base, const long long* h_offsets,
                          int n_polys, int prime_index, int inverse, void* stream)
{
    return guarded([&] {
        if (!ctx || !base || !h_offsets)
            throw std::invalid_argument("null argument");
        const Context& c = ctx->c;
        DeviceGuard dev_guard(c);
        if (prime_index < 0 || prime_index >= c.Qp)
            throw std::invalid_argument("prime index out of range");
        cudaStream_t st = (cudaStream_t) stream;
        long long* d_off = nullptr;
        if (cudaMallocAsync(&d_off, sizeof(long long) * n_polys, st) != cudaSuccess)
            throw std::runtime_error("cudaMallocAsync failed");
        cudaMemcpyAsync(d_off, h_offsets, sizeof(long long) * n_polys, cudaMemcpyHostToDevice, st);
        long long ext = 0;
        bool aligned = true;
        for (int i = 0; i < n_polys; ++i)
        {
            ext = std::max(ext, h_offsets[i] + c.n);
            aligned &= (h_offsets[i] & 15) == 0 && h_offsets[i] >= 0;
        }
        launch_ntt_scattered(c, base, d_off, n_polys, prime_index, inverse != 0, ext, aligned, st);
        cudaFreeAsync(d_off, st);
    });
}

#define HEON_OP_PROLOGUE                                                                           \
    if (!ctx)                                                                                      \
        throw std::invalid_argument("null context");                                               \
    const Context& c = ctx->c;                                                                     \
    DeviceGuard dev_guard(c);                                                                                \
    if (batch < 1)                                                                                 \
        throw std::invalid_argument("batch must be positive");                                     \
    cudaStream_t st = (cudaStream_t) stream;

int heon_add(heon_context_t ctx, const uint64_t* a, long long as, const uint64_t* b, long long bs,
             uint64_t* out, long long os, int comps, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!a || !b || !out)
            throw std::invalid_argument("null argument");
        if (comps < 1 || comps > 3)
            throw std::invalid_argument("Invalid Ciphertexts size!");
        op_add(c, a, as, b, bs, out, os, comps, depth, batch, 0, st);
    });
}
int heon_sub(heon_context_t ctx, const uint64_t* a, long long as, const uint64_t* b, long long bs,
             uint64_t* out, long long os, int comps, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!a || !b || !out)
            throw std::invalid_argument("null argument");
        if (comps < 1 || comps > 3)
            throw std::invalid_argument("Invalid Ciphertexts size!");
        op_add(c, a, as, b, bs, out, os, comps, depth, batch, 1, st);
    });
}
int heon_negate(heon_context_t ctx, const uint64_t* a, long long as, uint64_t* out, long long os,
                int comps, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!a || !out)
            throw std::invalid_argument("null argument");
        if (comps < 1 || comps > 3)
            throw std::invalid_argument("Invalid Ciphertexts size!");
        op_add(c, a, as, a, as, out, os, comps, depth, batch, 2, st);
    });
}

int heon_ckks_multiply(heon_context_t ctx, const uint64_t* a, long long as, const uint64_t* b,
                       long long bs, uint64_t* out, long long os, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!a || !b || !out)
            throw std::invalid_argument("null argument");
        op_multiply(c, a, as, b, bs, out, os, depth, batch, st);
    });
}

int heon_ckks_multiply_plain(heon_context_t ctx, const uint64_t* ct, long long cs, const uint64_t* pt,
                             long long ps, uint64_t* out, long long os, int comps, int depth, int batch,
                             void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!ct || !pt || !out)
            throw std::invalid_argument("null argument");
        op_plain(c, ct, cs, pt, ps, out, os, comps, depth, batch, 0, st);
    });
}
int heon_ckks_add_plain(heon_context_t ctx, const uint64_t* ct, long long cs, const uint64_t* pt, long long ps,
                        uint64_t* out, long long os, int comps, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!ct || !pt || !out)
            throw std::invalid_argument("null argument");
        op_plain(c, ct, cs, pt, ps, out, os, comps, depth, batch, 1, st);
    });
}
int heon_ckks_sub_plain(heon_context_t ctx, const uint64_t* ct, long long cs, const uint64_t* pt, long long ps,
                        uint64_t* out, long long os, int comps, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!ct || !pt || !out)
            throw std::invalid_argument("null argument");
        op_plain(c, ct, cs, pt, ps, out, os, comps, depth, batch, 2, st);
    });
}

int heon_ckks_keyswitch(heon_context_t ctx, const uint64_t* in, long long is, uint64_t* out, long long os,
                        const uint64_t* switch_key, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!in || !out || !switch_key || in == out)
            throw std::invalid_argument("keyswitch needs distinct, non-null buffers");
        op_keyswitch(c, in, is, out, os, switch_key, depth, batch, st);
    });
}

int heon_ckks_conjugate(heon_context_t ctx, const uint64_t* in, long long is, uint64_t* out, long long os,
                        const uint64_t* conjugate_key, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!in || !out || !conjugate_key || in == out)
            throw std::invalid_argument("conjugate needs distinct, non-null buffers");
        if (c.scheme != SCHEME_CKKS)
            throw std::invalid_argument("not a CKKS context");
        // galois_elt_zero = 2N - 1 (ckks/evaluationkey.cu: conjugation key)
        op_apply_galois(c, in, is, out, os, conjugate_key, (unsigned) (2 * c.n - 1), depth, batch, st);
    });
}

int heon_ckks_rotate_hoisted(heon_context_t ctx, const uint64_t* in, long long is, uint64_t* out, long long os,
                             long long out_rot_stride, const uint64_t* const* h_galois_keys,
                             const uint32_t* h_galois_elts, int count, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!in || !out || !h_galois_keys || !h_galois_elts)
            throw std::invalid_argument("null argument");
        for (int r = 0; r < count; ++r)
            if (!h_galois_keys[r])
                throw std::logic_error("Galois key not present!");
        op_rotate_hoisted(c, in, is, out, os, out_rot_stride, (const u64* const*) h_galois_keys, h_galois_elts, count,
                          depth, batch, st);
    });
}

int heon_ckks_multiply_matrix(heon_context_t ctx, const uint64_t* in, uint64_t* out, const uint64_t* diags,
                              const uint32_t* h_baby_elts, const uint64_t* const* h_baby_keys, int n1,
                              const uint32_t* h_giant_elts, const uint64_t* const* h_giant_keys,
                              const int* h_group_sizes, const int* h_term_baby, int n2, int depth, void* stream)
{
    return guarded([&] {
        const int batch = 1;
        HEON_OP_PROLOGUE
        if (!in || !out || !diags || !h_baby_elts || !h_baby_keys || !h_giant_elts || !h_giant_keys || !h_group_sizes ||
            !h_term_baby)
            throw std::invalid_argument("null argument");
        op_bsgs_matvec(c, in, out, diags, h_baby_elts, (const u64* const*) h_baby_keys, n1, h_giant_elts,
                       (const u64* const*) h_giant_keys, h_group_sizes, h_term_baby, n2, depth, st);
    });
}

int heon_ckks_multiply_plain_accumulate(heon_context_t ctx, const uint64_t* cts, const uint64_t* pts, uint64_t* out,
                                        int count, int depth, void* stream)
{
    return guarded([&] {
        const int batch = 1;
        HEON_OP_PROLOGUE
        if (!cts || !pts || !out)
            throw std::invalid_argument("null argument");
        op_multiply_plain_accumulate(c, cts, pts, out, count, depth, st);
    });
}

int heon_ckks_relinearize(heon_context_t ctx, uint64_t* ct, long long cs, const uint64_t* relin_key,
                          int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!ct || !relin_key)
            throw std::invalid_argument("null argument");
        op_relinearize(c, ct, cs, relin_key, depth, batch, st);
    });
}

int heon_ckks_rescale(heon_context_t ctx, uint64_t* ct, long long cs, int depth, int batch,
                      void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!ct)
            throw std::invalid_argument("null argument");
        op_rescale(c, ct, cs, depth, batch, st);
    });
}

int heon_ckks_mod_drop_inplace(heon_context_t ctx, uint64_t* ct, long long cs, int comps, int depth,
                               int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!ct)
            throw std::invalid_argument("null argument");
        if (comps < 1 || comps > 3)
            throw std::invalid_argument("Invalid Ciphertexts size!");
        op_mod_drop_inplace(c, ct, cs, comps, depth, batch, st);
    });
}

int heon_ckks_mod_drop(heon_context_t ctx, const uint64_t* in, long long is, uint64_t* out,
                       long long os, int depth, int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!in || !out)
            throw std::invalid_argument("null argument");
        op_mod_drop(c, in, is, out, os, depth, batch, st);
    });
}

int heon_ckks_apply_galois(heon_context_t ctx, const uint64_t* in, long long is, uint64_t* out,
                           long long os, const uint64_t* galois_key, uint32_t galois_elt, int depth,
                           int batch, void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!in || !out || !galois_key || in == out)
            throw std::invalid_argument("invalid buffers");
        op_apply_galois(c, in, is, out, os, galois_key, galois_elt, depth, batch, st);
    });
}

// ---- client side (client.cu) ----

#define HEON_CLIENT_PROLOGUE                                                                       \
    if (!ctx)                                                                                      \
        throw std::invalid_argument("null context");                                               \
    const Context& c = ctx->c;                                                                     \
    DeviceGuard dev_guard(c);                                                                      \
    cudaStream_t st = (cudaStream_t) stream;

int heon_ckks_multiply_relinearize_host(heon_context_t ctx, const uint64_t* h_a, const uint64_t* h_b, uint64_t* h_out,
                                        const uint64_t* relin_key, int depth, int rescale, int batch, int chunk,
                                        void* stream)
{
    return guarded([&] {
        HEON_OP_PROLOGUE
        if (!h_a || !h_b || !h_out || !relin_key)
            throw std::invalid_argument("null argument");
        if (c.scheme != SCHEME_CKKS)
            throw std::invalid_argument("not a CKKS context");
        op_mulrelin_host(c, h_a, h_b, h_out, relin_key, depth, rescale, batch, chunk, st);
    });
}

int heon_keygen_secret(heon_context_t ctx, uint64_t seed, int hamming_weight, uint64_t* sk, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!sk)
            throw std::invalid_argument("null argument");
        client_keygen_secret(c, seed, hamming_weight, sk, st);
    });
}
int heon_keygen_public(heon_context_t ctx, const uint64_t* sk, uint64_t seed, uint64_t* pk, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!sk || !pk)
            throw std::invalid_argument("null argument");
        client_keygen_public(c, sk, seed, pk, st);
    });
}
int heon_keygen_relin(heon_context_t ctx, const uint64_t* sk, uint64_t seed, uint64_t* key, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!sk || !key)
            throw std::invalid_argument("null argument");
        client_keygen_relin(c, sk, seed, key, st);
    });
}
int heon_keygen_galois(heon_context_t ctx, const uint64_t* sk, uint32_t galois_elt, uint64_t seed, uint64_t* key,
                       void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!sk || !key)
            throw std::invalid_argument("null argument");
        client_keygen_galois(c, sk, galois_elt, seed, key, st);
    });
}
int heon_keygen_switch(heon_context_t ctx, const uint64_t* new_sk, const uint64_t* old_sk, uint64_t seed,
                       uint64_t* key, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!new_sk || !old_sk || !key)
            throw std::invalid_argument("null argument");
        client_keygen_evk(c, new_sk, old_sk, seed, key, st);
    });
}
int heon_encrypt(heon_context_t ctx, const uint64_t* pk, const uint64_t* pt, uint64_t seed, uint64_t* ct, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!pk || !ct)
            throw std::invalid_argument("null argument");
        client_encrypt(c, pk, pt, seed, ct, st);
    });
}
int heon_ckks_decrypt(heon_context_t ctx, const uint64_t* sk, const uint64_t* ct, int components, int depth,
                      uint64_t* pt, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!sk || !ct || !pt)
            throw std::invalid_argument("null argument");
        if (c.scheme != SCHEME_CKKS)
            throw std::invalid_argument("not a CKKS context");
        client_decrypt_ckks(c, sk, ct, components, depth, pt, st);
    });
}
int heon_bfv_decrypt(heon_context_t ctx, const uint64_t* sk, const uint64_t* ct, int components, uint64_t* pt,
                     void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!sk || !ct || !pt)
            throw std::invalid_argument("null argument");
        if (c.scheme != SCHEME_BFV)
            throw std::invalid_argument("not a BFV context");
        client_decrypt_bfv(c, sk, ct, components, pt, st);
    });
}
int heon_bfv_noise_budget(heon_context_t ctx, const uint64_t* sk, const uint64_t* ct, int components, int* bits,
                          void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!sk || !ct || !bits)
            throw std::invalid_argument("null argument");
        if (c.scheme != SCHEME_BFV)
            throw std::invalid_argument("not a BFV context");
        client_noise_budget_bfv(c, sk, ct, components, bits, st);
    });
}
int heon_ckks_encode(heon_context_t ctx, const double* h_values, int count, double scale, int depth, uint64_t* pt,
                     void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!pt || (count > 0 && !h_values))
            throw std::invalid_argument("null argument");
        if (c.scheme != SCHEME_CKKS)
            throw std::invalid_argument("not a CKKS context");
        client_ckks_encode(c, h_values, count, scale, depth, pt, st);
    });
}
int heon_ckks_decode(heon_context_t ctx, const uint64_t* pt, int depth, double scale, double* h_out, int count,
                     void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!pt || !h_out)
            throw std::invalid_argument("null argument");
        if (c.scheme != SCHEME_CKKS)
            throw std::invalid_argument("not a CKKS context");
        client_ckks_decode(c, pt, depth, scale, h_out, count, st);
    });
}
int heon_bfv_encode(heon_context_t ctx, const uint64_t* h_message, int count, uint64_t* pt, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!pt || (count > 0 && !h_message))
            throw std::invalid_argument("null argument");
        client_bfv_encode(c, h_message, count, pt, st);
    });
}
int heon_bfv_decode(heon_context_t ctx, const uint64_t* pt, uint64_t* h_message, int count, void* stream)
{
    return guarded([&] {
        HEON_CLIENT_PROLOGUE
        if (!pt || !h_message)
            throw std::invalid_argument("null argument");
        client_bfv_decode(c, pt, h_message, count, st);
    });
}

size_t heon_compress_bound(size_t n) { return (size_t) compressBound((uLong) n); }
int heon_compress(const uint8_t* in, size_t n, uint8_t* out, size_t* out_len)
{
    if (!in || !out || !out_len)
        return HEON_ERR_INVALID;
    uLongf len = (uLongf) *out_len;
    if (::compress(out, &len, in, (uLong) n) != Z_OK)
        return HEON_ERR_RUNTIME;
    *out_len = (size_t) len;
    return HEON_OK;
}
int heon_decompress(const uint8_t* in, size_t n, uint8_t* out, size_t* out_len)
{
    if (!in || !out || !out_len)
        return HEON_ERR_INVALID;
    uLongf len = (uLongf) *out_len;
    const int rc = ::uncompress(out, &len, in, (uLong) n);
    if (rc == Z_BUF_ERROR)
        return HEON_ERR_LOGIC; // output buffer too small: the caller retries with a larger one
    if (rc != Z_OK)
        return HEON_ERR_RUNTIME;
    *out_len = (size_t) len;
    return HEON_OK;
}

int heon_profile_begin(void)
{
    profile_begin();
    return HEON_OK;
}
int heon_profile_end(double* ms, long long* launches, int capacity)
{
    if (!ms || !launches || capacity < KC_COUNT)
        return HEON_ERR_INVALID;
    profile_end(ms, launches);
    return KC_COUNT;
}
// ---- TFHE gate bootstrapping (csrc/tfhe.cu) ----
struct heon_tfhe_s {
    TfheContext* c;
};
struct TfheGuard {
    int prev = -1;
    explicit TfheGuard(heon_tfhe_s* h)
    {
        if (!h || !h->c)
            throw std::invalid_argument("null TFHE context");
        cudaGetDevice(&prev);
        cudaSetDevice(tfhe_device(h->c));
        cudaGetLastError();
    }
    ~TfheGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

int heon_tfhe_create(int device, heon_tfhe_t* out)
{
    return guarded([&] {
        if (!out)
            throw std::invalid_argument("null argument");
        *out = nullptr;
        auto h = std::make_unique<heon_tfhe_s>();
        h->c = tfhe_create(device);
        *out = h.release();
    });
}
void heon_tfhe_destroy(heon_tfhe_t h)
{
    if (h)
    {
        tfhe_destroy(h->c);
        delete h;
    }
}
int heon_tfhe_params(heon_tfhe_t h, int* out7)
{
    return guarded([&] {
        if (!h || !h->c || !out7)
            throw std::invalid_argument("null argument");
        tfhe_params(h->c, out7);
    });
}
int heon_tfhe_gate_linear(heon_tfhe_t h, int gate, const int32_t* a1, const int32_t* b1, const int32_t* a2, const int32_t* b2,
                          int32_t* out_a, int32_t* out_b, int n, int shape, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!a1 || !b1 || !out_a || !out_b || (gate != 7 && (!a2 || !b2)))
            throw std::invalid_argument("null argument");
        tfhe_gate_linear(*h->c, gate, a1, b1, a2, b2, out_a, out_b, n, shape, (cudaStream_t) stream);
    });
}
int heon_tfhe_bootstrap(heon_tfhe_t h, const int32_t* in_a, const int32_t* in_b, int32_t* out_a, int32_t* out_b,
                        const uint64_t* boot_key, int shape, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!in_a || !in_b || !out_a || !out_b || !boot_key)
            throw std::invalid_argument("null argument");
        tfhe_bootstrap(*h->c, in_a, in_b, out_a, out_b, (const u64*) boot_key, shape, (cudaStream_t) stream);
    });
}
int heon_tfhe_keyswitch(heon_tfhe_t h, const int32_t* in_a, const int32_t* in_b, int32_t* out_a, int32_t* out_b,
                        const int32_t* ks_a, const int32_t* ks_b, int shape, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!in_a || !in_b || !out_a || !out_b || !ks_a || !ks_b)
            throw std::invalid_argument("null argument");
        tfhe_keyswitch(*h->c, in_a, in_b, out_a, out_b, ks_a, ks_b, shape, (cudaStream_t) stream);
    });
}
int heon_tfhe_gate(heon_tfhe_t h, int gate, const int32_t* a1, const int32_t* b1, const int32_t* a2, const int32_t* b2,
                   const int32_t* a3, const int32_t* b3, int32_t* out_a, int32_t* out_b, const uint64_t* boot_key,
                   const int32_t* ks_a, const int32_t* ks_b, int shape, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!a1 || !b1 || !out_a || !out_b)
            throw std::invalid_argument("null argument");
        if (gate != 7 && (!a2 || !b2 || !boot_key || !ks_a || !ks_b))
            throw std::invalid_argument("null argument");
        tfhe_gate(*h->c, gate, a1, b1, a2, b2, a3, b3, out_a, out_b, (const u64*) boot_key, ks_a, ks_b, shape,
                  (cudaStream_t) stream);
    });
}
int heon_tfhe_keygen_secret(heon_tfhe_t h, uint64_t seed, int32_t* lwe_key, int32_t* tlwe_key, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!lwe_key || !tlwe_key)
            throw std::invalid_argument("null argument");
        tfhe_keygen_secret(*h->c, seed, lwe_key, tlwe_key, (cudaStream_t) stream);
    });
}
int heon_tfhe_keygen_boot(heon_tfhe_t h, const int32_t* lwe_key, const int32_t* tlwe_key, uint64_t seed, uint64_t* boot_key,
                          int32_t* ks_a, int32_t* ks_b, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!lwe_key || !tlwe_key || !boot_key || !ks_a || !ks_b)
            throw std::invalid_argument("null argument");
        tfhe_keygen_boot(*h->c, lwe_key, tlwe_key, seed, (u64*) boot_key, ks_a, ks_b, (cudaStream_t) stream);
    });
}
int heon_tfhe_encrypt(heon_tfhe_t h, const int32_t* lwe_key, const int32_t* d_messages, uint64_t seed, int32_t* out_a,
                      int32_t* out_b, int shape, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!lwe_key || !d_messages || !out_a || !out_b)
            throw std::invalid_argument("null argument");
        tfhe_encrypt(*h->c, lwe_key, d_messages, seed, out_a, out_b, shape, (cudaStream_t) stream);
    });
}
int heon_tfhe_phase(heon_tfhe_t h, const int32_t* lwe_key, const int32_t* in_a, const int32_t* in_b, int32_t* d_phase, int n,
                    int shape, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!lwe_key || !in_a || !in_b || !d_phase)
            throw std::invalid_argument("null argument");
        tfhe_phase(*h->c, lwe_key, in_a, in_b, d_phase, n, shape, (cudaStream_t) stream);
    });
}
int heon_tfhe_ntt(heon_tfhe_t h, uint64_t* data, int count, int inverse, void* stream)
{
    return guarded([&] {
        TfheGuard g(h);
        if (!data)
            throw std::invalid_argument("null argument");
        tfhe_ntt(*h->c, (u64*) data, count, inverse != 0, (cudaStream_t) stream);
    });
}

const char* heon_profile_class_name(int cls)
{
    static const char* names[KC_COUNT] = {"ntt_fwd_col_pass", "ntt_fwd_row_pass", "ntt_inv_row_pass",
                                          "ntt_inv_col_pass", "keyswitch_mac",   "modup_method2",
                                          "moddown",          "cross_multiply",  "elementwise", "keyswitch_row_mac",
                                          "tfhe_blind_rotate", "tfhe_keyswitch"};
    return (cls >= 0 && cls < KC_COUNT) ? names[cls] : "";
}

long long heon_kernel_launches(int reset)
{
    long long v = g_launches.load();
    if (reset)
        g_launches.store(0);
    return v;
}

} // extern "C"
