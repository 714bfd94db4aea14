// TEST INFRASTRUCTURE ONLY.
// C wrapper over the reference's OWN context classes, HEContextImpl<BFV> and
// HEContextImpl<CKKS> (src/lib/host/bfv/context.cu, src/lib/host/ckks/context.cu),
// compiled UNMODIFIED where they lie under /root/reference into
// oracle/_ref/libref_ctx.so.  The RMM / RNGonGPU / OpenSSL infrastructure those
// sources include is replaced by host-only stand-ins in oracle/shim/heongpu/util
// (a DeviceVector that keeps its words on the host), so generate() runs on the
// CPU and every table it builds can be read back and compared word for word
// with the product's (tests/test_oracle_vs_ref_host.py).  Table codes are the
// HEON_TBL_* numbers of include/heon_b200.h.
#define private public
#include <heongpu/host/bfv/context.cuh>
#include <heongpu/host/ckks/context.cuh>
#undef private
#include <cstring>

using namespace heongpu;

namespace {
template <class V> long long put(const V& v, Data64* out, long long cap)
{
    const long long n = (long long) v.size();
    if (out)
        for (long long i = 0; i < n && i < cap; ++i)
            out[i] = (Data64) v[i];
    return n;
}
long long put_mod(const std::vector<Modulus64>& v, Data64* out, long long cap)
{
    const long long n = 3 * (long long) v.size();
    if (out)
        for (long long i = 0; i < (long long) v.size() && 3 * i + 2 < cap; ++i)
        {
            out[3 * i] = v[i].value;
            out[3 * i + 1] = v[i].bit;
            out[3 * i + 2] = v[i].mu;
        }
    return n;
}
} // namespace

extern "C" {

// ---- BFV ------------------------------------------------------------------
void* refctx_bfv_create(int n, const Data64* q, int nq, const Data64* p, int np, int plain_modulus)
{
    try
    {
        auto* c = new HEContextImpl<Scheme::BFV>(sec_level_type::none);
        c->set_poly_modulus_degree((size_t) n);
        c->set_coeff_modulus_values(std::vector<Data64>(q, q + nq), std::vector<Data64>(p, p + np));
        c->set_plain_modulus(plain_modulus);
        c->generate();
        return c;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "refctx_bfv_create: %s\n", e.what());
        return nullptr;
    }
}

void* refctx_bfv_create_default(int n, int p_size, int plain_modulus)
{
    try
    {
        auto* c = new HEContextImpl<Scheme::BFV>(sec_level_type::sec128);
        c->set_poly_modulus_degree((size_t) n);
        c->set_coeff_modulus_default_values(p_size);
        c->set_plain_modulus(plain_modulus);
        c->generate();
        return c;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "refctx_bfv_create_default: %s\n", e.what());
        return nullptr;
    }
}

void refctx_bfv_destroy(void* h) { delete (HEContextImpl<Scheme::BFV>*) h; }

// returns the word count of table `which` (writes at most `cap` words when out != NULL); -1: unknown
long long refctx_bfv_table(void* h, int which, Data64* out, long long cap)
{
    auto* c = (HEContextImpl<Scheme::BFV>*) h;
    switch (which)
    {
        case 0:
        {
            std::vector<Modulus64> m = c->prime_vector_;
            for (auto& b : *c->base_Bsk_)
                m.push_back(b);
            return put_mod(m, out, cap);
        }
        case 2: return put(*c->ntt_table_, out, cap);
        case 3: return put(*c->intt_table_, out, cap);
        case 4: return put(*c->n_inverse_, out, cap);
        case 5: return put(*c->last_q_modinv_, out, cap);
        case 6: return put(*c->half_p_, out, cap);
        case 7: return put(*c->half_mod_, out, cap);
        case 8: return put(*c->factor_, out, cap);
        case 12: return c->base_change_matrix_D_to_Q_tilda_ ? put(*c->base_change_matrix_D_to_Q_tilda_, out, cap) : 0;
        case 13: return c->Mi_inv_D_to_Q_tilda_ ? put(*c->Mi_inv_D_to_Q_tilda_, out, cap) : 0;
        case 14: return c->prod_D_to_Q_tilda_ ? put(*c->prod_D_to_Q_tilda_, out, cap) : 0;
        case 15: return c->I_j_ ? put(*c->I_j_, out, cap) : 0;
        case 16: return c->I_location_ ? put(*c->I_location_, out, cap) : 0;
        case 20: return put(*c->base_change_matrix_Bsk_, out, cap);
        case 21: return put(*c->inv_punctured_prod_mod_base_array_, out, cap);
        case 22: return put(*c->base_change_matrix_m_tilde_, out, cap);
        case 23: return put(*c->inv_m_tilde_mod_Bsk_, out, cap);
        case 24: return put(*c->prod_q_mod_Bsk_, out, cap);
        case 25: return put(*c->inv_prod_q_mod_Bsk_, out, cap);
        case 26: return put(*c->base_change_matrix_q_, out, cap);
        case 27: return put(*c->base_change_matrix_msk_, out, cap);
        case 28: return put(*c->inv_punctured_prod_mod_B_array_, out, cap);
        case 29: return put(*c->prod_B_mod_q_, out, cap);
        case 30:
        {
            std::vector<Data64> s = {c->inv_prod_q_mod_m_tilde_, c->inv_prod_B_mod_m_sk_, (Data64) c->bsk_modulus,
                                     c->plain_modulus_.value};
            return put(s, out, cap);
        }
        case 31:
        {
            // {Q mod t, upper_threshold, coeff_div_plainmod[Q], upper_halfincrement[Q]}
            std::vector<Data64> s = {c->Q_mod_t_, c->upper_threshold_};
            for (auto v : *c->coeeff_div_plainmod_)
                s.push_back(v);
            for (auto v : *c->upper_halfincrement_)
                s.push_back(v);
            return put(s, out, cap);
        }
        // the merged q||Bsk NTT tables the BEHZ multiply transforms with
        case 40: return put_mod(std::vector<Modulus64>(c->q_Bsk_merge_modulus_->begin(), c->q_Bsk_merge_modulus_->end()), out, cap);
        case 41: return put(*c->q_Bsk_merge_ntt_tables_, out, cap);
        case 42: return put(*c->q_Bsk_merge_intt_tables_, out, cap);
        case 43: return put(*c->q_Bsk_n_inverse_, out, cap);
        case 44:
        {
            std::vector<Data64> s = {(Data64) c->m, (Data64) c->l, (Data64) c->l_tilda, (Data64) c->d};
            return put(s, out, cap);
        }
    }
    return -1;
}

// ---- CKKS -----------------------------------------------------------------
void* refctx_ckks_create(int n, const Data64* q, int nq, const Data64* p, int np)
{
    try
    {
        auto* c = new HEContextImpl<Scheme::CKKS>(sec_level_type::none);
        c->set_poly_modulus_degree((size_t) n);
        c->set_coeff_modulus_values(std::vector<Data64>(q, q + nq), std::vector<Data64>(p, p + np));
        c->generate();
        return c;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "refctx_ckks_create: %s\n", e.what());
        return nullptr;
    }
}

void refctx_ckks_destroy(void* h) { delete (HEContextImpl<Scheme::CKKS>*) h; }

long long refctx_ckks_table(void* h, int which, int depth, Data64* out, long long cap)
{
    auto* c = (HEContextImpl<Scheme::CKKS>*) h;
    switch (which)
    {
        case 0: return put_mod(c->prime_vector_, out, cap);
        case 2: return put(*c->ntt_table_, out, cap);
        case 3: return put(*c->intt_table_, out, cap);
        case 4: return put(*c->n_inverse_, out, cap);
        case 5: return put(*c->last_q_modinv_, out, cap);
        case 6: return put(*c->half_p_, out, cap);
        case 7: return put(*c->half_mod_, out, cap);
        case 8: return put(*c->factor_, out, cap);
        case 9: return put(*c->rescaled_last_q_modinv_, out, cap);
        case 10: return put(*c->rescaled_half_mod_, out, cap);
        case 11: return put(*c->rescaled_half_, out, cap);
        case 12: return c->base_change_matrix_D_to_Qtilda_leveled ? put((*c->base_change_matrix_D_to_Qtilda_leveled)[depth], out, cap) : 0;
        case 13: return c->Mi_inv_D_to_Qtilda_leveled ? put((*c->Mi_inv_D_to_Qtilda_leveled)[depth], out, cap) : 0;
        case 14: return c->prod_D_to_Qtilda_leveled ? put((*c->prod_D_to_Qtilda_leveled)[depth], out, cap) : 0;
        case 15: return c->I_j_leveled ? put((*c->I_j_leveled)[depth], out, cap) : 0;
        case 16: return c->I_location_leveled ? put((*c->I_location_leveled)[depth], out, cap) : 0;
    }
    return -1;
}

}
