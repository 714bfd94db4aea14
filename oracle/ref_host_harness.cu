// Thin C wrapper over the reference's OWN host code (compiled unmodified from
// /root/reference by oracle/Makefile into oracle/_ref/libref_host.so).
// TEST INFRASTRUCTURE ONLY: used to pin the CPU oracle's table generators and
// CPU NTT against the reference in this container (no GPU needed).
#include <heongpu/util/util.cuh>
// the level_* generators are private (friended to the context class); the
// harness only needs to call them, layout is unaffected by the access keyword
#define private public
#include <heongpu/kernel/contextpool.hpp>
#undef private
#include "gpuntt/ntt_merge/ntt_cpu.cuh"
#include <cstring>

using namespace heongpu;

extern "C" {

void ref_modulus(Data64 p, Data64* out3)
{
    Modulus64 m(p);
    out3[0] = m.value;
    out3[1] = m.bit;
    out3[2] = m.mu;
}

Data64 ref_mult(Data64 a, Data64 b, Data64 p)
{
    Modulus64 m(p);
    return OPERATOR64::mult(a, b, m);
}

int ref_generate_primes(int n, const int* bits, int count, Data64* out)
{
    try
    {
        std::vector<int> b(bits, bits + count);
        std::vector<Modulus64> pv = generate_primes((size_t) n, b);
        for (int i = 0; i < count; i++)
            out[i] = pv[i].value;
        return 0;
    }
    catch (...)
    {
        return -1;
    }
}

void ref_ntt_tables(const Data64* primes, int count, int n_power, Data64* psi_out, Data64* fwd,
                    Data64* inv, Data64* ninv)
{
    std::vector<Modulus64> pv;
    for (int i = 0; i < count; i++)
        pv.push_back(Modulus64(primes[i]));
    size_t n = (size_t) 1 << n_power;
    std::vector<Data64> psi = generate_primitive_root_of_unity(n, pv);
    std::vector<Root64> f = generate_ntt_table(psi, pv, n_power);
    std::vector<Root64> g = generate_intt_table(psi, pv, n_power);
    std::vector<Ninverse64> ni = generate_n_inverse(n, pv);
    std::memcpy(psi_out, psi.data(), sizeof(Data64) * count);
    std::memcpy(fwd, f.data(), sizeof(Data64) * f.size());
    std::memcpy(inv, g.data(), sizeof(Data64) * g.size());
    std::memcpy(ninv, ni.data(), sizeof(Data64) * count);
}

int ref_moddown_tables(const Data64* primes, int Qp, int K, int Q, Data64* last_q_modinv,
                       Data64* half, Data64* half_mod, Data64* factor)
{
    std::vector<Modulus64> pv;
    for (int i = 0; i < Qp; i++)
        pv.push_back(Modulus64(primes[i]));
    std::vector<Data64> a = calculate_last_q_modinv(pv, Qp, K);
    std::vector<Data64> h = calculate_half(pv, K);
    std::vector<Data64> hm = calculate_half_mod(pv, h, Qp, K);
    std::vector<Data64> f = calculate_factor(pv, Q, K);
    std::memcpy(last_q_modinv, a.data(), sizeof(Data64) * a.size());
    std::memcpy(half, h.data(), sizeof(Data64) * h.size());
    std::memcpy(half_mod, hm.data(), sizeof(Data64) * hm.size());
    std::memcpy(factor, f.data(), sizeof(Data64) * f.size());
    return (int) a.size();
}

// KeySwitchParameterGenerator, CKKS Method II level tables for one depth.
int ref_method2_tables(int n, const Data64* primes, int Qp, int K, int depth, Data64* base_change,
                       Data64* mi_inv, Data64* prod, int* I_j, int* I_loc, int* counts)
{
    std::vector<Data64> base(primes, primes + Qp);
    KeySwitchParameterGenerator pool(n, base, K, scheme_type::ckks,
                                     keyswitching_type::KEYSWITCHING_METHOD_II);
    auto bc = pool.level_base_change_matrix_D_to_Qtilda();
    auto mi = pool.level_Mi_inv_D_to_Qtilda();
    auto pr = pool.level_prod_D_to_Qtilda();
    auto ij = pool.level_I_j();
    auto il = pool.level_I_location();
    std::memcpy(base_change, bc[depth].data(), sizeof(Data64) * bc[depth].size());
    std::memcpy(mi_inv, mi[depth].data(), sizeof(Data64) * mi[depth].size());
    std::memcpy(prod, pr[depth].data(), sizeof(Data64) * pr[depth].size());
    std::memcpy(I_j, ij[depth].data(), sizeof(int) * ij[depth].size());
    std::memcpy(I_loc, il[depth].data(), sizeof(int) * il[depth].size());
    counts[0] = (int) bc[depth].size();
    counts[1] = (int) mi[depth].size();
    counts[2] = (int) pr[depth].size();
    return pool.level_d_[depth];
}

// gpuntt::NTTCPU<Data64> forward / inverse on one polynomial.
void ref_ntt_cpu(Data64* a, int n_power, Data64 p, Data64 psi, int inverse)
{
    gpuntt::NTTFactors<Data64> factors(Modulus64(p), 0, psi);
    gpuntt::NTTParameters<Data64> params(n_power, factors, gpuntt::ReductionPolynomial::X_N_plus);
    gpuntt::NTTCPU<Data64> gen(params);
    std::vector<Data64> in(a, a + ((size_t) 1 << n_power));
    std::vector<Data64> out = inverse ? gen.intt(in) : gen.ntt(in);
    std::memcpy(a, out.data(), sizeof(Data64) * out.size());
}

}
