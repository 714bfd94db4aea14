// Reference-kernel harness.  TEST / BASELINE INFRASTRUCTURE ONLY.
//
// Links the reference's OWN CUDA kernels, compiled unmodified for sm_100a from
// /root/reference (thirdparty/GPU-NTT ntt.cu, src/lib/kernel/{switchkey,
// multiplication,addition}.cu -- see oracle/Makefile), and replays the launch
// sequences of src/lib/host/ckks/operator.cu (multiply :796-837, relinearize
// I :899-1023, II :1025-1154, rescale :1156-1244, apply_galois I :1422-1559,
// II :1561-1720) on caller-provided device buffers.  The full reference
// library cannot be built offline (RMM, GMP headers, NTL, GoogleTest), so the
// host classes are replaced by this ~300-line driver; every kernel, launch
// shape and table layout is the reference's.
//
// Used (a) by the -m gpu parity tests as the bit-exact oracle and (b) by
// `bench.py --impl reference` as the timed baseline.  Temporaries are
// allocated once per handle (the reference takes them from a stream-ordered
// pool, so allocation is not part of its steady-state cost either).
#include <heongpu/kernel/switchkey.cuh>
#include <heongpu/kernel/multiplication.cuh>
#include <heongpu/kernel/addition.cuh>
#include "gpuntt/ntt_merge/ntt.cuh"
#include <vector>
#include <cstdio>

using namespace heongpu;

struct RefLevel2 {
    Data64 *bc = nullptr, *mi = nullptr, *pr = nullptr;
    int *Ij = nullptr, *Iloc = nullptr;
    int d = 0;
};

struct RefGpu {
    int n, n_power, Q, K, Qp, method;
    Modulus64* modulus;
    Root64 *ntt_table, *intt_table;
    Ninverse64* n_inverse;
    Data64 *last_q_modinv, *half, *half_mod;
    Data64 *r_modinv, *r_half_mod, *r_half;
    int *new_prime_locations, *new_input_locations, *prime_location_leveled;
    std::vector<RefLevel2> lvl2;
    Data64* temp; // big scratch
    size_t temp_words;
    Data64* bsgs_ws = nullptr; // multiply_matrix_v2 temporaries (the reference draws them from its stream-ordered pool)
    size_t bsgs_words = 0;
};

template <class T> static T* up(const T* h, size_t count)
{
    T* d = nullptr;
    cudaMalloc(&d, sizeof(T) * (count ? count : 1));
    if (count)
        cudaMemcpy(d, h, sizeof(T) * count, cudaMemcpyHostToDevice);
    return d;
}

extern "C" {

void* refgpu_create(int n_power, int Q, int K, const Data64* primes, const Data64* fwd,
                    const Data64* inv, const Data64* ninv, const Data64* last_q_modinv, int n_lqm,
                    const Data64* half, const Data64* half_mod, const Data64* r_modinv,
                    const Data64* r_half_mod, int n_r, const Data64* r_half)
{
    RefGpu* h = new RefGpu();
    h->n_power = n_power;
    h->n = 1 << n_power;
    h->Q = Q;
    h->K = K;
    h->Qp = Q + K;
    h->method = (K == 1) ? 1 : 2;
    std::vector<Modulus64> mods;
    for (int i = 0; i < h->Qp; i++)
        mods.push_back(Modulus64(primes[i]));
    h->modulus = up(mods.data(), mods.size());
    h->ntt_table = up(fwd, (size_t) h->Qp * h->n);
    h->intt_table = up(inv, (size_t) h->Qp * h->n);
    h->n_inverse = up(ninv, h->Qp);
    h->last_q_modinv = up(last_q_modinv, n_lqm);
    h->half = up(half, K);
    h->half_mod = up(half_mod, n_lqm);
    h->r_modinv = up(r_modinv, n_r);
    h->r_half_mod = up(r_half_mod, n_r);
    h->r_half = up(r_half, Q > 1 ? Q - 1 : 0);

    // index tables: ckks/operator.cu:24-56, ckks/context.cu:423-440
    std::vector<int> prime_loc, input_loc, prime_loc_lvl;
    int counter = Q;
    for (int i = 0; i < Q; i++)
    {
        for (int j = 0; j < counter; j++)
            prime_loc.push_back(j);
        counter--;
        for (int j = 0; j < K; j++)
            prime_loc.push_back(Q + j);
    }
    counter = h->Qp;
    for (int i = 0; i < h->Qp - 1; i++)
    {
        int sum = counter - 1;
        for (int j = 0; j < 2; j++)
        {
            input_loc.push_back(sum);
            sum += counter;
        }
        counter--;
    }
    counter = Q;
    for (int i = 0; i < Q - 1; i++)
    {
        for (int j = 0; j < counter; j++)
            prime_loc_lvl.push_back(j);
        counter--;
        for (int j = 0; j < K; j++)
            prime_loc_lvl.push_back(Q + j);
    }
    h->new_prime_locations = up(prime_loc.data(), prime_loc.size());
    h->new_input_locations = up(input_loc.data(), input_loc.size());
    h->prime_location_leveled = up(prime_loc_lvl.data(), prime_loc_lvl.size());
    h->lvl2.resize(Q);

    // temp sizing as in the reference (depth-0 sizes, operator.cu:924-930,1432-1441)
    size_t n = h->n;
    h->temp_words = 4 * n * Q + n * Q * h->Qp + 2 * n * h->Qp + 4 * n * h->Qp;
    cudaMalloc(&h->temp, sizeof(Data64) * h->temp_words);
    return h;
}

void refgpu_set_method2(void* hv, int depth, const Data64* bc, int nbc, const Data64* mi, int nmi,
                        const Data64* pr, int npr, const int* Ij, const int* Iloc, int d)
{
    RefGpu* h = (RefGpu*) hv;
    RefLevel2& l = h->lvl2[depth];
    l.bc = up(bc, nbc);
    l.mi = up(mi, nmi);
    l.pr = up(pr, npr);
    l.Ij = up(Ij, d);
    l.Iloc = up(Iloc, d);
    l.d = d;
}

void refgpu_destroy(void* hv)
{
    RefGpu* h = (RefGpu*) hv;
    cudaFree(h->modulus);
    cudaFree(h->ntt_table);
    cudaFree(h->intt_table);
    cudaFree(h->n_inverse);
    cudaFree(h->last_q_modinv);
    cudaFree(h->half);
    cudaFree(h->half_mod);
    cudaFree(h->r_modinv);
    cudaFree(h->r_half_mod);
    cudaFree(h->r_half);
    cudaFree(h->new_prime_locations);
    cudaFree(h->new_input_locations);
    cudaFree(h->prime_location_leveled);
    for (auto& l : h->lvl2)
    {
        cudaFree(l.bc);
        cudaFree(l.mi);
        cudaFree(l.pr);
        cudaFree(l.Ij);
        cudaFree(l.Iloc);
    }
    cudaFree(h->temp);
    cudaFree(h->bsgs_ws);
    delete h;
}

static gpuntt::ntt_rns_configuration<Data64> cfg_of(RefGpu* h, bool inverse, Ninverse64* ninv,
                                                    cudaStream_t st)
{
    gpuntt::ntt_rns_configuration<Data64> cfg = {
        .n_power = h->n_power,
        .ntt_type = inverse ? gpuntt::INVERSE : gpuntt::FORWARD,
        .ntt_layout = gpuntt::PerPolynomial,
        .reduction_poly = gpuntt::ReductionPolynomial::X_N_plus,
        .zero_padding = false,
        .mod_inverse = ninv,
        .stream = st};
    return cfg;
}

// plain GPU_NTT_Inplace / GPU_INTT_Inplace over the first mod_count primes
int refgpu_ntt(void* hv, Data64* data, int n_polys, int mod_count, int inverse, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t st = (cudaStream_t) stream;
    if (inverse)
        gpuntt::GPU_INTT_Inplace(data, h->intt_table, h->modulus, cfg_of(h, true, h->n_inverse, st),
                                 n_polys, mod_count);
    else
        gpuntt::GPU_NTT_Inplace(data, h->ntt_table, h->modulus, cfg_of(h, false, nullptr, st),
                                n_polys, mod_count);
    return (int) cudaGetLastError();
}

// GPU_NTT_Modulus_Ordered_Inplace over the levelled limb set of `depth`
int refgpu_ntt_level(void* hv, Data64* data, int n_polys, int depth, int inverse, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t st = (cudaStream_t) stream;
    int counter = h->Qp, location = 0;
    for (int i = 0; i < depth; i++)
    {
        location += counter;
        counter--;
    }
    int Ql = h->Qp - depth;
    gpuntt::GPU_NTT_Modulus_Ordered_Inplace(data, inverse ? h->intt_table : h->ntt_table, h->modulus,
                                            cfg_of(h, inverse != 0, h->n_inverse, st), n_polys, Ql,
                                            h->new_prime_locations + location);
    return (int) cudaGetLastError();
}

// multiply_ckks: operator.cu:796-837
int refgpu_multiply(void* hv, Data64* in1, Data64* in2, Data64* out, int depth, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    int L = h->Q - depth;
    cross_multiplication<<<dim3((h->n >> 8), L, 1), 256, 0, (cudaStream_t) stream>>>(
        in1, in2, out, h->modulus, h->n_power, L);
    return (int) cudaGetLastError();
}

int refgpu_addsub(void* hv, Data64* a, Data64* b, Data64* out, int comps, int depth, int op,
                  void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    int L = h->Q - depth;
    dim3 g((h->n >> 8), L, comps);
    cudaStream_t st = (cudaStream_t) stream;
    if (op == 0)
        addition<<<g, 256, 0, st>>>(a, b, out, h->modulus, h->n_power);
    else if (op == 1)
        substraction<<<g, 256, 0, st>>>(a, b, out, h->modulus, h->n_power);
    else
        negation<<<g, 256, 0, st>>>(a, out, h->modulus, h->n_power);
    return (int) cudaGetLastError();
}

// relinearize_seal_method_inplace_ckks (operator.cu:899-1023) /
// relinearize_external_product_method2_inplace_ckks (:1025-1154)
int refgpu_relinearize(void* hv, Data64* ct, Data64* relin_key, int depth, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t stream_ = (cudaStream_t) stream;
    const int n = h->n, n_power = h->n_power;
    int first_rns_mod_count = h->Qp;
    int current_rns_mod_count = h->Qp - depth;
    int first_decomp_count = h->Q;
    int current_decomp_count = h->Q - depth;

    auto cfg_intt = cfg_of(h, true, h->n_inverse, stream_);
    auto cfg_ntt = cfg_of(h, false, nullptr, stream_);

    int counter = first_rns_mod_count, location = 0;
    for (int i = 0; i < depth; i++)
    {
        location += counter;
        counter--;
    }

    gpuntt::GPU_INTT_Inplace(ct + (current_decomp_count << (n_power + 1)), h->intt_table,
                             h->modulus, cfg_intt, current_decomp_count, current_decomp_count);

    Data64* temp1_relin = h->temp;
    Data64* temp2_relin = temp1_relin + ((size_t) n * h->Q * h->Qp);

    if (h->method == 1)
    {
        cipher_broadcast_leveled_kernel<<<dim3((n >> 8), current_decomp_count, 1), 256, 0,
                                          stream_>>>(ct + (current_decomp_count << (n_power + 1)),
                                                     temp1_relin, h->modulus, first_rns_mod_count,
                                                     current_rns_mod_count, n_power);
        gpuntt::GPU_NTT_Modulus_Ordered_Inplace(temp1_relin, h->ntt_table, h->modulus, cfg_ntt,
                                                current_decomp_count * current_rns_mod_count,
                                                current_rns_mod_count,
                                                h->new_prime_locations + location);
        int iteration_count_1 = current_decomp_count / 4;
        int iteration_count_2 = current_decomp_count % 4;
        keyswitch_multiply_accumulate_leveled_kernel<<<dim3((n >> 8), current_rns_mod_count, 1),
                                                       256, 0, stream_>>>(
            temp1_relin, relin_key, temp2_relin, h->modulus, first_rns_mod_count,
            current_decomp_count, iteration_count_1, iteration_count_2, n_power);

        auto cfg_intt2 = cfg_of(h, true, h->n_inverse + first_decomp_count, stream_);
        gpuntt::GPU_NTT_Poly_Ordered_Inplace(temp2_relin,
                                             h->intt_table + ((size_t) first_decomp_count << n_power),
                                             h->modulus + first_decomp_count, cfg_intt2, 2, 1,
                                             h->new_input_locations + (depth * 2));

        divide_round_lastq_leveled_stage_one_kernel<<<dim3((n >> 8), 2, 1), 256, 0, stream_>>>(
            temp2_relin, temp1_relin, h->modulus, h->half, h->half_mod, n_power, first_decomp_count,
            current_decomp_count);

        gpuntt::GPU_NTT_Inplace(temp1_relin, h->ntt_table, h->modulus, cfg_ntt,
                                2 * current_decomp_count, current_decomp_count);

        divide_round_lastq_leveled_stage_two_kernel<<<dim3((n >> 8), current_decomp_count, 2), 256,
                                                      0, stream_>>>(
            temp1_relin, temp2_relin, ct, ct, h->modulus, h->last_q_modinv, n_power,
            current_decomp_count);
    }
    else
    {
        RefLevel2& l = h->lvl2[depth];
        base_conversion_DtoQtilde_relin_leveled_kernel<<<dim3((n >> 8), l.d, 1), 256, 0, stream_>>>(
            ct + (current_decomp_count << (n_power + 1)), temp1_relin, h->modulus, l.bc, l.mi, l.pr,
            l.Ij, l.Iloc, n_power, l.d, current_rns_mod_count, current_decomp_count, depth,
            h->prime_location_leveled + location);
        gpuntt::GPU_NTT_Modulus_Ordered_Inplace(temp1_relin, h->ntt_table, h->modulus, cfg_ntt,
                                                l.d * current_rns_mod_count, current_rns_mod_count,
                                                h->new_prime_locations + location);
        int iteration_count_1 = l.d / 4;
        int iteration_count_2 = l.d % 4;
        keyswitch_multiply_accumulate_leveled_method_II_kernel<<<
            dim3((n >> 8), current_rns_mod_count, 1), 256, 0, stream_>>>(
            temp1_relin, relin_key, temp2_relin, h->modulus, first_rns_mod_count,
            current_decomp_count, current_rns_mod_count, iteration_count_1, iteration_count_2, depth,
            n_power);
        gpuntt::GPU_NTT_Modulus_Ordered_Inplace(temp2_relin, h->intt_table, h->modulus, cfg_intt,
                                                2 * current_rns_mod_count, current_rns_mod_count,
                                                h->new_prime_locations + location);
        divide_round_lastq_extended_leveled_kernel<<<dim3((n >> 8), current_decomp_count, 2), 256,
                                                     0, stream_>>>(
            temp2_relin, temp1_relin, h->modulus, h->half, h->half_mod, h->last_q_modinv, n_power,
            current_rns_mod_count, current_decomp_count, first_rns_mod_count, first_decomp_count,
            h->K);
        gpuntt::GPU_NTT_Inplace(temp1_relin, h->ntt_table, h->modulus, cfg_ntt,
                                2 * current_decomp_count, current_decomp_count);
        addition<<<dim3((n >> 8), current_decomp_count, 2), 256, 0, stream_>>>(
            temp1_relin, ct, ct, h->modulus, n_power);
    }
    return (int) cudaGetLastError();
}

// multiply_matrix_v2, one matrix of the chain (operator.cu:2898-3390): BSGS diagonal mat-vec with double
// hoisting in PQ_l.  The BSGS plan arrives resolved (Galois elements, key pointers, baby-step index per term)
// exactly as the reference host code resolves it from diags_matrices_bsgs_ / rot_n1_ / rot_n2_ (:3176-3193).
// The trailing rescale_inplace (:3387) is refgpu_rescale.
int refgpu_bsgs_matvec(void* hv, Data64* ct, Data64* out, Data64* matrix, const int* baby_elts, Data64** baby_keys,
                       int n1, const int* giant_elts, Data64** giant_keys, const int* group_sizes, const int* term_baby,
                       int n2, int depth, void* stream_v)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t stream = (cudaStream_t) stream_v;
    const int n = h->n, n_power = h->n_power, Q_size = h->Q, P_size = h->K;
    const int first_rns_mod_count = h->Qp;
    const int current_level = depth;
    const int current_decomp_count = Q_size - current_level;
    const int current_rns_mod_count = h->Qp - current_level;
    const int pql_count = current_decomp_count + P_size;

    std::vector<Data64> primes(h->Qp);
    {
        std::vector<Modulus64> m(h->Qp);
        cudaMemcpy(m.data(), h->modulus, sizeof(Modulus64) * h->Qp, cudaMemcpyDeviceToHost);
        for (int i = 0; i < h->Qp; i++)
            primes[i] = m[i].value;
    }
    std::vector<Modulus64> pq_mod_host;
    for (int j = 0; j < current_decomp_count; j++)
        pq_mod_host.push_back(Modulus64(primes[j]));
    for (int k = 0; k < P_size; k++)
        pq_mod_host.push_back(Modulus64(primes[Q_size + k]));
    Modulus64* pq_modulus_dev = up(pq_mod_host.data(), pq_mod_host.size());
    std::vector<Data64> P_mod_q_host(current_decomp_count);
    for (int j = 0; j < current_decomp_count; j++)
    {
        uint64_t p_mod_qj = 1, qj = primes[j];
        for (int k = 0; k < P_size; k++)
        {
            __uint128_t tmp = (__uint128_t) p_mod_qj * (primes[Q_size + k] % qj);
            p_mod_qj = (uint64_t) (tmp % qj);
        }
        P_mod_q_host[j] = p_mod_qj;
    }
    Data64* P_mod_q_dev = up(P_mod_q_host.data(), P_mod_q_host.size());

    int counter_loc = first_rns_mod_count, location = 0;
    for (int i = 0; i < current_level; i++)
    {
        location += counter_loc;
        counter_loc--;
    }
    auto cfg_intt = cfg_of(h, true, h->n_inverse, stream);
    auto cfg_ntt = cfg_of(h, false, nullptr, stream);
    RefLevel2& l = h->lvl2[current_level];
    const int d_level = l.d;
    const int iteration_count_1 = d_level / 4, iteration_count_2 = d_level % 4;
    const size_t baby_ct_size = (size_t) 2 * pql_count * n;

    const size_t ws_need = (size_t) 2 * n * Q_size + 2 * ((size_t) 2 * n * d_level * current_rns_mod_count) + baby_ct_size * n1 +
                           2 * ((size_t) 2 * n * current_rns_mod_count) + (size_t) pql_count * n + 3 * baby_ct_size +
                           (size_t) current_decomp_count * n;
    if (h->bsgs_words < ws_need)
    {
        cudaFree(h->bsgs_ws);
        cudaMalloc(&h->bsgs_ws, ws_need * sizeof(Data64));
        h->bsgs_words = ws_need;
    }
    Data64* ws_next = h->bsgs_ws;
    auto dmalloc = [&](size_t words) {
        Data64* p = ws_next;
        ws_next += words;
        return p;
    };
    Data64* temp0 = dmalloc((size_t) 2 * n * Q_size);
    Data64* temp3 = dmalloc((size_t) 2 * n * d_level * current_rns_mod_count);
    Data64* baby_results = dmalloc(baby_ct_size * n1);
    Data64* temp4 = dmalloc((size_t) 2 * n * current_rns_mod_count);
    Data64* Pc0 = dmalloc((size_t) pql_count * n);
    Data64* gs_accum = dmalloc(baby_ct_size);
    Data64* u_pql = dmalloc(baby_ct_size);
    Data64* u1_Q = dmalloc((size_t) current_decomp_count * n);
    Data64* temp3_gs = dmalloc((size_t) 2 * n * d_level * current_rns_mod_count);
    Data64* temp4_gs = dmalloc((size_t) 2 * n * current_rns_mod_count);
    Data64* permuted_gs = dmalloc(baby_ct_size);
    int total_terms = 0;
    for (int j = 0; j < n2; j++)
        total_terms += group_sizes[j];
    int* ct_indices_all = up(term_baby, total_terms);

    gpuntt::GPU_INTT(ct, temp0, h->intt_table, h->modulus, cfg_intt, 2 * current_decomp_count, current_decomp_count);
    base_conversion_DtoQtilde_relin_leveled_kernel<<<dim3((n >> 8), d_level, 1), 256, 0, stream>>>(
        temp0 + (current_decomp_count << n_power), temp3, h->modulus, l.bc, l.mi, l.pr, l.Ij, l.Iloc, n_power, d_level,
        current_rns_mod_count, current_decomp_count, current_level, h->prime_location_leveled + location);
    gpuntt::GPU_NTT_Modulus_Ordered_Inplace(temp3, h->ntt_table, h->modulus, cfg_ntt, d_level * current_rns_mod_count,
                                            current_rns_mod_count, h->new_prime_locations + location);
    broadcast_scale_P_kernel<<<dim3((n >> 8), pql_count, 1), 256, 0, stream>>>(ct, Pc0, P_mod_q_dev, pq_modulus_dev, n_power,
                                                                             current_decomp_count, pql_count);
    for (int i = 0; i < n1; i++)
    {
        Data64* baby = baby_results + baby_ct_size * i;
        if (baby_elts[i] == 0)
        {
            cudaMemcpyAsync(baby, Pc0, (size_t) pql_count * n * sizeof(Data64), cudaMemcpyDeviceToDevice, stream);
            broadcast_scale_P_kernel<<<dim3((n >> 8), pql_count, 1), 256, 0, stream>>>(
                ct + ((size_t) current_decomp_count * n), baby + ((size_t) pql_count * n), P_mod_q_dev, pq_modulus_dev, n_power,
                current_decomp_count, pql_count);
            continue;
        }
        keyswitch_multiply_accumulate_leveled_method_II_kernel<<<dim3((n >> 8), current_rns_mod_count, 1), 256, 0, stream>>>(
            temp3, baby_keys[i], temp4, h->modulus, first_rns_mod_count, current_decomp_count, current_rns_mod_count,
            iteration_count_1, iteration_count_2, current_level, n_power);
        addition_pql_kernel<<<dim3((n >> 8), pql_count, 1), 256, 0, stream>>>(temp4, Pc0, temp4, pq_modulus_dev, n_power,
                                                                            pql_count);
        galois_permute_ntt_pql_kernel<<<dim3((n >> 8), pql_count, 2), 256, 0, stream>>>(temp4, baby, baby_elts[i], n_power,
                                                                                      pql_count);
    }
    cudaMemsetAsync(gs_accum, 0, baby_ct_size * sizeof(Data64), stream);
    int counter = 0;
    for (int j = 0; j < n2; j++)
    {
        const int inner_n1 = group_sizes[j];
        cipherplain_multiply_accumulate_indexed_kernel<<<dim3((n >> 8), pql_count, 2), 256, 0, stream>>>(
            baby_results, matrix + (((size_t) counter * pql_count) << n_power), u_pql, pq_modulus_dev, ct_indices_all + counter,
            inner_n1, pql_count, pql_count, n_power);
        counter += inner_n1;
        if (giant_elts[j] == 0)
        {
            addition_pql_kernel<<<dim3((n >> 8), pql_count, 2), 256, 0, stream>>>(gs_accum, u_pql, gs_accum, pq_modulus_dev,
                                                                                n_power, pql_count);
            continue;
        }
        gpuntt::GPU_NTT_Modulus_Ordered_Inplace(u_pql + ((size_t) pql_count * n), h->intt_table, h->modulus, cfg_intt, pql_count,
                                                pql_count, h->new_prime_locations + location);
        divide_round_lastq_extended_leveled_kernel<<<dim3((n >> 8), current_decomp_count, 1), 256, 0, stream>>>(
            u_pql + ((size_t) pql_count * n), u1_Q, h->modulus, h->half, h->half_mod, h->last_q_modinv, n_power, pql_count,
            current_decomp_count, first_rns_mod_count, Q_size, P_size);
        base_conversion_DtoQtilde_relin_leveled_kernel<<<dim3((n >> 8), d_level, 1), 256, 0, stream>>>(
            u1_Q, temp3_gs, h->modulus, l.bc, l.mi, l.pr, l.Ij, l.Iloc, n_power, d_level, current_rns_mod_count,
            current_decomp_count, current_level, h->prime_location_leveled + location);
        gpuntt::GPU_NTT_Modulus_Ordered_Inplace(temp3_gs, h->ntt_table, h->modulus, cfg_ntt, d_level * current_rns_mod_count,
                                                current_rns_mod_count, h->new_prime_locations + location);
        keyswitch_multiply_accumulate_leveled_method_II_kernel<<<dim3((n >> 8), current_rns_mod_count, 1), 256, 0, stream>>>(
            temp3_gs, giant_keys[j], temp4_gs, h->modulus, first_rns_mod_count, current_decomp_count, current_rns_mod_count,
            iteration_count_1, iteration_count_2, current_level, n_power);
        addition_pql_kernel<<<dim3((n >> 8), pql_count, 1), 256, 0, stream>>>(temp4_gs, u_pql, temp4_gs, pq_modulus_dev, n_power,
                                                                            pql_count);
        galois_permute_ntt_pql_kernel<<<dim3((n >> 8), pql_count, 2), 256, 0, stream>>>(temp4_gs, permuted_gs, giant_elts[j],
                                                                                      n_power, pql_count);
        addition_pql_kernel<<<dim3((n >> 8), pql_count, 2), 256, 0, stream>>>(gs_accum, permuted_gs, gs_accum, pq_modulus_dev,
                                                                            n_power, pql_count);
    }
    gpuntt::GPU_NTT_Modulus_Ordered_Inplace(gs_accum, h->intt_table, h->modulus, cfg_intt, 2 * pql_count, pql_count,
                                            h->new_prime_locations + location);
    divide_round_lastq_extended_leveled_kernel<<<dim3((n >> 8), current_decomp_count, 2), 256, 0, stream>>>(
        gs_accum, out, h->modulus, h->half, h->half_mod, h->last_q_modinv, n_power, pql_count, current_decomp_count,
        first_rns_mod_count, Q_size, P_size);
    gpuntt::GPU_NTT_Inplace(out, h->ntt_table, h->modulus, cfg_ntt, 2 * current_decomp_count, current_decomp_count);
    cudaStreamSynchronize(stream);
    int err = (int) cudaGetLastError();
    cudaFree(P_mod_q_dev);
    cudaFree(pq_modulus_dev);
    cudaFree(ct_indices_all);
    return err;
}

// rescale_inplace_ckks_leveled: operator.cu:1156-1244
int refgpu_rescale(void* hv, Data64* ct, int depth, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t stream_ = (cudaStream_t) stream;
    const int n = h->n, n_power = h->n_power;
    int first_decomp_count = h->Q;
    int current_decomp_count = h->Q - depth;

    auto cfg_intt = cfg_of(h, true, h->n_inverse + (current_decomp_count - 1), stream_);
    auto cfg_ntt = cfg_of(h, false, nullptr, stream_);

    int counter = first_decomp_count - 1, location = 0;
    for (int i = 0; i < depth; i++)
    {
        location += counter;
        counter--;
    }
    Data64* temp1_rescale = h->temp;
    Data64* temp2_rescale = temp1_rescale + ((size_t) 2 * n * h->Qp);

    gpuntt::GPU_NTT_Poly_Ordered_Inplace(
        ct, h->intt_table + ((size_t) (current_decomp_count - 1) << n_power),
        h->modulus + (current_decomp_count - 1), cfg_intt, 2, 1,
        h->new_input_locations + ((depth + h->K) * 2));

    divide_round_lastq_leveled_stage_one_kernel<<<dim3((n >> 8), 2, 1), 256, 0, stream_>>>(
        ct, temp1_rescale, h->modulus, h->r_half + depth, h->r_half_mod + location, n_power,
        current_decomp_count - 1, current_decomp_count - 1);

    gpuntt::GPU_NTT_Inplace(temp1_rescale, h->ntt_table, h->modulus, cfg_ntt,
                            2 * (current_decomp_count - 1), (current_decomp_count - 1));

    move_cipher_leveled_kernel<<<dim3((n >> 8), current_decomp_count - 1, 2), 256, 0, stream_>>>(
        ct, temp2_rescale, n_power, current_decomp_count - 1);

    divide_round_lastq_rescale_kernel<<<dim3((n >> 8), current_decomp_count - 1, 2), 256, 0,
                                        stream_>>>(temp1_rescale, temp2_rescale, ct, h->modulus,
                                                   h->r_modinv + location, n_power,
                                                   current_decomp_count - 1);
    return (int) cudaGetLastError();
}

// apply_galois_ckks_method_I (operator.cu:1422-1559) / _II (:1561-1720)
int refgpu_apply_galois(void* hv, Data64* in, Data64* out, Data64* galois_key, int galois_elt,
                        int depth, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t stream_ = (cudaStream_t) stream;
    const int n = h->n, n_power = h->n_power;
    int first_rns_mod_count = h->Qp;
    int current_rns_mod_count = h->Qp - depth;
    int first_decomp_count = h->Q;
    int current_decomp_count = h->Q - depth;

    Data64* temp0_rotation = h->temp;
    Data64* temp1_rotation = temp0_rotation + ((size_t) 2 * n * h->Q);
    Data64* temp2_rotation = temp1_rotation + ((size_t) 2 * n * h->Q);
    Data64* temp3_rotation = temp2_rotation + ((size_t) n * h->Q * h->Qp);

    auto cfg_intt = cfg_of(h, true, h->n_inverse, stream_);
    auto cfg_ntt = cfg_of(h, false, nullptr, stream_);

    gpuntt::GPU_INTT(in, temp0_rotation, h->intt_table, h->modulus, cfg_intt,
                     2 * current_decomp_count, current_decomp_count);

    int counter = first_rns_mod_count, location = 0;
    for (int i = 0; i < depth; i++)
    {
        location += counter;
        counter--;
    }
    int d;
    if (h->method == 1)
    {
        d = current_decomp_count;
        ckks_duplicate_kernel<<<dim3((n >> 8), current_decomp_count, 1), 256, 0, stream_>>>(
            temp0_rotation, temp2_rotation, h->modulus, n_power, first_rns_mod_count,
            current_rns_mod_count, current_decomp_count);
    }
    else
    {
        RefLevel2& l = h->lvl2[depth];
        d = l.d;
        base_conversion_DtoQtilde_relin_leveled_kernel<<<dim3((n >> 8), l.d, 1), 256, 0, stream_>>>(
            temp0_rotation + (current_decomp_count << n_power), temp2_rotation, h->modulus, l.bc,
            l.mi, l.pr, l.Ij, l.Iloc, n_power, l.d, current_rns_mod_count, current_decomp_count,
            depth, h->prime_location_leveled + location);
    }
    gpuntt::GPU_NTT_Modulus_Ordered_Inplace(temp2_rotation, h->ntt_table, h->modulus, cfg_ntt,
                                            d * current_rns_mod_count, current_rns_mod_count,
                                            h->new_prime_locations + location);
    int iteration_count_1 = d / 4;
    int iteration_count_2 = d % 4;
    if (h->method == 1)
        keyswitch_multiply_accumulate_leveled_kernel<<<dim3((n >> 8), current_rns_mod_count, 1),
                                                       256, 0, stream_>>>(
            temp2_rotation, galois_key, temp3_rotation, h->modulus, first_rns_mod_count,
            current_decomp_count, iteration_count_1, iteration_count_2, n_power);
    else
        keyswitch_multiply_accumulate_leveled_method_II_kernel<<<
            dim3((n >> 8), current_rns_mod_count, 1), 256, 0, stream_>>>(
            temp2_rotation, galois_key, temp3_rotation, h->modulus, first_rns_mod_count,
            current_decomp_count, current_rns_mod_count, iteration_count_1, iteration_count_2, depth,
            n_power);

    gpuntt::GPU_NTT_Modulus_Ordered_Inplace(temp3_rotation, h->intt_table, h->modulus, cfg_intt,
                                            2 * current_rns_mod_count, current_rns_mod_count,
                                            h->new_prime_locations + location);

    divide_round_lastq_permute_ckks_kernel<<<dim3((n >> 8), current_decomp_count, 2), 256, 0,
                                             stream_>>>(
        temp3_rotation, temp0_rotation, out, h->modulus, h->half, h->half_mod, h->last_q_modinv,
        galois_elt, n_power, current_rns_mod_count, current_decomp_count, first_rns_mod_count,
        first_decomp_count, h->K);

    gpuntt::GPU_NTT_Inplace(out, h->ntt_table, h->modulus, cfg_ntt, 2 * current_decomp_count,
                            current_decomp_count);
    return (int) cudaGetLastError();
}

// ---------------------------------------------------------------------------
// BFV: multiply_bfv (bfv/operator.cu:336-430) and relinearize (505-671)
// ---------------------------------------------------------------------------
struct RefBfv {
    int n, n_power, Q, m;
    Modulus64 *ibase, *obase, *merged;
    Root64 *ntt, *intt;
    Ninverse64* ninv;
    Modulus64 m_tilde, plain;
    Data64 inv_prod_q_mod_m_tilde, inv_prod_B_mod_m_sk;
    Data64 *bcm_bsk, *inv_punct, *bcm_mt, *inv_mt_bsk, *prod_q_bsk, *inv_prod_q_bsk, *bcm_q, *bcm_msk, *inv_punct_B,
        *prod_B_q;
    Data64* temp;
};

void* refgpu_bfv_create(int n_power, int Q, int m, const Data64* q, const Data64* bsk, Data64 plain_modulus,
                        const Data64* fwd, const Data64* inv, const Data64* ninv, const Data64* bcm_bsk,
                        const Data64* inv_punct, const Data64* bcm_mt, const Data64* inv_mt_bsk,
                        const Data64* prod_q_bsk, const Data64* inv_prod_q_bsk, const Data64* bcm_q,
                        const Data64* bcm_msk, const Data64* inv_punct_B, const Data64* prod_B_q,
                        Data64 inv_prod_q_mod_m_tilde, Data64 inv_prod_B_mod_m_sk)
{
    RefBfv* h = new RefBfv();
    h->n_power = n_power;
    h->n = 1 << n_power;
    h->Q = Q;
    h->m = m;
    std::vector<Modulus64> ib, ob, mg;
    for (int i = 0; i < Q; i++)
        ib.push_back(Modulus64(q[i]));
    for (int i = 0; i < m; i++)
        ob.push_back(Modulus64(bsk[i]));
    mg = ib;
    mg.insert(mg.end(), ob.begin(), ob.end());
    h->ibase = up(ib.data(), ib.size());
    h->obase = up(ob.data(), ob.size());
    h->merged = up(mg.data(), mg.size());
    const int W = Q + m;
    h->ntt = up(fwd, (size_t) W * h->n);
    h->intt = up(inv, (size_t) W * h->n);
    h->ninv = up(ninv, W);
    h->m_tilde = Modulus64(1ULL << 32);
    h->plain = Modulus64(plain_modulus);
    h->inv_prod_q_mod_m_tilde = inv_prod_q_mod_m_tilde;
    h->inv_prod_B_mod_m_sk = inv_prod_B_mod_m_sk;
    h->bcm_bsk = up(bcm_bsk, (size_t) m * Q);
    h->inv_punct = up(inv_punct, Q);
    h->bcm_mt = up(bcm_mt, Q);
    h->inv_mt_bsk = up(inv_mt_bsk, m);
    h->prod_q_bsk = up(prod_q_bsk, m);
    h->inv_prod_q_bsk = up(inv_prod_q_bsk, m);
    h->bcm_q = up(bcm_q, (size_t) Q * (m - 1));
    h->bcm_msk = up(bcm_msk, m - 1);
    h->inv_punct_B = up(inv_punct_B, m - 1);
    h->prod_B_q = up(prod_B_q, Q);
    cudaMalloc(&h->temp, sizeof(Data64) * (size_t) 7 * W * h->n);
    cudaDeviceSetLimit(cudaLimitStackSize, 2048); // bfv/context.cu sets this for the per-thread arrays
    return h;
}

void refgpu_bfv_destroy(void* hv)
{
    RefBfv* h = (RefBfv*) hv;
    cudaFree(h->ibase);
    cudaFree(h->obase);
    cudaFree(h->merged);
    cudaFree(h->ntt);
    cudaFree(h->intt);
    cudaFree(h->ninv);
    cudaFree(h->bcm_bsk);
    cudaFree(h->inv_punct);
    cudaFree(h->bcm_mt);
    cudaFree(h->inv_mt_bsk);
    cudaFree(h->prod_q_bsk);
    cudaFree(h->inv_prod_q_bsk);
    cudaFree(h->bcm_q);
    cudaFree(h->bcm_msk);
    cudaFree(h->inv_punct_B);
    cudaFree(h->prod_B_q);
    cudaFree(h->temp);
    delete h;
}

int refgpu_bfv_multiply(void* hv, Data64* in1, Data64* in2, Data64* out, void* stream)
{
    RefBfv* h = (RefBfv*) hv;
    cudaStream_t stream_ = (cudaStream_t) stream;
    const int n = h->n, n_power = h->n_power, Q = h->Q, m = h->m, W = Q + m;
    Data64* temp1_mul = h->temp;
    Data64* temp2_mul = temp1_mul + ((size_t) 4 * n * W);
    fast_convertion<<<dim3((n >> 8), 4, 1), 256, 0, stream_>>>(
        in1, in2, temp1_mul, h->ibase, h->obase, h->m_tilde, h->inv_prod_q_mod_m_tilde, h->inv_mt_bsk,
        h->prod_q_bsk, h->bcm_bsk, h->bcm_mt, h->inv_punct, n_power, Q, m);
    gpuntt::ntt_rns_configuration<Data64> cfg_ntt = {.n_power = n_power,
                                                    .ntt_type = gpuntt::FORWARD,
                                                    .ntt_layout = gpuntt::PerPolynomial,
                                                    .reduction_poly = gpuntt::ReductionPolynomial::X_N_plus,
                                                    .zero_padding = false,
                                                    .stream = stream_};
    gpuntt::ntt_rns_configuration<Data64> cfg_intt = {.n_power = n_power,
                                                     .ntt_type = gpuntt::INVERSE,
                                                     .ntt_layout = gpuntt::PerPolynomial,
                                                     .reduction_poly = gpuntt::ReductionPolynomial::X_N_plus,
                                                     .zero_padding = false,
                                                     .mod_inverse = h->ninv,
                                                     .stream = stream_};
    gpuntt::GPU_NTT_Inplace(temp1_mul, h->ntt, h->merged, cfg_ntt, W * 4, W);
    cross_multiplication<<<dim3((n >> 8), W, 1), 256, 0, stream_>>>(temp1_mul, temp1_mul + ((size_t) W * 2 * n),
                                                                   temp2_mul, h->merged, n_power, W);
    gpuntt::GPU_INTT_Inplace(temp2_mul, h->intt, h->merged, cfg_intt, 3 * W, W);
    fast_floor<<<dim3((n >> 8), 3, 1), 256, 0, stream_>>>(
        temp2_mul, out, h->ibase, h->obase, h->plain, h->inv_punct, h->bcm_bsk, h->inv_prod_q_bsk, h->inv_punct_B,
        h->bcm_q, h->bcm_msk, h->inv_prod_B_mod_m_sk, h->prod_B_q, n_power, Q, m);
    return (int) cudaGetLastError();
}

// relinearize_seal_method_inplace / relinearize_external_product_method2_inplace on a RefGpu handle
// (the Q' chain tables are the same objects as for CKKS; the un-levelled Method-II tables equal depth 0)
int refgpu_bfv_relinearize(void* hv, Data64* ct, Data64* relin_key, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t stream_ = (cudaStream_t) stream;
    const int n = h->n, n_power = h->n_power, Q = h->Q, Qp = h->Qp;
    Data64* temp1_relin = h->temp;
    Data64* temp2_relin = temp1_relin + ((size_t) n * Q * Qp);
    auto cfg_ntt = cfg_of(h, false, nullptr, stream_);
    auto cfg_intt = cfg_of(h, true, h->n_inverse, stream_);
    int d;
    if (h->method == 1)
    {
        d = Q;
        cipher_broadcast_kernel<<<dim3((n >> 8), Q, 1), 256, 0, stream_>>>(ct + (Q << (n_power + 1)), temp1_relin,
                                                                          h->modulus, n_power, Qp);
    }
    else
    {
        RefLevel2& l = h->lvl2[0];
        d = l.d;
        base_conversion_DtoQtilde_relin_kernel<<<dim3((n >> 8), d, 1), 256, 0, stream_>>>(
            ct + (Q << (n_power + 1)), temp1_relin, h->modulus, l.bc, l.mi, l.pr, l.Ij, l.Iloc, n_power, Q, Qp, d);
    }
    gpuntt::GPU_NTT_Inplace(temp1_relin, h->ntt_table, h->modulus, cfg_ntt, d * Qp, Qp);
    keyswitch_multiply_accumulate_kernel<<<dim3((n >> 8), Qp, 1), 256, 0, stream_>>>(
        temp1_relin, relin_key, temp2_relin, h->modulus, n_power, Qp, d / 4, d % 4);
    gpuntt::GPU_INTT_Inplace(temp2_relin, h->intt_table, h->modulus, cfg_intt, 2 * Qp, Qp);
    if (h->method == 1)
        divide_round_lastq_kernel<<<dim3((n >> 8), Q, 2), 256, 0, stream_>>>(
            temp2_relin, ct, ct, h->modulus, h->half, h->half_mod, h->last_q_modinv, n_power, Q);
    else
        divide_round_lastq_extended_kernel<<<dim3((n >> 8), Q, 2), 256, 0, stream_>>>(
            temp2_relin, ct, ct, h->modulus, h->half, h->half_mod, h->last_q_modinv, n_power, Qp, Q, h->K);
    return (int) cudaGetLastError();
}

// apply_galois_method_I / _II (bfv/operator.cu:771-973) on a RefGpu handle
int refgpu_bfv_apply_galois(void* hv, Data64* in, Data64* out, Data64* galois_key, int galois_elt, void* stream)
{
    RefGpu* h = (RefGpu*) hv;
    cudaStream_t stream_ = (cudaStream_t) stream;
    const int n = h->n, n_power = h->n_power, Q = h->Q, Qp = h->Qp;
    Data64* temp0_rotation = h->temp;
    Data64* temp1_rotation = temp0_rotation + ((size_t) 2 * n * Q);
    Data64* temp2_rotation = temp1_rotation + ((size_t) n * Q * Qp);
    auto cfg_ntt = cfg_of(h, false, nullptr, stream_);
    auto cfg_intt = cfg_of(h, true, h->n_inverse, stream_);
    int d;
    if (h->method == 1)
    {
        d = Q;
        bfv_duplicate_kernel<<<dim3((n >> 8), Q, 2), 256, 0, stream_>>>(in, temp0_rotation, temp1_rotation,
                                                                       h->modulus, n_power, Qp);
    }
    else
    {
        RefLevel2& l = h->lvl2[0];
        d = l.d;
        global_memory_replace_kernel<<<dim3((n >> 8), Q, 1), 256, 0, stream_>>>(in, temp0_rotation, n_power);
        base_conversion_DtoQtilde_relin_kernel<<<dim3((n >> 8), d, 1), 256, 0, stream_>>>(
            in + (Q << n_power), temp1_rotation, h->modulus, l.bc, l.mi, l.pr, l.Ij, l.Iloc, n_power, Q, Qp, d);
    }
    gpuntt::GPU_NTT_Inplace(temp1_rotation, h->ntt_table, h->modulus, cfg_ntt, d * Qp, Qp);
    keyswitch_multiply_accumulate_kernel<<<dim3((n >> 8), Qp, 1), 256, 0, stream_>>>(
        temp1_rotation, galois_key, temp2_rotation, h->modulus, n_power, Qp, d / 4, d % 4);
    gpuntt::GPU_INTT_Inplace(temp2_rotation, h->intt_table, h->modulus, cfg_intt, 2 * Qp, Qp);
    divide_round_lastq_permute_bfv_kernel<<<dim3((n >> 8), Q, 2), 256, 0, stream_>>>(
        temp2_rotation, temp0_rotation, out, h->modulus, h->half, h->half_mod, h->last_q_modinv, galois_elt,
        n_power, Qp, Q, h->K);
    return (int) cudaGetLastError();
}

} // extern "C"
