// Reference-kernel harness for TFHE gate bootstrapping.  TEST / BASELINE INFRASTRUCTURE ONLY.
//
// One translation unit that includes the reference's OWN kernel sources where they lie under /root/reference
// (src/lib/kernel/small_ntt.cu and src/lib/kernel/bootstrapping.cu, unmodified; SmallForwardNTT is a __device__
// function defined in the first and used in the second, so they are compiled together instead of with -rdc) and
// replays the launch sequences of src/lib/host/tfhe/operator.cu (:24-196 gate pre-computation, :198-266
// bootstrapping = 1 + 2*511 + 1 launches, :268-290 key switching) and the context constants of
// src/lib/host/tfhe/context.cu:23-104 on caller-provided device buffers.  Used by tests/test_gpu_tfhe.py as the
// bit-exact oracle and by `bench.py --impl reference --workload M5_tfhe_nand` as the timed baseline.
#include REF_SMALL_NTT
#include REF_BOOTSTRAPPING
#include <vector>
#include <cmath>

using namespace heongpu;

struct RefTfhe {
    Modulus64 prime_;
    Root64 *ntt_table_, *intt_table_;
    Ninverse64 n_inverse_;
    int n_, N_, k_, bk_l_, bk_bg_bit_, bg_, half_bg_, mask_mod_, offset_, Npower_;
    int ks_base_bit_, ks_length_;
    int32_t encode_mu;
    Data64* temp_boot = nullptr;
    int32_t* temp_boot2 = nullptr;
    int temp_shape = 0;
};

static std::vector<Root64> compute_ntt_table(Data64 psi, Modulus64 primes, int n_power)
{
    // tfhe/context.cu:80-104
    int n = 1 << n_power;
    std::vector<Root64> forward_table, table;
    table.push_back(1);
    for (int j = 1; j < n; j++)
        table.push_back(OPERATOR64::mult(table[(j - 1)], psi, primes));
    for (int j = 0; j < n; j++)
        forward_table.push_back(table[gpuntt::bitreverse(j, n_power)]);
    return forward_table;
}

static int32_t encode_to_torus32(uint32_t mu, uint32_t m_size)
{
    uint64_t interval = ((1ULL << 63) / m_size) * 2;
    uint64_t phase64 = mu * interval;
    return static_cast<int32_t>(phase64 >> 32);
}

extern "C" {

void* reftfhe_create()
{
    RefTfhe* h = new RefTfhe();
    h->prime_ = Modulus64(1152921504606877697ULL);
    Data64 psi = 1689264667710614ULL;
    Data64 psi_inv = OPERATOR64::modinv(psi, h->prime_);
    std::vector<Root64> f = compute_ntt_table(psi, h->prime_, 10), b = compute_ntt_table(psi_inv, h->prime_, 10);
    cudaMalloc(&h->ntt_table_, sizeof(Root64) * 1024);
    cudaMalloc(&h->intt_table_, sizeof(Root64) * 1024);
    cudaMemcpy(h->ntt_table_, f.data(), sizeof(Root64) * 1024, cudaMemcpyHostToDevice);
    cudaMemcpy(h->intt_table_, b.data(), sizeof(Root64) * 1024, cudaMemcpyHostToDevice);
    h->n_inverse_ = OPERATOR64::modinv(1024, h->prime_);
    h->ks_base_bit_ = 2;
    h->ks_length_ = 8;
    h->n_ = 512;
    h->N_ = 1024;
    h->k_ = 1;
    h->bk_l_ = 2;
    h->bk_bg_bit_ = 10;
    h->bg_ = 1 << h->bk_bg_bit_;
    h->half_bg_ = h->bg_ >> 1;
    h->mask_mod_ = h->bg_ - 1;
    int64_t sum = 0;
    for (int i = 1; i <= h->bk_l_; ++i)
        sum += static_cast<int64_t>(1) << (32 - i * h->bk_bg_bit_);
    h->offset_ = static_cast<int>(sum * h->half_bg_);
    h->Npower_ = 10;
    h->encode_mu = encode_to_torus32(1, 8);
    return h;
}

void reftfhe_destroy(void* hv)
{
    RefTfhe* h = (RefTfhe*) hv;
    cudaFree(h->ntt_table_);
    cudaFree(h->intt_table_);
    cudaFree(h->temp_boot);
    cudaFree(h->temp_boot2);
    delete h;
}

// *_pre_computation / NOT_computation (tfhe/operator.cu:24-196); gate codes of include/heon_b200.h
int reftfhe_gate_linear(void* hv, int gate, int32_t* a1, int32_t* b1, int32_t* a2, int32_t* b2, int32_t* oa, int32_t* ob, int n,
                        int shape, void* stream_v)
{
    cudaStream_t stream = (cudaStream_t) stream_v;
    int32_t e8 = encode_to_torus32(1, 8), e4 = encode_to_torus32(1, 4);
    switch (gate)
    {
    case 0: tfhe_nand_pre_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, a2, b2, e8, n); break;
    case 1: tfhe_and_pre_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, a2, b2, -e8, n); break;
    case 2: tfhe_nor_pre_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, a2, b2, -e8, n); break;
    case 3: tfhe_or_pre_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, a2, b2, e8, n); break;
    case 4: tfhe_xnor_pre_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, a2, b2, -e4, n); break;
    case 5: tfhe_xor_pre_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, a2, b2, e4, n); break;
    case 6: tfhe_and_first_not_pre_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, a2, b2, -e8, n); break;
    case 7: tfhe_not_comp_kernel<<<shape, 512, 0, stream>>>(oa, ob, a1, b1, n); break;
    default: return -1;
    }
    return (int) cudaGetLastError();
}

// HELogicOperator<TFHE>::bootstrapping (tfhe/operator.cu:198-266)
int reftfhe_bootstrap(void* hv, int32_t* in_a, int32_t* in_b, int32_t* out_a, int32_t* out_b, Data64* boot_key, int shape_,
                      void* stream_v)
{
    RefTfhe* c = (RefTfhe*) hv;
    cudaStream_t stream = (cudaStream_t) stream_v;
    if (c->temp_shape < shape_)
    {
        cudaFree(c->temp_boot);
        cudaFree(c->temp_boot2);
        // the reference draws these two from its stream-ordered pool per call
        cudaMalloc(&c->temp_boot, sizeof(Data64) * (size_t) shape_ * (c->k_ + 1) * (c->bk_l_ + 1) * (c->k_ + 1) * c->N_);
        cudaMalloc(&c->temp_boot2, sizeof(int32_t) * (size_t) shape_ * (c->k_ + 1) * c->N_);
        c->temp_shape = shape_;
    }
    Data64* temp_boot = c->temp_boot;
    int32_t* temp_boot2 = c->temp_boot2;
    tfhe_bootstrapping_kernel_unique_step1<<<dim3(shape_, (c->k_ + 1), c->bk_l_), 512, 0, stream>>>(
        in_a, in_b, temp_boot, boot_key, c->ntt_table_, c->prime_, c->encode_mu, c->offset_, c->mask_mod_, c->half_bg_, c->n_,
        c->N_, c->Npower_, c->k_, c->bk_bg_bit_, c->bk_l_);
    tfhe_bootstrapping_kernel_unique_step2<<<dim3(shape_, (c->k_ + 1)), 512, 0, stream>>>(
        temp_boot, in_b, temp_boot2, c->intt_table_, c->n_inverse_, c->prime_, c->encode_mu, c->n_, c->N_, c->Npower_, c->k_,
        c->bk_l_);
    for (int i = 1; i < c->n_; i++)
    {
        tfhe_bootstrapping_kernel_regular_step1<<<dim3(shape_, (c->k_ + 1), c->bk_l_), 512, 0, stream>>>(
            in_a, in_b, temp_boot2, temp_boot, boot_key, i, c->ntt_table_, c->prime_, c->offset_, c->mask_mod_, c->half_bg_,
            c->n_, c->N_, c->Npower_, c->k_, c->bk_bg_bit_, c->bk_l_);
        tfhe_bootstrapping_kernel_regular_step2<<<dim3(shape_, (c->k_ + 1)), 512, 0, stream>>>(
            temp_boot, temp_boot2, c->intt_table_, c->n_inverse_, c->prime_, c->n_, c->N_, c->k_, c->bk_l_);
    }
    tfhe_sample_extraction_kernel<<<dim3(shape_, c->k_), 512, 0, stream>>>(temp_boot2, out_a, out_b, c->N_, c->k_, 0);
    return (int) cudaGetLastError();
}

// HELogicOperator<TFHE>::key_switching (tfhe/operator.cu:268-290)
int reftfhe_keyswitch(void* hv, int32_t* in_a, int32_t* in_b, int32_t* out_a, int32_t* out_b, int32_t* ks_a, int32_t* ks_b,
                      int shape_, void* stream_v)
{
    RefTfhe* c = (RefTfhe*) hv;
    tfhe_key_switching_kernel<<<shape_, 512, 0, (cudaStream_t) stream_v>>>(in_a, in_b, out_a, out_b, ks_a, ks_b, c->ks_base_bit_,
                                                                        c->ks_length_, c->n_, c->N_, c->k_);
    return (int) cudaGetLastError();
}

// SmallForwardNTT / SmallInverseNTT on `count` polynomials in place (to build NTT-domain test keys)
__global__ void reftfhe_ntt_kernel(Data64* data, const Root64* table, Modulus64 modulus, Ninverse64 ninv, int inverse)
{
    __shared__ Data64 sh[1024];
    Data64* poly = data + (size_t) blockIdx.x * 1024;
    sh[threadIdx.x] = poly[threadIdx.x];
    sh[threadIdx.x + 512] = poly[threadIdx.x + 512];
    __syncthreads();
    if (inverse)
        SmallInverseNTT(sh, table, modulus, ninv, false);
    else
        SmallForwardNTT(sh, table, modulus, false);
    poly[threadIdx.x] = sh[threadIdx.x];
    poly[threadIdx.x + 512] = sh[threadIdx.x + 512];
}
int reftfhe_ntt(void* hv, Data64* data, int count, int inverse, void* stream_v)
{
    RefTfhe* c = (RefTfhe*) hv;
    reftfhe_ntt_kernel<<<count, 512, 0, (cudaStream_t) stream_v>>>(data, inverse ? c->intt_table_ : c->ntt_table_, c->prime_,
                                                                 c->n_inverse_, inverse);
    return (int) cudaGetLastError();
}

} // extern "C"
