// Operator entry points behind the C ABI (implemented in ckks_ops.cu).
#pragma once
#include <atomic>
#include "heon_internal.hpp"

namespace heon {

extern std::atomic<long long> g_launches;

// Kernel classes for the launch counter / per-kernel CUDA-event profiler.
enum KernelClass {
    KC_NTT_FWD_COL = 0,
    KC_NTT_FWD_ROW,
    KC_NTT_INV_ROW,
    KC_NTT_INV_COL,
    KC_KEYSWITCH_MAC,
    KC_MODUP2,
    KC_MODDOWN,
    KC_CROSS_MULTIPLY,
    KC_ELEMENTWISE,
    KC_ROW_MAC,
    KC_TFHE_BLIND_ROTATE,
    KC_TFHE_KEYSWITCH,
    KC_COUNT
};

// RAII scope around one kernel launch: counts it and, when profiling is on,
// brackets it with CUDA events on the launching stream.
struct LaunchScope {
    int cls;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr;
    LaunchScope(int cls, cudaStream_t st);
    ~LaunchScope();
};
void profile_begin();
void profile_end(double* ms, long long* launches);

void op_add(const Context& c, const u64* a, long long a_bs, const u64* b, long long b_bs, u64* out,
            long long o_bs, int comps, int depth, int batch, int op, cudaStream_t st);
void op_plain(const Context& c, const u64* ct, long long ct_bs, const u64* pt, long long pt_bs, u64* out,
              long long o_bs, int comps, int depth, int batch, int op, cudaStream_t st);
void op_rotate_hoisted(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                       long long out_rs, const u64* const* galois_keys, const unsigned* galois_elts, int count,
                       int depth, int batch, cudaStream_t st);
void op_bsgs_matvec(const Context& c, const u64* in, u64* out, const u64* diags, const unsigned* baby_elts,
                    const u64* const* baby_keys, int n1, const unsigned* giant_elts, const u64* const* giant_keys,
                    const int* group_sizes, const int* term_baby, int n2, int depth, cudaStream_t st);
void op_multiply_plain_accumulate(const Context& c, const u64* cts, const u64* pts, u64* out, int count, int depth,
                                  cudaStream_t st);
void op_keyswitch(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                  const u64* switch_key, int depth, int batch, cudaStream_t st);
void op_multiply(const Context& c, const u64* a, long long a_bs, const u64* b, long long b_bs,
                 u64* out, long long o_bs, int depth, int batch, cudaStream_t st);
void op_relinearize(const Context& c, u64* ct, long long ct_bs, const u64* relin_key, int depth,
                    int batch, cudaStream_t st);
void op_mulrelin_host(const Context& c, const u64* h_a, const u64* h_b, u64* h_out, const u64* relin_key, int depth,
                      int rescale, int batch, int chunk, cudaStream_t st);
void op_rescale(const Context& c, u64* ct, long long ct_bs, int depth, int batch, cudaStream_t st);
void op_mod_drop_inplace(const Context& c, u64* ct, long long ct_bs, int comps, int depth, int batch,
                         cudaStream_t st);
void op_mod_drop(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                 int depth, int batch, cudaStream_t st);
void op_bfv_multiply(const Context& c, const u64* a, long long a_bs, const u64* b, long long b_bs, u64* out,
                     long long o_bs, int batch, cudaStream_t st);
void op_bfv_addsub_plain(const Context& c, const u64* ct, long long ct_bs, const u64* pt, long long pt_bs, u64* out,
                         long long o_bs, int comps, int batch, int op, cudaStream_t st);
void op_bfv_multiply_plain(const Context& c, const u64* ct, long long ct_bs, const u64* pt, long long pt_bs, u64* out,
                           long long o_bs, int batch, cudaStream_t st);
void op_bfv_relinearize(const Context& c, u64* ct, long long ct_bs, const u64* relin_key, int batch,
                        cudaStream_t st);
void op_apply_galois(const Context& c, const u64* in, long long in_bs, u64* out, long long out_bs,
                     const u64* galois_key, unsigned galois_elt, int depth, int batch,
                     cudaStream_t st);


// TFHE gate bootstrapping (csrc/tfhe.cu)
struct TfheContext;
TfheContext* tfhe_create(int device);
void tfhe_destroy(TfheContext* c);
int tfhe_device(const TfheContext* c);
void tfhe_params(const TfheContext* c, int* out7);
void tfhe_gate_linear(const TfheContext& c, int gate, const int* a1, const int* b1, const int* a2, const int* b2, int* oa,
                      int* ob, int n, int shape, cudaStream_t st);
void tfhe_bootstrap(const TfheContext& c, const int* in_a, const int* in_b, int* out_a, int* out_b, const u64* bk, int shape,
                    cudaStream_t st);
void tfhe_keyswitch(const TfheContext& c, const int* in_a, const int* in_b, int* out_a, int* out_b, const int* ks_a,
                    const int* ks_b, int shape, cudaStream_t st);
void tfhe_gate(const TfheContext& c, int gate, const int* a1, const int* b1, const int* a2, const int* b2, const int* a3,
               const int* b3, int* oa, int* ob, const u64* bk, const int* ks_a, const int* ks_b, int shape, cudaStream_t st);
void tfhe_keygen_secret(const TfheContext& c, u64 seed, int* lwe_key, int* tlwe_key, cudaStream_t st);
void tfhe_keygen_boot(const TfheContext& c, const int* lwe_key, const int* tlwe_key, u64 seed, u64* bk, int* ks_a, int* ks_b,
                      cudaStream_t st);
void tfhe_encrypt(const TfheContext& c, const int* lwe_key, const int* d_messages, u64 seed, int* out_a, int* out_b, int shape,
                  cudaStream_t st);
void tfhe_phase(const TfheContext& c, const int* lwe_key, const int* in_a, const int* in_b, int* d_phase, int n, int shape,
                cudaStream_t st);
void tfhe_ntt(const TfheContext& c, u64* data, int count, bool inverse, cudaStream_t st);

} // namespace heon
