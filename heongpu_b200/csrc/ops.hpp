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

} // namespace heon
