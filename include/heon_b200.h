/*
 * heon_b200.h -- C ABI of the B200-native RNS-FHE arithmetic engine.
 *
 * Drop-in boundary for ONE hot path of Alisah-Ozcan/HEonGPU: batched
 * negacyclic NTT/INTT over RNS limbs, RNS base conversion (mod-up, mod-down,
 * rescale) and the key-switch inner product behind multiply -> relinearize
 * (-> rescale) and rotate.  The reference exposes this path as a C++ class
 * layer (no FFI); each entry point below names the reference host function
 * (file:line under the reference tree) whose launch sequence it replaces.
 * The `heongpu::` C++ classes in heongpu_b200/include/heongpu/ forward to
 * these functions with batch = 1.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers to 64-bit words unless the name
 *    starts with h_.  Residues are canonical (0 <= x < p).
 *  - Ciphertext layout is the reference's: [components][L][N] words with
 *    L = Q_size - depth (src/lib/host/ckks/ciphertext.cu:20-31); CKKS
 *    ciphertexts live in the NTT domain.
 *  - Batched calls take a batch count and, per buffer, a batch stride in
 *    words: element b of the batch starts at ptr + b*stride.
 *  - Evaluation keys use the reference layout [digit][2][Q'_0][N]
 *    (src/lib/kernel/keygeneration.cu:180-183), NTT domain.
 *  - `stream` is a cudaStream_t passed as void*; everything is asynchronous
 *    on it.  Scratch memory is stream-ordered (cudaMallocAsync).
 *  - Every function returns HEON_OK (0) or a negative status; the message of
 *    the last failure is available from heon_last_error().  No exceptions
 *    cross the boundary.  (The reference throws std::invalid_argument /
 *    std::logic_error / std::runtime_error; the class layer re-raises them.)
 */
#ifndef HEON_B200_H
#define HEON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct heon_context_s* heon_context_t;

enum {
    HEON_OK = 0,
    HEON_ERR_INVALID = -1, /* std::invalid_argument in the reference */
    HEON_ERR_LOGIC = -2,   /* std::logic_error */
    HEON_ERR_RUNTIME = -3, /* std::runtime_error / CUDA failure */
    HEON_ERR_NO_DEVICE = -4
};

enum { HEON_SCHEME_BFV = 1, HEON_SCHEME_CKKS = 2 };

/* table selectors for heon_context_table() (reference member names) */
enum {
    HEON_TBL_MODULUS = 0,       /* Modulus64 {value,bit,mu} x Q'           */
    HEON_TBL_PSI = 1,           /* minimal primitive 2N-th roots x Q'     */
    HEON_TBL_NTT = 2,           /* ntt_table_   [Q'][N]                    */
    HEON_TBL_INTT = 3,          /* intt_table_  [Q'][N]                    */
    HEON_TBL_N_INVERSE = 4,     /* n_inverse_   [Q']                       */
    HEON_TBL_LAST_Q_MODINV = 5, /* last_q_modinv_                          */
    HEON_TBL_HALF = 6,          /* half_p_                                 */
    HEON_TBL_HALF_MOD = 7,      /* half_mod_                               */
    HEON_TBL_FACTOR = 8,        /* factor_                                 */
    HEON_TBL_RESCALED_LAST_Q_MODINV = 9,
    HEON_TBL_RESCALED_HALF_MOD = 10,
    HEON_TBL_RESCALED_HALF = 11,
    HEON_TBL_II_BASE_CHANGE = 12, /* Method II, per depth (arg `depth`)    */
    HEON_TBL_II_MI_INV = 13,
    HEON_TBL_II_PROD = 14,
    HEON_TBL_II_I_J = 15,      /* int32 widened to u64                     */
    HEON_TBL_II_I_LOCATION = 16,
    /* BFV BEHZ tables (reference member names, src/lib/host/bfv/context.cu:543-660) */
    HEON_TBL_BFV_BASE_CHANGE_BSK = 20,    /* base_change_matrix_Bsk_            [bsk][Q]   */
    HEON_TBL_BFV_INV_PUNCT_Q = 21,        /* inv_punctured_prod_mod_base_array_ [Q]        */
    HEON_TBL_BFV_BASE_CHANGE_MTILDE = 22, /* base_change_matrix_m_tilde_        [Q]        */
    HEON_TBL_BFV_INV_MTILDE_MOD_BSK = 23, /* inv_m_tilde_mod_Bsk_               [bsk]      */
    HEON_TBL_BFV_PROD_Q_MOD_BSK = 24,     /* prod_q_mod_Bsk_                    [bsk]      */
    HEON_TBL_BFV_INV_PROD_Q_MOD_BSK = 25, /* inv_prod_q_mod_Bsk_                [bsk]      */
    HEON_TBL_BFV_BASE_CHANGE_Q = 26,      /* base_change_matrix_q_              [Q][bsk-1] */
    HEON_TBL_BFV_BASE_CHANGE_MSK = 27,    /* base_change_matrix_msk_            [bsk-1]    */
    HEON_TBL_BFV_INV_PUNCT_B = 28,        /* inv_punctured_prod_mod_B_array_    [bsk-1]    */
    HEON_TBL_BFV_PROD_B_MOD_Q = 29,       /* prod_B_mod_q_                      [Q]        */
    HEON_TBL_BFV_SCALARS = 30, /* {inv_prod_q_mod_m_tilde_, inv_prod_B_mod_m_sk_, bsk_modulus, plain_modulus} */
    /* plaintext-operand constants: coeeff_div_plainmod_ [Q], upper_halfincrement_ [Q], Q_mod_t_, upper_threshold_ */
    HEON_TBL_BFV_PLAIN = 31
};

typedef struct heon_info {
    int scheme, n, log_n, q_size, p_size, keyswitch_method, device;
} heon_info;

const char* heon_last_error(void);
const char* heon_version(void);

/* ---- context: HEContext<Scheme::CKKS> ---------------------------------
 * replaces HEContextImpl<CKKS>::set_poly_modulus_degree /
 * set_coeff_modulus_bit_sizes / set_coeff_modulus_values / generate
 * (src/lib/host/ckks/context.cu:26-260,267-539) and the table builders in
 * src/lib/util/util.cu:219-276,356-464,700-767 and
 * src/lib/kernel/contextpool.cpp:11-438.
 * device < 0 builds the host tables only (no CUDA call is made). */
int heon_ckks_context_create(int device, int log_n, const int* q_bits, int n_q, const int* p_bits,
                             int n_p, heon_context_t* out);
int heon_ckks_context_create_values(int device, int log_n, const uint64_t* q, int n_q,
                                    const uint64_t* p, int n_p, heon_context_t* out);
/* ---- context: HEContext<Scheme::BFV> ------------------------------------
 * replaces HEContextImpl<BFV>::generate (src/lib/host/bfv/context.cu:397-700):
 * the Q' chain as for CKKS, plus the BEHZ auxiliary base Bsk (61-bit internal
 * primes, util.cu:278-310) with its NTT tables and conversion constants.
 * HEON_TBL_MODULUS then lists Q' followed by the bsk_modulus Bsk primes. */
int heon_bfv_context_create(int device, int log_n, const int* q_bits, int n_q, const int* p_bits, int n_p,
                            uint64_t plain_modulus, heon_context_t* out);
/* set_coeff_modulus_values / set_coeff_modulus_default_values (bfv/context.cu:223-300): explicit
 * primes, e.g. the default 128-bit-security chains of src/lib/util/defaultmodulus.cpp:12-90. */
int heon_bfv_context_create_values(int device, int log_n, const uint64_t* q, int n_q, const uint64_t* p,
                                   int n_p, uint64_t plain_modulus, heon_context_t* out);
void heon_context_destroy(heon_context_t ctx);
int heon_context_info(heon_context_t ctx, heon_info* out);
/* Copies a host table into h_out (capacity `cap` words); *count receives the
 * table length.  Pass h_out = NULL to query the length. */
int heon_context_table(heon_context_t ctx, int which, int depth, uint64_t* h_out, size_t cap,
                       size_t* count);
/* steps_to_galois_elt (src/lib/kernel/keygeneration.cu:684-727); group_order 5 for CKKS, 3 for BFV */
int heon_steps_to_galois_elt(int steps, int n, int group_order);

/* ---- NTT: gpuntt::GPU_NTT / GPU_INTT and the *_Ordered variants --------
 * (thirdparty/GPU-NTT/src/lib/ntt_merge/ntt.cu:2563-3103,3603-3783,4284-4466)
 * n_polys polynomials of N words, polynomial z uses prime
 * h_prime_index[z % mod_count] (an index into the context's Q' chain; this is
 * `order[z % mod_count]` of GPU_NTT_Modulus_Ordered; pass NULL for the
 * identity 0..mod_count-1 of plain GPU_NTT).  in == out is allowed. */
int heon_ntt(heon_context_t ctx, const uint64_t* in, uint64_t* out, long long n_polys,
             const int* h_prime_index, int mod_count, int inverse, void* stream);
/* GPU_NTT_Poly_Ordered_Inplace: polynomials at word offsets h_offsets[z]
 * from base, all with prime `prime_index`. */
int heon_ntt_poly_ordered(heon_context_t ctx, uint64_t* base, const long long* h_offsets,
                          int n_polys, int prime_index, int inverse, void* stream);

/* ---- element-wise: HEOperator::add / sub / negate ----------------------
 * (src/lib/host/ckks/operator.cu:66-140; kernels src/lib/kernel/addition.cu:10-49) */
int heon_add(heon_context_t ctx, const uint64_t* a, long long a_stride, const uint64_t* b,
             long long b_stride, uint64_t* out, long long out_stride, int components, int depth,
             int batch, void* stream);
int heon_sub(heon_context_t ctx, const uint64_t* a, long long a_stride, const uint64_t* b,
             long long b_stride, uint64_t* out, long long out_stride, int components, int depth,
             int batch, void* stream);
int heon_negate(heon_context_t ctx, const uint64_t* a, long long a_stride, uint64_t* out,
                long long out_stride, int components, int depth, int batch, void* stream);

/* ---- HEOperator<CKKS>::multiply_ckks (operator.cu:796-837) -------------
 * a, b: [2][L][N]; out: [3][L][N]. */
int heon_ckks_multiply(heon_context_t ctx, const uint64_t* a, long long a_stride, const uint64_t* b,
                       long long b_stride, uint64_t* out, long long out_stride, int depth, int batch,
                       void* stream);

/* ---- multiply_plain_ckks (operator.cu:839-871; kernel multiplication.cu:313-331),
 *      add_plain_ckks (:302-345) / sub_plain_ckks (:434-477; kernels addition.cu:175-217).
 * ct, out: [components][L][N]; pt: [L][N] (a CKKS Plaintext lives in the NTT domain).
 * multiply scales every component; add/sub touch component 0 and copy the rest. */
int heon_ckks_multiply_plain(heon_context_t ctx, const uint64_t* ct, long long ct_stride, const uint64_t* pt,
                             long long pt_stride, uint64_t* out, long long out_stride, int components,
                             int depth, int batch, void* stream);
int heon_ckks_add_plain(heon_context_t ctx, const uint64_t* ct, long long ct_stride, const uint64_t* pt,
                        long long pt_stride, uint64_t* out, long long out_stride, int components, int depth,
                        int batch, void* stream);
int heon_ckks_sub_plain(heon_context_t ctx, const uint64_t* ct, long long ct_stride, const uint64_t* pt,
                        long long pt_stride, uint64_t* out, long long out_stride, int components, int depth,
                        int batch, void* stream);

/* ---- switchkey_ckks_method_I / _II (operator.cu:1722-2025): HEOperator::keyswitch.
 * out = (c0, 0) + KeySwitch(c1) under `switch_key` ([digit][2][Q'_0][N]).
 * in, out: [2][L][N], distinct buffers. */
int heon_ckks_keyswitch(heon_context_t ctx, const uint64_t* in, long long in_stride, uint64_t* out,
                        long long out_stride, const uint64_t* switch_key, int depth, int batch, void* stream);

/* ---- conjugate_ckks_method_I / _II (operator.cu:2027-2311): HEOperator::conjugate.
 * apply_galois with galois_elt_zero = 2N-1 and the key's conjugation entry (Galoiskey::c_data()). */
int heon_ckks_conjugate(heon_context_t ctx, const uint64_t* in, long long in_stride, uint64_t* out,
                        long long out_stride, const uint64_t* conjugate_key, int depth, int batch, void* stream);

/* ---- relinearize_seal_method_inplace_ckks (operator.cu:899-1023) and
 *      relinearize_external_product_method2_inplace_ckks (:1025-1154);
 * the method follows the context (P_size == 1 -> I, else II).
 * ct: [3][L][N] in place; on return components 0 and 1 are the result and
 * component 2 holds INTT(c2), exactly as the reference leaves it. */
int heon_ckks_relinearize(heon_context_t ctx, uint64_t* ct, long long ct_stride,
                          const uint64_t* relin_key, int depth, int batch, void* stream);

/* ---- rescale_inplace_ckks_leveled (operator.cu:1156-1244) --------------
 * ct: [2][L][N] -> [2][L-1][N] compacted in place. */
int heon_ckks_rescale(heon_context_t ctx, uint64_t* ct, long long ct_stride, int depth, int batch,
                      void* stream);

/* ---- mod_drop_ckks_leveled[_inplace] (operator.cu:1246-1300) ----------- */
int heon_ckks_mod_drop_inplace(heon_context_t ctx, uint64_t* ct, long long ct_stride, int components,
                               int depth, int batch, void* stream);
int heon_ckks_mod_drop(heon_context_t ctx, const uint64_t* in, long long in_stride, uint64_t* out,
                       long long out_stride, int depth, int batch, void* stream);

/* ---- apply_galois_ckks_method_I / _II (operator.cu:1422-1720) ----------
 * in, out: [2][L][N] (distinct buffers); galois_key: the key for galois_elt
 * ([digit][2][Q'_0][N]).  rotate_rows(shift) = apply_galois with
 * galois_elt = heon_steps_to_galois_elt(shift, N, 5). */
int heon_ckks_apply_galois(heon_context_t ctx, const uint64_t* in, long long in_stride,
                           uint64_t* out, long long out_stride, const uint64_t* galois_key,
                           uint32_t galois_elt, int depth, int batch, void* stream);

/* ---- hoisted rotations: the baby-step loop of the BSGS matrix-vector product,
 *      fast_single_hoisting_rotation_ckks_method_I / _II (operator.cu:4674-4954, 5092-5446).
 * `count` automorphisms of the same ciphertext(s): h_galois_keys[r] (HOST array of DEVICE key
 * pointers) with element h_galois_elts[r] (HOST array).  Rotation r of batch element b lands at
 * out + r*out_rot_stride + b*out_stride, layout [2][L][N]; each equals heon_ckks_apply_galois
 * with the same key bit for bit, but INTT, mod-up and the d*Q' forward NTTs run once. */
int heon_ckks_rotate_hoisted(heon_context_t ctx, const uint64_t* in, long long in_stride, uint64_t* out,
                             long long out_stride, long long out_rot_stride, const uint64_t* const* h_galois_keys,
                             const uint32_t* h_galois_elts, int count, int depth, int batch, void* stream);

/* ---- HEOperator<CKKS>::multiply_matrix_v2 (src/lib/host/ckks/operator.cu:2898-3390), one matrix of
 *      the chain: BSGS diagonal matrix-vector product with DOUBLE hoisting in the PQ_l domain, the
 *      linear-transform core of CKKS bootstrapping (CoeffToSlot / SlotToCoeff).  Method II only, as in
 *      the reference.  in, out: [2][L][N], NTT domain, same depth (the reference applies
 *      rescale_inplace afterwards: heon_ckks_rescale).  PQ_l = limbs {q_0..q_{L-1}, p_0..p_{K-1}}.
 *      The BSGS plan is passed resolved, as the reference host code resolves it (:3176-3193), all
 *      index arrays on the HOST:
 *        baby step i  (n1 of them): Galois element h_baby_elts[i] (0 = no rotation) and DEVICE key
 *                     h_baby_keys[i] (ignored for element 0);
 *        giant step j (n2 of them): element h_giant_elts[j] (0 = none), key h_giant_keys[j],
 *                     h_group_sizes[j] terms; term t (counted over the groups in order) multiplies baby
 *                     step h_term_baby[t] with the diagonal plaintext diags + t*(L+K)*N, layout
 *                     [L+K][N] in the NTT domain over PQ_l (DEVICE). */
int heon_ckks_multiply_matrix(heon_context_t ctx, const uint64_t* in, uint64_t* out, const uint64_t* diags,
                              const uint32_t* h_baby_elts, const uint64_t* const* h_baby_keys, int n1,
                              const uint32_t* h_giant_elts, const uint64_t* const* h_giant_keys,
                              const int* h_group_sizes, const int* h_term_baby, int n2, int depth, void* stream);

/* ---- the giant-step inner sum of the single-hoisting BSGS product (HEOperator<CKKS>::multiply_matrix,
 *      operator.cu:2803-2895): cipherplain_multiply_accumulate_kernel (multiplication.cu:374-403).
 *      out = sum_i cts[i] * pts[i];  cts: [count][2][L][N] (e.g. the output of heon_ckks_rotate_hoisted),
 *      pts: [count][L][N], out: [2][L][N], all in the NTT domain at `depth`. */
int heon_ckks_multiply_plain_accumulate(heon_context_t ctx, const uint64_t* cts, const uint64_t* pts, uint64_t* out,
                                        int count, int depth, void* stream);

/* ---- HEOperator<BFV>::multiply_bfv (src/lib/host/bfv/operator.cu:336-430) --
 * BEHZ multiplication.  a, b: [2][Q][N], out: [3][Q][N], all in the
 * COEFFICIENT domain (BFV ciphertexts are not kept in the NTT domain). */
int heon_bfv_multiply(heon_context_t ctx, const uint64_t* a, long long a_stride, const uint64_t* b,
                      long long b_stride, uint64_t* out, long long out_stride, int batch, void* stream);

/* ---- relinearize_seal_method_inplace (bfv/operator.cu:505-590) and
 *      relinearize_external_product_method2_inplace (:592-671).
 * ct: [3][Q][N] coefficient domain, in place (components 0 and 1 updated). */
int heon_bfv_relinearize(heon_context_t ctx, uint64_t* ct, long long ct_stride, const uint64_t* relin_key,
                         int batch, void* stream);

/* ---- HEOperator<BFV>::apply_galois_method_I / _II (bfv/operator.cu:771-973),
 *      rotate_rows (galois element from heon_steps_to_galois_elt(shift, N, 3)) and
 *      rotate_columns (galois element 2N-1).  in, out: [2][Q][N] coefficient domain. */
int heon_bfv_apply_galois(heon_context_t ctx, const uint64_t* in, long long in_stride, uint64_t* out,
                          long long out_stride, const uint64_t* galois_key, uint32_t galois_elt, int batch,
                          void* stream);

/* ---- add_plain_bfv / sub_plain_bfv (bfv/operator.cu:216-340; kernels addition.cu:50-173) and
 *      multiply_plain_bfv (bfv/operator.cu:432-503; threshold_kernel + cipherplain_kernel).
 * ct, out: [components][Q][N] coefficient domain; pt: [N] values below the plain modulus
 * (pt_stride 0 shares one plaintext across the batch).  multiply: 2 components, out != ct. */
int heon_bfv_add_plain(heon_context_t ctx, const uint64_t* ct, long long ct_stride, const uint64_t* pt,
                       long long pt_stride, uint64_t* out, long long out_stride, int components, int batch,
                       void* stream);
int heon_bfv_sub_plain(heon_context_t ctx, const uint64_t* ct, long long ct_stride, const uint64_t* pt,
                       long long pt_stride, uint64_t* out, long long out_stride, int components, int batch,
                       void* stream);
int heon_bfv_multiply_plain(heon_context_t ctx, const uint64_t* ct, long long ct_stride, const uint64_t* pt,
                            long long pt_stride, uint64_t* out, long long out_stride, int batch, void* stream);

/* ---- HEOperator<BFV>::switchkey_method_I / _II (bfv/operator.cu:975-1372, HEOperator::keyswitch):
 *      out = (c0, 0) + KeySwitch(c1) under `switch_key`; in, out: [2][Q][N] coefficient domain. */
int heon_bfv_keyswitch(heon_context_t ctx, const uint64_t* in, long long in_stride, uint64_t* out,
                       long long out_stride, const uint64_t* switch_key, int batch, void* stream);

/* ---- HOST-resident operands (ExecutionOptions::set_storage_type(storage_type::HOST),
 *      src/include/heongpu/util/storagemanager.cuh:113-167, truth table README.md:349-366).
 * multiply + relinearize_inplace (+ rescale_inplace when `rescale` != 0) on `batch` ciphertext pairs whose
 * words live in HOST memory (pinned for full PCIe speed): h_a, h_b [batch][2][L][N], h_out
 * [batch][2][L - rescale][N].  The batch is processed in chunks of `chunk` ciphertexts (0 = default) with the
 * copies of neighbouring chunks overlapped with the compute; asynchronous, ordered on `stream`.  Calls issued
 * on different streams overlap each other (the copies of one call start while the last chunks of the previous
 * one are still computing / leaving): alternate two streams to keep the link busy across calls. */
int heon_ckks_multiply_relinearize_host(heon_context_t ctx, const uint64_t* h_a, const uint64_t* h_b, uint64_t* h_out,
                                        const uint64_t* relin_key, int depth, int rescale, int batch, int chunk,
                                        void* stream);

/* ---- client side: key generation, encryption, decryption, encoding ------------
 * Secret / public / evaluation keys in the reference's layouts, so they interoperate with every operator
 * above.  Randomness is a counter-based generator seeded by the caller (reproducible); the reference draws
 * from RNGonGPU's AES-CTR DRBG, so key WORDS differ by construction and parity here is decrypt-level.
 *
 * heon_keygen_secret: HEKeyGenerator::generate_secret_key (ckks/keygenerator.cu:28-82; kernels
 *   keygeneration.cu:13-91).  sk: [Q'][N] NTT domain, ternary with `hamming_weight` non-zeros.
 * heon_keygen_public: generate_public_key (ckks/keygenerator.cu:167-243, publickey_gen_kernel :93-116).
 *   pk: [2][Q'][N] = (-(a*s + e), a).
 * heon_keygen_relin: generate_relin_key Method I / II (ckks/keygenerator.cu:245-415; relinkey_gen_kernel
 *   keygeneration.cu:145-185, relinkey_gen_II_kernel :584-629).  key: [d][2][Q'][N].
 * heon_keygen_galois: generate_galois_key (ckks/keygenerator.cu:416-995; galoiskey_gen_kernel :757-860):
 *   the key for apply_galois(galois_elt); heon_keygen_switch: generate_switch_key (:996-1200,
 *   switchkey_gen_kernel :896-1030): re-encrypts a ciphertext under old_sk to new_sk. */
int heon_keygen_secret(heon_context_t ctx, uint64_t seed, int hamming_weight, uint64_t* sk, void* stream);
int heon_keygen_public(heon_context_t ctx, const uint64_t* sk, uint64_t seed, uint64_t* pk, void* stream);
int heon_keygen_relin(heon_context_t ctx, const uint64_t* sk, uint64_t seed, uint64_t* key, void* stream);
int heon_keygen_galois(heon_context_t ctx, const uint64_t* sk, uint32_t galois_elt, uint64_t seed, uint64_t* key,
                       void* stream);
int heon_keygen_switch(heon_context_t ctx, const uint64_t* new_sk, const uint64_t* old_sk, uint64_t seed,
                       uint64_t* key, void* stream);
/* HEEncryptor::encrypt (host/{ckks,bfv}/encryptor.cu; kernels encryption.cu): ct = round((pk*u + e)/P) + pt.
 * CKKS: pt [Q][N] NTT domain, ct [2][Q][N] NTT domain.  BFV: pt [N] below the plain modulus, ct
 * [2][Q][N] coefficient domain.  pt == NULL encrypts zero. */
int heon_encrypt(heon_context_t ctx, const uint64_t* pk, const uint64_t* pt, uint64_t seed, uint64_t* ct,
                 void* stream);
/* HEDecryptor::decrypt (host/ckks/decryptor.cu; sk_multiplication_ckks decryption.cu:349-370):
 * pt [L][N] NTT domain = c0 + c1*s (+ c2*s^2) at `depth`. */
int heon_ckks_decrypt(heon_context_t ctx, const uint64_t* sk, const uint64_t* ct, int components, int depth,
                      uint64_t* pt, void* stream);
/* HEDecryptor<BFV>::decrypt (host/bfv/decryptor.cu): pt [N] = round(t/Q * [c0 + c1*s]_Q) mod t.  The
 * final scaling runs on the host (exact CRT fraction); the call synchronises the stream. */
int heon_bfv_decrypt(heon_context_t ctx, const uint64_t* sk, const uint64_t* ct, int components, uint64_t* pt,
                     void* stream);
/* HEDecryptor<BFV>::remainder_noise_budget (host/bfv/decryptor.cu): -log2(2 * |v|_inf) of the invariant
 * noise v = t/Q * [ct(s)]_Q - m, in bits, from the same host CRT fraction as heon_bfv_decrypt (80-bit long
 * double: budgets above ~58 bits saturate). */
int heon_bfv_noise_budget(heon_context_t ctx, const uint64_t* sk, const uint64_t* ct, int components, int* bits,
                          void* stream);
/* HEEncoder<CKKS>::encode / decode (host/ckks/encoder.cu; kernels encoding.cu:43-400): canonical
 * embedding with the 5^j slot order.  h_values / h_out: `count` complex slots as (re, im) pairs on the
 * HOST; pt: [L][N] NTT domain on the device.  The FFT and the CRT composition run on the host. */
int heon_ckks_encode(heon_context_t ctx, const double* h_values, int count, double scale, int depth, uint64_t* pt,
                     void* stream);
int heon_ckks_decode(heon_context_t ctx, const uint64_t* pt, int depth, double scale, double* h_out, int count,
                     void* stream);
/* HEEncoder<BFV>::encode / decode (host/bfv/encoder.cu; encode_kernel_bfv / decode_kernel_bfv
 * encoding.cu:11-41): batching with generator 3 and the negacyclic transform modulo t. */
int heon_bfv_encode(heon_context_t ctx, const uint64_t* h_message, int count, uint64_t* pt, void* stream);
int heon_bfv_decode(heon_context_t ctx, const uint64_t* pt, uint64_t* h_message, int count, void* stream);

/* ---- serialization helpers (heongpu::serializer, src/lib/util/serializer.cpp:20-50): zlib compress /
 * uncompress of a byte buffer, so that consumers of the header-only class layer need not link zlib.
 * *out_len: capacity on entry, bytes written on return.  heon_compress_bound(n) = a sufficient capacity. */
size_t heon_compress_bound(size_t n);
int heon_compress(const uint8_t* in, size_t n, uint8_t* out, size_t* out_len);
int heon_decompress(const uint8_t* in, size_t n, uint8_t* out, size_t* out_len);

/* ==== TFHE gate bootstrapping (BASELINE config 5) ==================================
 * The reference's fixed parameter set (src/lib/host/tfhe/context.cu:23-56): LWE n = 512, ring N = 1024, k = 1,
 * bootstrapping-key decomposition l = 2 / Bg = 2^10, key-switch base 4 x 8 digits, NTT prime
 * 1152921504606877697.  Ciphertexts are batches of `shape` LWE samples: a = int32[shape][n], b = int32[shape]
 * (Ciphertext<TFHE>::a_device_location_, b_device_location_).  Keys (Bootstrappingkey<TFHE>):
 *   boot_key uint64[n][k+1][l][k+1][N], NTT domain (boot_key_device_location_),
 *   ks_a int32[kN][8][3][n], ks_b int32[kN][8][3] (switch_key_device_location_a_/_b_).
 * All pointers are DEVICE pointers unless named h_*. */
typedef struct heon_tfhe_s* heon_tfhe_t;
int heon_tfhe_create(int device, heon_tfhe_t* out); /* HEContextImpl<TFHE>::HEContextImpl */
void heon_tfhe_destroy(heon_tfhe_t ctx);
/* out7 = {n, N, k, l, bg_bit, ks_base_bit, ks_length} */
int heon_tfhe_params(heon_tfhe_t ctx, int* out7);
/* gate codes */
#define HEON_TFHE_NAND 0
#define HEON_TFHE_AND 1
#define HEON_TFHE_NOR 2
#define HEON_TFHE_OR 3
#define HEON_TFHE_XNOR 4
#define HEON_TFHE_XOR 5
#define HEON_TFHE_ANDNY 6 /* AND with the first input negated (AND_N_pre_computation) */
#define HEON_TFHE_NOT 7
#define HEON_TFHE_MUX 8
/* HELogicOperator<TFHE>::NAND / AND / NOR / OR / XNOR / XOR / NOT / MUX (src/include/heongpu/host/tfhe/
 * operator.cuh:53-812): linear part, blind rotation + sample extraction, key switch.  MUX(in1, in2, control):
 * a3, b3 = control.  NOT needs no keys. */
int heon_tfhe_gate(heon_tfhe_t ctx, int gate, const int32_t* a1, const int32_t* b1, const int32_t* a2, const int32_t* b2,
                   const int32_t* a3, const int32_t* b3, int32_t* out_a, int32_t* out_b, const uint64_t* boot_key,
                   const int32_t* ks_a, const int32_t* ks_b, int shape, void* stream);
/* the three steps of a gate, separately:
 *  *_pre_computation / NOT_computation (tfhe/operator.cu:24-196; bootstrapping.cu:378-660) on samples of n words, */
int heon_tfhe_gate_linear(heon_tfhe_t ctx, int gate, const int32_t* a1, const int32_t* b1, const int32_t* a2, const int32_t* b2,
                          int32_t* out_a, int32_t* out_b, int n, int shape, void* stream);
/*  HELogicOperator<TFHE>::bootstrapping (tfhe/operator.cu:198-266: tfhe_bootstrapping_kernel_unique_step1/2,
 *  _regular_step1/2 x 511, tfhe_sample_extraction_kernel) as ONE launch: out_a int32[shape][kN], out_b int32[shape], */
int heon_tfhe_bootstrap(heon_tfhe_t ctx, const int32_t* in_a, const int32_t* in_b, int32_t* out_a, int32_t* out_b,
                        const uint64_t* boot_key, int shape, void* stream);
/*  HELogicOperator<TFHE>::key_switching (tfhe/operator.cu:268-290; tfhe_key_switching_kernel). */
int heon_tfhe_keyswitch(heon_tfhe_t ctx, const int32_t* in_a, const int32_t* in_b, int32_t* out_a, int32_t* out_b,
                        const int32_t* ks_a, const int32_t* ks_b, int shape, void* stream);
/* client side: HEKeyGenerator<TFHE>::generate_secret_key / generate_bootstrapping_key (tfhe/keygenerator.cu),
 * HEEncryptor<TFHE>::encrypt_lwe_symmetric (d_messages: torus32 words, +-2^29 for true / false),
 * HEDecryptor<TFHE>::decrypt_lwe (phase = b - <a, s>; the bit is phase > 0). */
int heon_tfhe_keygen_secret(heon_tfhe_t ctx, uint64_t seed, int32_t* lwe_key, int32_t* tlwe_key, void* stream);
int heon_tfhe_keygen_boot(heon_tfhe_t ctx, const int32_t* lwe_key, const int32_t* tlwe_key, uint64_t seed, uint64_t* boot_key,
                          int32_t* ks_a, int32_t* ks_b, void* stream);
int heon_tfhe_encrypt(heon_tfhe_t ctx, const int32_t* lwe_key, const int32_t* d_messages, uint64_t seed, int32_t* out_a,
                      int32_t* out_b, int shape, void* stream);
int heon_tfhe_phase(heon_tfhe_t ctx, const int32_t* lwe_key, const int32_t* in_a, const int32_t* in_b, int32_t* d_phase, int n,
                    int shape, void* stream);
/* SmallForwardNTT / SmallInverseNTT (src/lib/kernel/small_ntt.cu) of `count` polynomials of 1024 words, in place */
int heon_tfhe_ntt(heon_tfhe_t ctx, uint64_t* data, int count, int inverse, void* stream);

/* Per-kernel-class CUDA-event profiler (used by bench.py for the roofline
 * line): begin() arms it, end() synchronises the device and returns, per
 * class, the summed device time in ms and the launch count; the return value
 * is the number of classes.  heon_profile_class_name(i) names class i. */
int heon_profile_begin(void);
int heon_profile_end(double* h_ms, long long* h_launches, int capacity);
const char* heon_profile_class_name(int cls);

/* Number of launches of this library's own kernels since the counter was
 * last reset (bench.py reports it as gpu_launches). */
long long heon_kernel_launches(int reset);

#ifdef __cplusplus
}
#endif
#endif /* HEON_B200_H */
