/* Minimal declarations of the GMP entry points the reference's util.cu
 * (src/lib/util/util.cu:771-887) and bfv/context.cu (:958-985) use.  The image ships libgmp.so.10 without
 * headers; this shim only declares the ABI so the reference source compiles
 * unmodified.  Test infrastructure only. */
#ifndef HEON_GMP_SHIM_H
#define HEON_GMP_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long mp_limb_t;
typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t* _mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
void __gmpz_init(__mpz_struct*);
void __gmpz_clear(__mpz_struct*);
void __gmpz_set_ui(__mpz_struct*, unsigned long);
void __gmpz_mul_ui(__mpz_struct*, const __mpz_struct*, unsigned long);
void __gmpz_add_ui(__mpz_struct*, const __mpz_struct*, unsigned long);
void __gmpz_fdiv_q_2exp(__mpz_struct*, const __mpz_struct*, unsigned long);
void* __gmpz_export(void*, size_t*, int, size_t, int, size_t, const __mpz_struct*);
/* bfv/context.cu:958-985 (floor(Q/t) mod q_i) */
unsigned long __gmpz_fdiv_q_ui(__mpz_struct*, const __mpz_struct*, unsigned long);
unsigned long __gmpz_fdiv_r_ui(__mpz_struct*, const __mpz_struct*, unsigned long);
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_set_ui __gmpz_set_ui
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_add_ui __gmpz_add_ui
#define mpz_div_2exp __gmpz_fdiv_q_2exp
#define mpz_export __gmpz_export
#define mpz_div_ui __gmpz_fdiv_q_ui
#define mpz_mod_ui __gmpz_fdiv_r_ui
#ifdef __cplusplus
}
#endif
#endif
