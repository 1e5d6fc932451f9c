/* Stand-in for <gmp.h> (the image ships libgmp.so.10 but not its headers): declarations of the mpz entry
 * points the reference's prover headers use, written from GMP's documented C ABI (struct layout and the
 * __gmpz_ symbol prefix are part of that ABI).  Test infrastructure only: it lets tests/cpp/ref_contexts.cpp
 * compile the reference's own stage contexts against the CUDA executor; nothing under ligero-prover_b200/
 * includes it. */
#ifndef LGR_TEST_GMP_STANDIN_H
#define LGR_TEST_GMP_STANDIN_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef long mp_size_t;
typedef unsigned long mp_bitcnt_t;

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#define LGR_GMPZ(name) __gmpz_##name

void LGR_GMPZ(init)(mpz_ptr);
void LGR_GMPZ(init2)(mpz_ptr, mp_bitcnt_t);
void LGR_GMPZ(clear)(mpz_ptr);
void LGR_GMPZ(set)(mpz_ptr, mpz_srcptr);
void LGR_GMPZ(set_ui)(mpz_ptr, unsigned long);
void LGR_GMPZ(set_si)(mpz_ptr, long);
int LGR_GMPZ(set_str)(mpz_ptr, const char *, int);
char *LGR_GMPZ(get_str)(char *, int, mpz_srcptr);
unsigned long LGR_GMPZ(get_ui)(mpz_srcptr);
long LGR_GMPZ(get_si)(mpz_srcptr);
void LGR_GMPZ(swap)(mpz_ptr, mpz_ptr);
void LGR_GMPZ(add)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(add_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void LGR_GMPZ(sub)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(sub_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void LGR_GMPZ(ui_sub)(mpz_ptr, unsigned long, mpz_srcptr);
void LGR_GMPZ(mul)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(mul_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void LGR_GMPZ(mul_si)(mpz_ptr, mpz_srcptr, long);
void LGR_GMPZ(mul_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void LGR_GMPZ(neg)(mpz_ptr, mpz_srcptr);
void LGR_GMPZ(abs)(mpz_ptr, mpz_srcptr);
void LGR_GMPZ(tdiv_q)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(tdiv_r)(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long LGR_GMPZ(tdiv_q_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void LGR_GMPZ(fdiv_q)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(fdiv_r)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(fdiv_qr)(mpz_ptr, mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(fdiv_q_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void LGR_GMPZ(fdiv_r_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void LGR_GMPZ(tdiv_q_2exp)(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void LGR_GMPZ(mod)(mpz_ptr, mpz_srcptr, mpz_srcptr);
int LGR_GMPZ(invert)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(powm)(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(powm_ui)(mpz_ptr, mpz_srcptr, unsigned long, mpz_srcptr);
void LGR_GMPZ(pow_ui)(mpz_ptr, mpz_srcptr, unsigned long);
void LGR_GMPZ(ui_pow_ui)(mpz_ptr, unsigned long, unsigned long);
int LGR_GMPZ(cmp)(mpz_srcptr, mpz_srcptr);
int LGR_GMPZ(cmp_ui)(mpz_srcptr, unsigned long);
int LGR_GMPZ(cmp_si)(mpz_srcptr, long);
void LGR_GMPZ(and)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(ior)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(xor)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(com)(mpz_ptr, mpz_srcptr);
int LGR_GMPZ(tstbit)(mpz_srcptr, mp_bitcnt_t);
void LGR_GMPZ(setbit)(mpz_ptr, mp_bitcnt_t);
void LGR_GMPZ(clrbit)(mpz_ptr, mp_bitcnt_t);
size_t LGR_GMPZ(sizeinbase)(mpz_srcptr, int);
void LGR_GMPZ(import)(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *LGR_GMPZ(export)(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
int LGR_GMPZ(fits_ulong_p)(mpz_srcptr);
int LGR_GMPZ(fits_slong_p)(mpz_srcptr);
int LGR_GMPZ(fits_uint_p)(mpz_srcptr);
int LGR_GMPZ(fits_sint_p)(mpz_srcptr);
void LGR_GMPZ(gcd)(mpz_ptr, mpz_srcptr, mpz_srcptr);
void LGR_GMPZ(sqrt)(mpz_ptr, mpz_srcptr);
mp_bitcnt_t LGR_GMPZ(popcount)(mpz_srcptr);
int __gmp_printf(const char *, ...);

#ifdef __cplusplus
}
#endif

#define mpz_init __gmpz_init
#define mpz_init2 __gmpz_init2
#define mpz_clear __gmpz_clear
#define mpz_set __gmpz_set
#define mpz_set_ui __gmpz_set_ui
#define mpz_set_si __gmpz_set_si
#define mpz_set_str __gmpz_set_str
#define mpz_get_str __gmpz_get_str
#define mpz_get_ui __gmpz_get_ui
#define mpz_get_si __gmpz_get_si
#define mpz_swap __gmpz_swap
#define mpz_add __gmpz_add
#define mpz_add_ui __gmpz_add_ui
#define mpz_sub __gmpz_sub
#define mpz_sub_ui __gmpz_sub_ui
#define mpz_ui_sub __gmpz_ui_sub
#define mpz_mul __gmpz_mul
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_mul_si __gmpz_mul_si
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_neg __gmpz_neg
#define mpz_abs __gmpz_abs
#define mpz_tdiv_q __gmpz_tdiv_q
#define mpz_tdiv_r __gmpz_tdiv_r
#define mpz_tdiv_q_ui __gmpz_tdiv_q_ui
#define mpz_fdiv_q __gmpz_fdiv_q
#define mpz_fdiv_r __gmpz_fdiv_r
#define mpz_fdiv_qr __gmpz_fdiv_qr
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_fdiv_r_2exp __gmpz_fdiv_r_2exp
#define mpz_tdiv_q_2exp __gmpz_tdiv_q_2exp
#define mpz_mod __gmpz_mod
#define mpz_invert __gmpz_invert
#define mpz_powm __gmpz_powm
#define mpz_powm_ui __gmpz_powm_ui
#define mpz_pow_ui __gmpz_pow_ui
#define mpz_ui_pow_ui __gmpz_ui_pow_ui
#define mpz_cmp __gmpz_cmp
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_cmp_si __gmpz_cmp_si
#define mpz_and __gmpz_and
#define mpz_ior __gmpz_ior
#define mpz_xor __gmpz_xor
#define mpz_com __gmpz_com
#define mpz_tstbit __gmpz_tstbit
#define mpz_setbit __gmpz_setbit
#define mpz_clrbit __gmpz_clrbit
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_import __gmpz_import
#define mpz_export __gmpz_export
#define mpz_fits_ulong_p __gmpz_fits_ulong_p
#define mpz_fits_slong_p __gmpz_fits_slong_p
#define mpz_fits_uint_p __gmpz_fits_uint_p
#define mpz_fits_sint_p __gmpz_fits_sint_p
#define mpz_gcd __gmpz_gcd
#define mpz_sqrt __gmpz_sqrt
#define mpz_popcount __gmpz_popcount
#define gmp_printf __gmp_printf
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)
#define mpz_odd_p(z) (((z)->_mp_size != 0) & (int)((z)->_mp_d[0] & 1))
#define mpz_even_p(z) (!mpz_odd_p(z))

#endif
