/*
 * Declaration-only stand-in for <gmp.h> (GMP 6.3.0 ABI, libgmp.so.10).
 *
 * TEST INFRASTRUCTURE ONLY.  This image ships the GMP runtime library but not its
 * development header.  The unmodified reference sources under /root/reference include
 * "gmp.h"; to compile them for oracle/_ref we only need prototypes of the entry points
 * they call, bound to the symbols libgmp.so.10 exports (__gmpz_* / __gmp_*).
 * Struct layouts follow the stable public GMP ABI.  Nothing here is product code.
 */
#ifndef MPRES_ORACLE_SHIM_GMP_H
#define MPRES_ORACLE_SHIM_GMP_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef long mp_exp_t;
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

typedef struct {
    mpz_t _mp_seed;
    int _mp_alg;
    union { void *_mp_lc; } _mp_algdata;
} __gmp_randstate_struct;
typedef __gmp_randstate_struct gmp_randstate_t[1];

#define mpz_init __gmpz_init
void __gmpz_init(mpz_ptr);
#define mpz_init2 __gmpz_init2
void __gmpz_init2(mpz_ptr, mp_bitcnt_t);
#define mpz_clear __gmpz_clear
void __gmpz_clear(mpz_ptr);
#define mpz_set __gmpz_set
void __gmpz_set(mpz_ptr, mpz_srcptr);
#define mpz_set_ui __gmpz_set_ui
void __gmpz_set_ui(mpz_ptr, unsigned long);
#define mpz_set_si __gmpz_set_si
void __gmpz_set_si(mpz_ptr, long);
#define mpz_set_str __gmpz_set_str
int __gmpz_set_str(mpz_ptr, const char *, int);
#define mpz_get_str __gmpz_get_str
char *__gmpz_get_str(char *, int, mpz_srcptr);
#define mpz_get_ui __gmpz_get_ui
unsigned long __gmpz_get_ui(mpz_srcptr);
#define mpz_get_si __gmpz_get_si
long __gmpz_get_si(mpz_srcptr);
#define mpz_add __gmpz_add
void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_add_ui __gmpz_add_ui
void __gmpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_sub __gmpz_sub
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_mul __gmpz_mul
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_mul_ui __gmpz_mul_ui
void __gmpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mul_si __gmpz_mul_si
void __gmpz_mul_si(mpz_ptr, mpz_srcptr, long);
#define mpz_mul_2exp __gmpz_mul_2exp
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_mod __gmpz_mod
void __gmpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_fdiv_q __gmpz_fdiv_q
void __gmpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_fdiv_r_ui __gmpz_fdiv_r_ui
unsigned long __gmpz_fdiv_r_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mod_ui __gmpz_fdiv_r_ui
#define mpz_fdiv_q_ui __gmpz_fdiv_q_ui
unsigned long __gmpz_fdiv_q_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_div_ui __gmpz_fdiv_q_ui
#define mpz_cmp __gmpz_cmp
int __gmpz_cmp(mpz_srcptr, mpz_srcptr);
#define mpz_cmp_ui __gmpz_cmp_ui
int __gmpz_cmp_ui(mpz_srcptr, unsigned long);
#define mpz_cmp_si __gmpz_cmp_si
int __gmpz_cmp_si(mpz_srcptr, long);
#define mpz_sizeinbase __gmpz_sizeinbase
size_t __gmpz_sizeinbase(mpz_srcptr, int);
#define mpz_urandomb __gmpz_urandomb
void __gmpz_urandomb(mpz_ptr, gmp_randstate_t, mp_bitcnt_t);
#define mpz_import __gmpz_import
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
#define mpz_export __gmpz_export
void *__gmpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
#define mpz_neg __gmpz_neg
void __gmpz_neg(mpz_ptr, mpz_srcptr);

#define gmp_randinit_default __gmp_randinit_default
void __gmp_randinit_default(gmp_randstate_t);
#define gmp_randseed_ui __gmp_randseed_ui
void __gmp_randseed_ui(gmp_randstate_t, unsigned long);
#define gmp_randclear __gmp_randclear
void __gmp_randclear(gmp_randstate_t);
#define gmp_printf __gmp_printf
int __gmp_printf(const char *, ...);
#define mp_get_memory_functions __gmp_get_memory_functions
void __gmp_get_memory_functions(void *(**)(size_t), void *(**)(void *, size_t, size_t),
                                void (**)(void *, size_t));

#ifdef __cplusplus
}
#endif
#endif
