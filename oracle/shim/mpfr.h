/*
 * Declaration-only stand-in for <mpfr.h> (MPFR 4.2.1 ABI, libmpfr.so.6).
 *
 * TEST INFRASTRUCTURE ONLY.  See shim/gmp.h for why this exists.  Only the entry points
 * the reference's hot path (and our MPFR cross-checks) call are declared.
 */
#ifndef MPRES_ORACLE_SHIM_MPFR_H
#define MPRES_ORACLE_SHIM_MPFR_H

#include "gmp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef long mpfr_prec_t;
typedef int mpfr_sign_t;
typedef long mpfr_exp_t;

typedef struct {
    mpfr_prec_t _mpfr_prec;
    mpfr_sign_t _mpfr_sign;
    mpfr_exp_t _mpfr_exp;
    mp_limb_t *_mpfr_d;
} __mpfr_struct;

typedef __mpfr_struct mpfr_t[1];
typedef __mpfr_struct *mpfr_ptr;
typedef const __mpfr_struct *mpfr_srcptr;

typedef enum {
    MPFR_RNDN = 0, MPFR_RNDZ, MPFR_RNDU, MPFR_RNDD, MPFR_RNDA, MPFR_RNDF, MPFR_RNDNA = -1
} mpfr_rnd_t;

void mpfr_init2(mpfr_ptr, mpfr_prec_t);
void mpfr_init(mpfr_ptr);
void mpfr_clear(mpfr_ptr);
int mpfr_set(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
int mpfr_set_ui(mpfr_ptr, unsigned long, mpfr_rnd_t);
int mpfr_set_si(mpfr_ptr, long, mpfr_rnd_t);
int mpfr_set_z(mpfr_ptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_set_z_2exp(mpfr_ptr, mpz_srcptr, mpfr_exp_t, mpfr_rnd_t);
int mpfr_set_str(mpfr_ptr, const char *, int, mpfr_rnd_t);
char *mpfr_get_str(char *, mpfr_exp_t *, int, size_t, mpfr_srcptr, mpfr_rnd_t);
void mpfr_free_str(char *);
double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t);
double mpfr_get_d_2exp(long *, mpfr_srcptr, mpfr_rnd_t);
long mpfr_get_si(mpfr_srcptr, mpfr_rnd_t);
unsigned long mpfr_get_ui(mpfr_srcptr, mpfr_rnd_t);
mpfr_exp_t mpfr_get_z_2exp(mpz_ptr, mpfr_srcptr);
#define mpfr_get_z_exp mpfr_get_z_2exp
int mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sub(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_div(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_fma(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_mul_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_mul_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_sqrt(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_rec_sqrt(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_log2(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_abs(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_neg(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_pow_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_ui_sub(mpfr_ptr, unsigned long, mpfr_srcptr, mpfr_rnd_t);
int mpfr_ui_div(mpfr_ptr, unsigned long, mpfr_srcptr, mpfr_rnd_t);
int mpfr_cmp3(mpfr_srcptr, mpfr_srcptr, int);
#define mpfr_cmp(a, b) mpfr_cmp3(a, b, 1)
int mpfr_cmp_ui_2exp(mpfr_srcptr, unsigned long, mpfr_exp_t);
#define mpfr_cmp_ui(a, b) mpfr_cmp_ui_2exp(a, b, 0)
int mpfr_zero_p(mpfr_srcptr);
int mpfr_sgn(mpfr_srcptr);
int mpfr_printf(const char *, ...);
int mpfr_sprintf(char *, const char *, ...);

#ifdef __cplusplus
}
#endif
#endif
