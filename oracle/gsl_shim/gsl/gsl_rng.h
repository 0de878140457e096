/* Minimal stand-in for <gsl/gsl_rng.h> — TEST INFRASTRUCTURE ONLY (oracle).
 *
 * GSL is an un-vendored, un-pinned dependency of the reference
 * (reference setup.py:10, libgbnet/Makefile:6, README.md:8-10) and is not
 * installed in this image.  This header declares only the symbols the
 * reference calls (GraphBase.cpp:8-9,18; RVNode.h:11) so that
 * the reference libgbnet sources compile UNMODIFIED.
 *
 * Extension (not in GSL): an injectable draw queue, so that parity tests can
 * feed the reference sampler the exact per-node draws the CUDA path used.
 */
#ifndef GBN_SHIM_GSL_RNG_H
#define GBN_SHIM_GSL_RNG_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { const char *name; } gsl_rng_type;

typedef struct gsl_rng {
    const gsl_rng_type *type;
    /* MT19937 state (Matsumoto & Nishimura 1998, 32-bit outputs) */
    unsigned long mt[624];
    int mti;
    /* injection queues (NULL = draw from MT19937) */
    const double *inj_flat;  size_t n_flat;  size_t i_flat;   /* uniforms for gsl_ran_flat   */
    const double *inj_beta;  size_t n_beta;  size_t i_beta;   /* ready-made Beta(a,b) draws  */
    const double *inj_exp;   size_t n_exp;                    /* Exp(1) draws, indexed by the
                                                                 T update they belong to
                                                                 (= i_beta - 1)              */
    unsigned long long n_uniform_calls;
} gsl_rng;

extern const gsl_rng_type *gsl_rng_mt19937;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long int seed);
void gsl_rng_free(gsl_rng *r);
unsigned long int gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
double gsl_rng_uniform_pos(gsl_rng *r);

/* shim-only */
void gslshim_rng_inject(gsl_rng *r,
                        const double *flat_u, size_t n_flat,
                        const double *beta_v, size_t n_beta,
                        const double *exp_v, size_t n_exp);
void gslshim_rng_clear_injection(gsl_rng *r);

#ifdef __cplusplus
}
#endif
#endif
