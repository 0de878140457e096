/* Minimal stand-in for <gsl/gsl_randist.h> — TEST INFRASTRUCTURE ONLY (oracle).
 * Call sites in the reference: Multinomial.cpp:92 (flat), Beta.cpp:50,56,82,83,87
 * (beta_pdf, beta, exponential), Dirichlet.cpp:53-95 (dead code, link only). */
#ifndef GBN_SHIM_GSL_RANDIST_H
#define GBN_SHIM_GSL_RANDIST_H
#include <stddef.h>
#include <gsl/gsl_rng.h>
#ifdef __cplusplus
extern "C" {
#endif
double gsl_ran_flat(const gsl_rng *r, const double a, const double b);
double gsl_ran_exponential(const gsl_rng *r, const double mu);
double gsl_ran_gamma(const gsl_rng *r, const double a, const double b);
double gsl_ran_beta(const gsl_rng *r, const double a, const double b);
double gsl_ran_beta_pdf(const double x, const double a, const double b);
void gsl_ran_dirichlet(const gsl_rng *r, const size_t K, const double alpha[], double theta[]);
double gsl_ran_dirichlet_pdf(const size_t K, const double alpha[], const double theta[]);
double gsl_ran_dirichlet_lnpdf(const size_t K, const double alpha[], const double theta[]);
#ifdef __cplusplus
}
#endif
#endif
