/* Minimal stand-in for <gsl/gsl_rstat.h> — TEST INFRASTRUCTURE ONLY (oracle).
 * Call sites: RVNode.cpp:18,58,72,117,122,127; Multinomial.cpp:104; Beta.cpp:100;
 * ModelBase.cpp:64-95,146-157,167-180,192-205.  Restates GSL 2.x rstat/rstat.c:
 * Welford mean / M2 (M3, M4 kept for fidelity of the update order; the P^2 median
 * estimator, min and max are omitted because nothing in the reference reads them). */
#ifndef GBN_SHIM_GSL_RSTAT_H
#define GBN_SHIM_GSL_RSTAT_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    double min, max, mean, M2, M3, M4;
    size_t n;
} gsl_rstat_workspace;
gsl_rstat_workspace *gsl_rstat_alloc(void);
void gsl_rstat_free(gsl_rstat_workspace *w);
size_t gsl_rstat_n(const gsl_rstat_workspace *w);
int gsl_rstat_add(const double x, gsl_rstat_workspace *w);
double gsl_rstat_mean(const gsl_rstat_workspace *w);
double gsl_rstat_variance(const gsl_rstat_workspace *w);
double gsl_rstat_sd(const gsl_rstat_workspace *w);
int gsl_rstat_reset(gsl_rstat_workspace *w);
#ifdef __cplusplus
}
#endif
#endif
