/* GSL stand-in for the oracle build — TEST INFRASTRUCTURE ONLY.
 *
 * The reference links GSL (setup.py:10, libgbnet/Makefile:6) without vendoring or pinning
 * it, and GSL is not installed here.  This file restates, from GSL 2.x's published
 * algorithms, exactly the 20 symbols the reference calls:
 *
 *   rng/mt.c            MT19937 (Matsumoto & Nishimura), gsl_rng_uniform = x / 2^32
 *   randist/flat.c      a*(1-u) + b*u
 *   randist/exponential.c  -mu * log1p(-u)
 *   randist/gamma.c     Marsaglia & Tsang (2000) for a >= 1, boost for a < 1
 *   randist/beta.c      X1/(X1+X2) from two gammas; pdf via lgamma/log/log1p
 *   rstat/rstat.c       Welford running mean / M2 / M3 / M4
 *
 * Only gsl_ran_beta_pdf and gsl_rstat_* affect deterministic outputs of the reference;
 * everything else only shapes the random stream, which the reference seeds from time(NULL)
 * (ModelORNOR.cpp:24), so stream-level equality with a real GSL build is neither possible
 * nor needed.  The normal deviate inside the gamma sampler uses Box-Muller instead of GSL's
 * ziggurat for the same reason.
 *
 * Shim-only extension: an injectable draw queue (gslshim_rng_inject) used by parity tests.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_rstat.h>

/* ------------------------------------------------------------------ rng: MT19937 */
static const gsl_rng_type mt19937_type = { "mt19937" };
const gsl_rng_type *gsl_rng_mt19937 = &mt19937_type;

#define MT_N 624
#define MT_M 397
static const unsigned long UPPER_MASK = 0x80000000UL;
static const unsigned long LOWER_MASK = 0x7fffffffUL;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    gsl_rng *r = (gsl_rng *) calloc(1, sizeof(gsl_rng));
    r->type = T;
    gsl_rng_set(r, 0);
    return r;
}

void gsl_rng_set(gsl_rng *r, unsigned long int s)
{
    if (s == 0) s = 4357;              /* GSL's default seed for mt19937 */
    r->mt[0] = s & 0xffffffffUL;
    for (int i = 1; i < MT_N; i++) {
        r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long) i);
        r->mt[i] &= 0xffffffffUL;
    }
    r->mti = MT_N;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

#define MAGIC(y) (((y) & 0x1) ? 0x9908b0dfUL : 0)
unsigned long int gsl_rng_get(gsl_rng *r)
{
    unsigned long k;
    unsigned long *const mt = r->mt;
    if (r->mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            unsigned long y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);
            mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ MAGIC(y);
        }
        for (; kk < MT_N - 1; kk++) {
            unsigned long y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);
            mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ MAGIC(y);
        }
        {
            unsigned long y = (mt[MT_N - 1] & UPPER_MASK) | (mt[0] & LOWER_MASK);
            mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ MAGIC(y);
        }
        r->mti = 0;
    }
    k = mt[r->mti];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680UL;
    k ^= (k << 15) & 0xefc60000UL;
    k ^= (k >> 18);
    r->mti++;
    return k & 0xffffffffUL;
}

double gsl_rng_uniform(gsl_rng *r)
{
    r->n_uniform_calls++;
    return gsl_rng_get(r) / 4294967296.0;
}

double gsl_rng_uniform_pos(gsl_rng *r)
{
    double x;
    do { x = gsl_rng_uniform(r); } while (x == 0);
    return x;
}

void gslshim_rng_inject(gsl_rng *r,
                        const double *flat_u, size_t n_flat,
                        const double *beta_v, size_t n_beta,
                        const double *exp_v, size_t n_exp)
{
    r->inj_flat = flat_u; r->n_flat = n_flat; r->i_flat = 0;
    r->inj_beta = beta_v; r->n_beta = n_beta; r->i_beta = 0;
    r->inj_exp = exp_v;   r->n_exp = n_exp;
}

void gslshim_rng_clear_injection(gsl_rng *r)
{
    gslshim_rng_inject(r, NULL, 0, NULL, 0, NULL, 0);
}

/* ------------------------------------------------------------------ randist */
double gsl_ran_flat(const gsl_rng *cr, const double a, const double b)
{
    gsl_rng *r = (gsl_rng *) cr;
    double u;
    if (r->inj_flat) {
        if (r->i_flat >= r->n_flat) abort();       /* a test bug: queue too short */
        u = r->inj_flat[r->i_flat++];
    } else {
        u = gsl_rng_uniform(r);
    }
    return a * (1 - u) + b * u;                    /* randist/flat.c */
}

double gsl_ran_exponential(const gsl_rng *cr, const double mu)
{
    gsl_rng *r = (gsl_rng *) cr;
    if (r->inj_exp) {
        /* the Exp(1) draw belongs to the T update whose Beta draw was consumed last */
        size_t k = r->i_beta - 1;
        if (r->i_beta == 0 || k >= r->n_exp) abort();
        return mu * r->inj_exp[k];
    }
    double u = gsl_rng_uniform(r);
    return -mu * log1p(-u);                        /* randist/exponential.c */
}

static double ran_gaussian(gsl_rng *r)
{
    /* polar Box-Muller (GSL's gsl_ran_gaussian); one deviate per call */
    double x, y, r2;
    do {
        x = -1 + 2 * gsl_rng_uniform_pos(r);
        y = -1 + 2 * gsl_rng_uniform_pos(r);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return y * sqrt(-2.0 * log(r2) / r2);
}

double gsl_ran_gamma(const gsl_rng *cr, const double a, const double b)
{
    gsl_rng *r = (gsl_rng *) cr;
    /* randist/gamma.c: Marsaglia-Tsang "A Simple Method for generating gamma variables" */
    if (a < 1) {
        double u = gsl_rng_uniform_pos(r);
        return gsl_ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
    }
    {
        double x, v, u;
        double d = a - 1.0 / 3.0;
        double c = (1.0 / 3.0) / sqrt(d);
        while (1) {
            do {
                x = ran_gaussian(r);
                v = 1.0 + c * x;
            } while (v <= 0);
            v = v * v * v;
            u = gsl_rng_uniform_pos(r);
            if (u < 1 - 0.0331 * x * x * x * x) break;
            if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
        }
        return b * d * v;
    }
}

double gsl_ran_beta(const gsl_rng *cr, const double a, const double b)
{
    gsl_rng *r = (gsl_rng *) cr;
    if (r->inj_beta) {
        if (r->i_beta >= r->n_beta) abort();
        return r->inj_beta[r->i_beta++];
    }
    if (a <= 1.0 && b <= 1.0) {
        /* Johnk's algorithm, as randist/beta.c does for small shapes */
        double U, V, X, Y;
        while (1) {
            U = gsl_rng_uniform_pos(r);
            V = gsl_rng_uniform_pos(r);
            X = pow(U, 1.0 / a);
            Y = pow(V, 1.0 / b);
            if ((X + Y) <= 1.0) {
                if (X + Y > 0) return X / (X + Y);
                {
                    double logX = log(U) / a, logY = log(V) / b;
                    double logM = logX > logY ? logX : logY;
                    logX -= logM; logY -= logM;
                    return exp(logX - log(exp(logX) + exp(logY)));
                }
            }
        }
    }
    {
        double x1 = gsl_ran_gamma(r, a, 1.0);
        double x2 = gsl_ran_gamma(r, b, 1.0);
        return x1 / (x1 + x2);
    }
}

double gsl_ran_beta_pdf(const double x, const double a, const double b)
{
    /* randist/beta.c */
    if (x < 0 || x > 1) {
        return 0;
    } else {
        double p;
        double gab = lgamma(a + b);
        double ga = lgamma(a);
        double gb = lgamma(b);
        if (x == 0.0 || x == 1.0) {
            if (a > 1.0 && b > 1.0)
                p = 0.0;
            else
                p = exp(gab - ga - gb) * pow(x, a - 1) * pow(1 - x, b - 1);
        } else {
            p = exp(gab - ga - gb + log(x) * (a - 1) + log1p(-x) * (b - 1));
        }
        return p;
    }
}

void gsl_ran_dirichlet(const gsl_rng *r, const size_t K, const double alpha[], double theta[])
{
    /* dead code in the reference (GraphORNOR.cpp:76-86 is commented out); kept so it links */
    double norm = 0.0;
    for (size_t i = 0; i < K; i++) theta[i] = gsl_ran_gamma(r, alpha[i], 1.0);
    for (size_t i = 0; i < K; i++) norm += theta[i];
    for (size_t i = 0; i < K; i++) theta[i] /= norm;
}

double gsl_ran_dirichlet_lnpdf(const size_t K, const double alpha[], const double theta[])
{
    double log_p = 0.0, sum_alpha = 0.0;
    for (size_t i = 0; i < K; i++) log_p += (alpha[i] - 1.0) * log(theta[i]);
    for (size_t i = 0; i < K; i++) sum_alpha += alpha[i];
    log_p += lgamma(sum_alpha);
    for (size_t i = 0; i < K; i++) log_p -= lgamma(alpha[i]);
    return log_p;
}

double gsl_ran_dirichlet_pdf(const size_t K, const double alpha[], const double theta[])
{
    return exp(gsl_ran_dirichlet_lnpdf(K, alpha, theta));
}

/* ------------------------------------------------------------------ rstat */
gsl_rstat_workspace *gsl_rstat_alloc(void)
{
    gsl_rstat_workspace *w = (gsl_rstat_workspace *) calloc(1, sizeof(gsl_rstat_workspace));
    gsl_rstat_reset(w);
    return w;
}

void gsl_rstat_free(gsl_rstat_workspace *w) { free(w); }

size_t gsl_rstat_n(const gsl_rstat_workspace *w) { return w->n; }

int gsl_rstat_add(const double x, gsl_rstat_workspace *w)
{
    /* rstat/rstat.c gsl_rstat_add: same update order as GSL */
    double delta = x - w->mean;
    double delta_n, delta_nsq, term1, n;
    if (w->n == 0) { w->min = x; w->max = x; }
    else { if (x < w->min) w->min = x; if (x > w->max) w->max = x; }
    n = (double) ++(w->n);
    delta_n = delta / n;
    delta_nsq = delta_n * delta_n;
    term1 = delta * delta_n * (n - 1.0);
    w->mean += delta_n;
    w->M4 += term1 * delta_nsq * (n * n - 3.0 * n + 3.0) + 6.0 * delta_nsq * w->M2 - 4.0 * delta_n * w->M3;
    w->M3 += term1 * delta_n * (n - 2.0) - 3.0 * delta_n * w->M2;
    w->M2 += term1;
    return 0;
}

double gsl_rstat_mean(const gsl_rstat_workspace *w) { return w->mean; }

double gsl_rstat_variance(const gsl_rstat_workspace *w)
{
    if (w->n > 1) {
        double n = (double) w->n;
        return w->M2 / (n - 1.0);
    }
    return 0.0;
}

double gsl_rstat_sd(const gsl_rstat_workspace *w) { return sqrt(gsl_rstat_variance(w)); }

int gsl_rstat_reset(gsl_rstat_workspace *w)
{
    w->min = 0.0; w->max = 0.0; w->mean = 0.0;
    w->M2 = 0.0; w->M3 = 0.0; w->M4 = 0.0;
    w->n = 0;
    return 0;
}
