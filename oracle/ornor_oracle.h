/* Plain-C restatement of gbnet's OR-NOR Gibbs sampler — TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (gbnet_b200/)
 * never links, imports or falls back to anything in oracle/.
 *
 * Pinning status: the reference ships no golden vectors and no assertions, so
 * this restatement is pinned against (a) the reference C++ itself, compiled unmodified into
 * oracle/_ref/libgbnet_ref.so (tests/test_oracle_vs_ref.py, run wherever that .so exists),
 * (b) fixtures that build produced, committed under tests/golden/ with their generating
 * script, and (c) the four known-answer vectors of SURVEY.md §8(c).
 */
#ifndef ORNOR_ORACLE_H
#define ORNOR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orn_model orn_model;

/* Hyper-parameters, same meaning and order as gbn::ModelORNOR's ctor (ModelORNOR.h:26-29). */
typedef struct orn_params {
    double sprior[9];
    double z_alpha, z_beta, z0_alpha, z0_beta, t_alpha, t_beta;
    unsigned int n_graphs;
    int noise_listen_children, comp_yprob, const_params;
    double t_focus, t_lmargin, t_hmargin, zn_focus, zn_lmargin, zn_hmargin;
    uint64_t seed;            /* Philox key (the reference seeds from time(NULL)) */
} orn_params;

const char *orn_last_error(void);

orn_model *orn_model_create(unsigned int n_edge, const unsigned int *src, const unsigned int *trg,
                            const int *mor,
                            unsigned int n_evid, const unsigned int *evid_trg, const int *evid_val,
                            unsigned int n_active, const unsigned int *active_tf,
                            const orn_params *params);
void orn_model_destroy(orn_model *m);

unsigned int orn_n_chains(const orn_model *m);
unsigned int orn_n_nodes(const orn_model *m);     /* random nodes: nX + (const_params ? 0 : nX) + E */
unsigned int orn_n_tf(const orn_model *m);
unsigned int orn_n_targets(const orn_model *m);
unsigned int orn_n_edges(const orn_model *m);
/* slot (position in the by-target CSR) of every edge, in edge input order */
void orn_edge_slots(const orn_model *m, unsigned int *slot_of_edge);
void orn_t_prior(const orn_model *m, unsigned int tf, double *a, double *b);

/* state in random_nodes order: X[nX], T[nX] (absent if const_params), S[E, edge input order] */
void orn_get_state(const orn_model *m, unsigned int chain, double *out);
void orn_set_state(orn_model *m, unsigned int chain, const double *in);

double orn_blanket_loglik(orn_model *m, unsigned int chain, unsigned int node, double value);
void orn_all_conditionals(orn_model *m, unsigned int chain, double *out /* [n_nodes*3] */);
void orn_target_likelihoods(orn_model *m, unsigned int chain, double *out /* [n_targets] */);
double orn_joint_logprob(orn_model *m, unsigned int chain);

/* draws: either the keyed Philox streams (default) or injected per-node draws */
void orn_inject(orn_model *m, unsigned int chain,
                const double *flat_u, size_t n_flat,     /* per sweep: X[nX] then S[E, edge order] */
                const double *beta_v, size_t n_beta,     /* per sweep: T[nX] Beta(a,b) draws       */
                const double *exp_v, size_t n_exp);      /* per sweep: T[nX] Exp(1) draws          */
void orn_clear_injection(orn_model *m, unsigned int chain);
/* the draws the keyed streams give for sweeps [sweep0, sweep0+n): same layout as orn_inject */
void orn_export_draws(const orn_model *m, unsigned int chain, unsigned int sweep0, unsigned int n,
                      double *flat_u, double *beta_v, double *exp_v);

void orn_sample_n(orn_model *m, unsigned int n);                 /* ModelBase.cpp:131-136 */
void orn_sample(orn_model *m, unsigned int N, unsigned int dN);  /* ModelBase.cpp:114-129 */
void orn_burn_stats(orn_model *m);                               /* ModelBase.cpp:138-142 */
double orn_time_sample_n(orn_model *m, unsigned int n);
int orn_omp_max_threads(void);
void orn_omp_set_threads(int n);
void orn_set_signal(int s);

double orn_chain_length(const orn_model *m);
double orn_node_mean(const orn_model *m, unsigned int chain, unsigned int node, unsigned int stat);
double orn_node_variance(const orn_model *m, unsigned int chain, unsigned int node, unsigned int stat);
void orn_gelman_rubin(const orn_model *m, double *out /* [n_nodes] */);   /* ModelBase.cpp:47-104 */
double orn_max_gelman_rubin(const orn_model *m);                          /* ModelBase.cpp:106-112 */
unsigned int orn_posterior_means(const orn_model *m, const char *name, double *out);  /* :160-183 */
unsigned int orn_posterior_sdevs(const orn_model *m, const char *name, double *out);  /* :185-209 */

/* building blocks exposed for known-answer tests */
void orn_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* P_model(h), h = 0,1,2, of one target from its parent list (HNodeORNOR.cpp:40-106) and the
 * marginal likelihood of each observed state L(y) = sum_h P(h) YNOISE[h][y] (HNode.cpp:33-53) */
void orn_target_kat(unsigned int n_parent, const double *t, const unsigned int *x,
                    const unsigned int *s, double z, const double yprob[3],
                    double p_out[3], double l_out[3]);
double orn_beta_pdf(double x, double a, double b);   /* GSL randist/beta.c semantics */
/* the keyed Beta(a,b) and Exp(1) draws of T node `tf` */
double orn_draw_beta(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int tf,
                     double a, double b);
double orn_draw_exp(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int tf);
double orn_draw_u(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int node,
                  unsigned int kind);

#ifdef __cplusplus
}
#endif
#endif
