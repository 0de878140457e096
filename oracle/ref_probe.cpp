/* C-ABI probe over the UNMODIFIED reference C++ — TEST INFRASTRUCTURE ONLY.
 *
 * Built by oracle/Makefile together with /root/reference/libgbnet/src/*.cpp (compiled where
 * they lie, against oracle/gsl_shim) into oracle/_ref/libgbnet_ref.so.  Nothing from the
 * reference is copied into this repository; this file only #includes its headers and calls
 * its public methods:
 *
 *   gbn::ModelORNOR ctor            libgbnet/src/ModelORNOR.cpp:13-28
 *   ModelBase::sample / sample_n    libgbnet/src/ModelBase.cpp:114-136
 *   ModelBase::get_gelman_rubin     libgbnet/src/ModelBase.cpp:47-112
 *   ModelBase::get_posterior_*      libgbnet/src/ModelBase.cpp:160-209
 *   RVNode::get_blanket_loglikelihood  libgbnet/src/RVNode.cpp:94-106
 *
 * Nodes are addressed by their index in GraphBase::random_nodes (X[sorted uid], T[sorted
 * uid] unless const_params, S[edge input order]: GraphORNOR.cpp:57-73,213-233,235-264).
 * Integer ids are turned into zero-padded decimal strings so that the reference's
 * lexicographic std::set order equals numeric order.
 */
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <omp.h>

#include "ModelORNOR.h"

/* defined at ModelBase.cpp:211-214 but declared in no header */
namespace gbn { void set_signal(int); }

namespace {

struct GraphPeek : gbn::GraphBase {
    /* legal access to the protected member through a derived class */
    static gsl_rng *rng_of(gbn::GraphBase *g) { return g->*(&GraphPeek::rng); }
};

std::string padded(unsigned int v)
{
    char buf[16];
    std::snprintf(buf, sizeof buf, "%08u", v);
    return std::string(buf);
}

struct RefModel {
    gbn::ModelORNOR *model;
    /* per chain, the latent H nodes in target-uid order (from norand_nodes) */
    std::vector<std::vector<gbn::HNode *>> hnodes;
    /* per chain, keeps injected arrays alive */
    std::vector<std::vector<double>> inj_flat, inj_beta, inj_exp;
};

char g_error[512] = "";

}  // namespace

extern "C" {

const char *ref_last_error(void) { return g_error; }

void *ref_model_create(unsigned int n_edge, const unsigned int *src, const unsigned int *trg,
                       const int *mor,
                       unsigned int n_evid, const unsigned int *evid_trg, const int *evid_val,
                       unsigned int n_active, const unsigned int *active_tf,
                       const double *sprior9,
                       double z_alpha, double z_beta, double z0_alpha, double z0_beta,
                       double t_alpha, double t_beta,
                       unsigned int n_graphs, int noise_listen_children, int comp_yprob,
                       int const_params,
                       double t_focus, double t_lmargin, double t_hmargin,
                       double zn_focus, double zn_lmargin, double zn_hmargin)
{
    try {
        gbn::network_t network;
        for (unsigned int e = 0; e < n_edge; e++)
            network.push_back(gbn::network_edge_t(gbn::src_trg_pair_t(padded(src[e]), padded(trg[e])), mor[e]));
        gbn::evidence_dict_t evidence;
        for (unsigned int i = 0; i < n_evid; i++)
            evidence.insert(gbn::trg_de_pair_t(padded(evid_trg[i]), evid_val[i]));
        gbn::prior_active_tf_set_t active;
        for (unsigned int i = 0; i < n_active; i++)
            active.insert(padded(active_tf[i]));
        double sp[9];
        std::memcpy(sp, sprior9, sizeof sp);

        RefModel *rm = new RefModel();
        rm->model = new gbn::ModelORNOR(network, evidence, active, sp, z_alpha, z_beta, z0_alpha, z0_beta,
                                        t_alpha, t_beta, n_graphs, noise_listen_children != 0,
                                        comp_yprob != 0, const_params != 0,
                                        t_focus, t_lmargin, t_hmargin, zn_focus, zn_lmargin, zn_hmargin);
        rm->hnodes.resize(n_graphs);
        for (unsigned int c = 0; c < n_graphs; c++)
            for (auto node : rm->model->graphs[c]->norand_nodes)
                if (node->name == "H")
                    rm->hnodes[c].push_back(static_cast<gbn::HNode *>(node));
        rm->inj_flat.resize(n_graphs);
        rm->inj_beta.resize(n_graphs);
        rm->inj_exp.resize(n_graphs);
        return rm;
    } catch (const std::exception &ex) {
        std::snprintf(g_error, sizeof g_error, "%s", ex.what());
        return nullptr;
    }
}

void ref_model_destroy(void *h)
{
    RefModel *rm = static_cast<RefModel *>(h);
    delete rm->model;
    delete rm;
}

unsigned int ref_n_chains(void *h) { return static_cast<RefModel *>(h)->model->n_graphs; }

unsigned int ref_n_nodes(void *h)
{
    return (unsigned int) static_cast<RefModel *>(h)->model->graphs[0]->random_nodes.size();
}

unsigned int ref_n_targets(void *h) { return (unsigned int) static_cast<RefModel *>(h)->hnodes[0].size(); }

/* kind: 'X', 'T' or 'S'; returns n_stats; id copied into id_out */
unsigned int ref_node_info(void *h, unsigned int node, char *kind, char *id_out, unsigned int id_cap)
{
    gbn::RVNode *n = static_cast<RefModel *>(h)->model->graphs[0]->random_nodes[node];
    *kind = n->name[0];
    if (id_out) std::snprintf(id_out, id_cap, "%s", n->id.c_str());
    return n->n_stats;
}

static double node_value(gbn::RVNode *n)
{
    if (n->is_discrete) return (double) static_cast<gbn::Multinomial *>(n)->value;
    return static_cast<gbn::Beta *>(n)->value;
}

static void node_set_value(gbn::RVNode *n, double v)
{
    if (n->is_discrete) static_cast<gbn::Multinomial *>(n)->value = (unsigned int) v;
    else static_cast<gbn::Beta *>(n)->value = v;
}

void ref_get_state(void *h, unsigned int chain, double *out)
{
    auto &nodes = static_cast<RefModel *>(h)->model->graphs[chain]->random_nodes;
    for (size_t i = 0; i < nodes.size(); i++) out[i] = node_value(nodes[i]);
}

void ref_set_state(void *h, unsigned int chain, const double *in)
{
    auto &nodes = static_cast<RefModel *>(h)->model->graphs[chain]->random_nodes;
    for (size_t i = 0; i < nodes.size(); i++) node_set_value(nodes[i], in[i]);
}

/* RVNode::get_blanket_loglikelihood with the node temporarily set to `value` */
double ref_blanket_loglik(void *h, unsigned int chain, unsigned int node, double value)
{
    gbn::RVNode *n = static_cast<RefModel *>(h)->model->graphs[chain]->random_nodes[node];
    double keep = node_value(n);
    node_set_value(n, value);
    double ll = n->get_blanket_loglikelihood();
    node_set_value(n, keep);
    return ll;
}

/* all candidates of all nodes: out[node*3 + v]; X uses v=0,1; S v=0,1,2;
 * T uses v=0 only (blanket at the current value) */
void ref_all_conditionals(void *h, unsigned int chain, double *out)
{
    auto &nodes = static_cast<RefModel *>(h)->model->graphs[chain]->random_nodes;
    for (size_t i = 0; i < nodes.size(); i++) {
        gbn::RVNode *n = nodes[i];
        out[3 * i + 0] = out[3 * i + 1] = out[3 * i + 2] = NAN;
        if (n->is_discrete) {
            for (unsigned int v = 0; v < n->n_outcomes; v++)
                out[3 * i + v] = ref_blanket_loglik(h, chain, (unsigned int) i, (double) v);
        } else {
            out[3 * i] = n->get_blanket_loglikelihood();
        }
    }
}

/* HNode::get_own_likelihood (HNode.cpp:33-53) per target, sorted-uid order */
void ref_target_likelihoods(void *h, unsigned int chain, double *out)
{
    auto &hn = static_cast<RefModel *>(h)->hnodes[chain];
    for (size_t i = 0; i < hn.size(); i++) out[i] = hn[i]->get_own_likelihood();
}

/* log P(X,T,S) + log P(Y=y | X,T,S) with H marginalised, from the reference's own methods:
 * sum of get_own_loglikelihood over random nodes plus over the latent H nodes */
double ref_joint_logprob(void *h, unsigned int chain)
{
    RefModel *rm = static_cast<RefModel *>(h);
    double lp = 0.;
    for (auto n : rm->model->graphs[chain]->random_nodes) lp += n->get_own_loglikelihood();
    for (auto hn : rm->hnodes[chain]) lp += hn->get_own_loglikelihood();
    return lp;
}

void ref_inject(void *h, unsigned int chain,
                const double *flat_u, size_t n_flat,
                const double *beta_v, size_t n_beta,
                const double *exp_v, size_t n_exp)
{
    RefModel *rm = static_cast<RefModel *>(h);
    rm->inj_flat[chain].assign(flat_u, flat_u + n_flat);
    rm->inj_beta[chain].assign(beta_v, beta_v + n_beta);
    rm->inj_exp[chain].assign(exp_v, exp_v + n_exp);
    gslshim_rng_inject(GraphPeek::rng_of(rm->model->graphs[chain]),
                       rm->inj_flat[chain].data(), n_flat,
                       rm->inj_beta[chain].data(), n_beta,
                       rm->inj_exp[chain].data(), n_exp);
}

void ref_clear_injection(void *h, unsigned int chain)
{
    gslshim_rng_clear_injection(GraphPeek::rng_of(static_cast<RefModel *>(h)->model->graphs[chain]));
}

void ref_seed(void *h, unsigned int chain, unsigned long seed)
{
    gsl_rng_set(GraphPeek::rng_of(static_cast<RefModel *>(h)->model->graphs[chain]), seed);
}

void ref_sample_n(void *h, unsigned int n) { static_cast<RefModel *>(h)->model->sample_n(n); }
void ref_sample(void *h, unsigned int N, unsigned int dN) { static_cast<RefModel *>(h)->model->sample(N, dN); }
void ref_burn_stats(void *h) { static_cast<RefModel *>(h)->model->burn_stats(); }

/* seconds spent inside ModelBase::sample_n(n), construction excluded */
double ref_time_sample_n(void *h, unsigned int n)
{
    double t0 = omp_get_wtime();
    static_cast<RefModel *>(h)->model->sample_n(n);
    return omp_get_wtime() - t0;
}

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { omp_set_num_threads(n); }

double ref_chain_length(void *h)
{
    return static_cast<RefModel *>(h)->model->graphs[0]->random_nodes[0]->chain_length();
}

double ref_node_mean(void *h, unsigned int chain, unsigned int node, unsigned int stat)
{
    return static_cast<RefModel *>(h)->model->graphs[chain]->random_nodes[node]->mean(stat);
}

double ref_node_variance(void *h, unsigned int chain, unsigned int node, unsigned int stat)
{
    return static_cast<RefModel *>(h)->model->graphs[chain]->random_nodes[node]->variance(stat);
}

double ref_max_gelman_rubin(void *h) { return static_cast<RefModel *>(h)->model->get_max_gelman_rubin(); }

void ref_gelman_rubin(void *h, double *out)
{
    auto v = static_cast<RefModel *>(h)->model->get_gelman_rubin();
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i].second;
}

/* returns the number of (node,stat) entries written; name is "X", "T" or "S" */
unsigned int ref_posterior_means(void *h, const char *name, double *out)
{
    auto v = static_cast<RefModel *>(h)->model->get_posterior_means(std::string(name));
    if (out) for (size_t i = 0; i < v.size(); i++) out[i] = v[i].second;
    return (unsigned int) v.size();
}

unsigned int ref_posterior_sdevs(void *h, const char *name, double *out)
{
    auto v = static_cast<RefModel *>(h)->model->get_posterior_sdevs(std::string(name));
    if (out) for (size_t i = 0; i < v.size(); i++) out[i] = v[i].second;
    return (unsigned int) v.size();
}

void ref_set_signal(int s) { gbn::set_signal(s); }

/* Beta prior parameters of a T node (GraphORNOR.cpp:213-223) */
void ref_t_prior(void *h, unsigned int node, double *a, double *b)
{
    gbn::Beta *t = static_cast<gbn::Beta *>(static_cast<RefModel *>(h)->model->graphs[0]->random_nodes[node]);
    *a = t->prior_alpha; *b = t->prior_beta;
}

}  // extern "C"
