/* gbnet_b200 — C ABI of the B200-native OR-NOR Gibbs sampler.
 *
 * The reference has no C ABI: its boundary is the C++ class gbn::ModelORNOR as seen by Cython
 * (reference gbnet/ModelORNOR.pxd:24-39).  Every entry point below names the reference method
 * it replaces.  Strings never cross this boundary: the caller (gbnet_b200/ModelORNOR.pyx) maps
 * uid strings to integers whose numeric order equals the reference's lexicographic std::set
 * order (GraphORNOR.cpp:31-42) and keeps the uid tables for building result ids.
 *
 * All functions return 0 on success and a negative gbn_status on failure; gbn_last_error()
 * gives the message (the reference throws C++ exceptions from the ctor only, pxd:27-30).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * GBN_ERR_CUDA.
 */
#ifndef GBNET_B200_H
#define GBNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBN_ABI_VERSION 2

typedef enum gbn_status {
    GBN_OK = 0,
    GBN_ERR_INVALID = -1,      /* std::invalid_argument / std::out_of_range in the reference */
    GBN_ERR_CUDA = -2,         /* no device, launch failure, out of memory */
    GBN_ERR_UNSUPPORTED = -3,  /* a request the selected kernel / mode cannot serve (e.g. new evidence for a simulation-mode model) */
    GBN_ERR_COMM = -4          /* NCCL */
} gbn_status;

typedef struct gbn_model gbn_model;

/* The network: what ModelORNOR.pyx:56-62 marshals into network_t, integer-coded.
 * Edge order is the caller's input order; it defines the S-node order (GraphORNOR.cpp:235-264). */
typedef struct gbn_network {
    uint32_t n_edge;
    const uint32_t *src;        /* [n_edge] TF id   */
    const uint32_t *trg;        /* [n_edge] target id */
    const int32_t *mor;         /* [n_edge] -1 / 0 / +1 (GraphORNOR.cpp:39-41) */
    uint32_t n_active;
    const uint32_t *active_tf;  /* prior_active_tf_set_t (GraphORNOR.cpp:64-67) */
} gbn_network;

/* Evidence: evidence_dict_t (ModelORNOR.pyx:64-66).  n_datasets > 1 batches contrasts that share
 * the network: dataset d uses evid_val[d*n_evid .. (d+1)*n_evid) and owns n_graphs chains. */
typedef struct gbn_evidence {
    uint32_t n_datasets;
    uint32_t n_evid;
    const uint32_t *evid_trg;   /* [n_evid] target ids (ids not in the network are ignored) */
    const int32_t *evid_val;    /* [n_datasets * n_evid] -1 / 0 / +1 */
} gbn_evidence;

/* Hyper-parameters: same meaning, order and defaults as gbn::ModelORNOR's ctor
 * (libgbnet/include/ModelORNOR.h:26-29; gbnet/ModelORNOR.pyx:13-16). */
typedef struct gbn_params {
    double sprior[9];
    double z_alpha, z_beta, z0_alpha, z0_beta;
    double t_alpha, t_beta;              /* accepted and ignored, as in GraphORNOR.cpp:217-218 */
    uint32_t n_graphs;                   /* chains per dataset */
    int32_t noise_listen_children, comp_yprob, const_params;
    double t_focus, t_lmargin, t_hmargin;
    double zn_focus, zn_lmargin, zn_hmargin;   /* accepted and unused, as in the reference */
    /* ---- extensions (zero-initialise for reference behaviour) ---- */
    uint64_t seed;                       /* Philox key; the reference seeds from time(NULL) (ModelORNOR.cpp:24) */
    int32_t device;                      /* CUDA device ordinal */
    int32_t kernel;                      /* gbn_kernel_kind */
    /* chain sharding across ranks: this model instance owns chains
     * [chain_offset, chain_offset + n_graphs_local) of every dataset; n_graphs stays the GLOBAL
     * chain count.  n_graphs_local == 0 means "all of them" (single rank). */
    uint32_t chain_offset;
    uint32_t n_graphs_local;
} gbn_params;

typedef enum gbn_kernel_kind {
    GBN_KERNEL_AUTO = 0,
    GBN_KERNEL_STRICT = 1,   /* warp per chain, every product recomputed in the reference's operand order */
    GBN_KERNEL_FAST = 2      /* CTA per chain, cached per-target products, speculative windows */
} gbn_kernel_kind;

void gbn_params_default(gbn_params *p);          /* pyx:13-16 defaults */
const char *gbn_last_error(void);
int gbn_abi_version(void);
int gbn_device_count(void);

/* new gbn::ModelORNOR(...)  (ModelORNOR.cpp:13-28 -> GraphORNOR::build_structure)
 *
 * SIMULATION MODE.  As in the reference (GraphORNOR.cpp:28), EMPTY evidence (ev == NULL or n_evid == 0)
 * builds a forward simulator instead of a sampler for inference: X is fixed (1 for the TFs of
 * active_tf, 0 otherwise) and S is fixed at mor + 1; the random nodes, in sweep order, are
 * H_0 Y_0 H_1 Y_1 ... (targets in sorted id order: latent and observed DE state, 3 outcomes each)
 * followed by the T nodes.  n_nodes = 2 n_trg + nT, n_stats = 6 n_trg + nT; get/set_state,
 * gelman_rubin, eval_conditionals and inject_draws (flat_u[chain][sweep][2 n_trg]) use that order;
 * the posterior getters answer to 'H', 'Y' and 'T'. */
int gbn_model_create(const gbn_network *net, const gbn_evidence *ev, const gbn_params *p, gbn_model **out);
/* delete model (ModelORNOR.cpp:7-11) */
void gbn_model_destroy(gbn_model *m);

typedef struct gbn_dims {
    uint32_t n_tf, n_trg, n_edge;
    uint32_t n_nodes;            /* random nodes per chain: nX + (const_params ? 0 : nX) + E */
    uint32_t n_datasets, n_graphs, n_graphs_local, chain_offset;
    uint32_t n_stats;            /* 2 nX + (const_params ? 0 : nX) + 3 E */
    uint32_t simulation;         /* 1: built with empty evidence (simulation mode, see below) */
} gbn_dims;
int gbn_model_dims(const gbn_model *m, gbn_dims *out);
/* the gbn_kernel_kind the model runs (GBN_KERNEL_AUTO resolved) */
int gbn_model_kernel(const gbn_model *m);
/* which sweep path serves the model: 0 = strict kernel (sweep_strict.cu), 1 = fast multi-kernel path with the
 * chain state in HBM (sweep_fast.cu), 2 = fast persistent kernel with the chain state in shared memory for a
 * whole batch of sweeps (sweep_small.cu: networks that fit one SM, non-listening T nodes) */
int gbn_model_path(const gbn_model *m);
/* sorted unique ids (std::set order of GraphORNOR.cpp:31-42), for building result ids */
int gbn_model_tf_ids(const gbn_model *m, uint32_t *out /* [n_tf] */);
int gbn_model_trg_ids(const gbn_model *m, uint32_t *out /* [n_trg] */);

/* ModelBase::sample(N, dN) (ModelBase.cpp:114-129), including its batch-overshoot quirk */
int gbn_model_sample(gbn_model *m, uint32_t N, uint32_t dN);
/* ModelBase::sample_n(n) (ModelBase.cpp:131-136) */
int gbn_model_sample_n(gbn_model *m, uint32_t n);
/* ModelBase::burn_stats() (ModelBase.cpp:138-142) */
int gbn_model_burn_stats(gbn_model *m);
/* RVNode::chain_length() of graphs[0]->random_nodes[0] (ModelBase.cpp:54) */
int gbn_model_chain_length(const gbn_model *m, double *out);

/* ModelBase::get_max_gelman_rubin() (ModelBase.cpp:106-112); max over datasets */
int gbn_model_max_gelman_rubin(gbn_model *m, double *out);
/* ModelBase::get_gelman_rubin() (ModelBase.cpp:47-104): one value per random node, in
 * random_nodes order X[sorted id], T[sorted id], S[edge input order] */
int gbn_model_gelman_rubin(gbn_model *m, uint32_t dataset, double *out /* [n_nodes] */);
/* ModelBase::get_posterior_means / get_posterior_sdevs (ModelBase.cpp:160-209): kind is 'X', 'T'
 * or 'S'; entries are node-major, stat-minor (2, 1 or 3 per node); chain 0 is excluded exactly
 * as the reference does.  *n receives the entry count; out may be NULL to query it. */
int gbn_model_posterior_means(gbn_model *m, uint32_t dataset, char kind, double *out, uint32_t *n);
int gbn_model_posterior_sdevs(gbn_model *m, uint32_t dataset, char kind, double *out, uint32_t *n);
/* ModelBase::print_stats() (ModelBase.cpp:144-158) for dataset 0, ids synthesised from integers */
int gbn_model_print_stats(gbn_model *m);

/* gbn::set_signal (ModelBase.cpp:211-214) and the SIGINT handler of ModelBase.cpp:22-27,40 */
void gbn_set_signal(int signum);
void gbn_install_sigint_handler(void);

/* Replace the evidence of every dataset (batched contrasts re-use the uploaded network);
 * resets chain state and statistics exactly as a fresh construction would. */
int gbn_model_set_evidence(gbn_model *m, const gbn_evidence *ev);

/* ---- multi-GPU: chains sharded over ranks, moment sums all-reduced with NCCL ---- */
#define GBN_COMM_ID_BYTES 128
int gbn_comm_unique_id(void *id_out /* [GBN_COMM_ID_BYTES] */);
int gbn_model_attach_comm(gbn_model *m, int rank, int world, const void *id /* [GBN_COMM_ID_BYTES] */);

/* ---- the burn-in / convergence protocol as one call (reference gbnet/utils.py:93-135, sample_model) ----
 * burn_its times sample(N, dN); burn_stats(); then sample(N, dN) followed by the max Gelman-Rubin statistic until
 * that is <= 1.1 or max_its calls (burn-in included) have been made - the schedule that defines "time to
 * R-hat < 1.1".  The reference evaluates R-hat inside every sample() call as well (ModelBase.cpp:122); during
 * burn-in that value can end a call early only when N > dN, so with N <= dN (the reference's sample(20),
 * dN = 30) the burn-in batches are enqueued back to back without a host round trip, and every later batch
 * costs one 8-byte read.  *its_out = calls made, *rhat_out = the last max R-hat (inf if none was evaluated).
 * SIGINT ends the protocol after the current batch (ModelBase.cpp:121). */
int gbn_model_run_protocol(gbn_model *m, uint32_t burn_its, uint32_t max_its, uint32_t N, uint32_t dN,
                           uint32_t *its_out, double *rhat_out);

/* ---- enrichment pre-screen counts (reference gbnet/utils.py:10-35, get_tests) for many contrasts at once ----
 * Per contrast d and source TF (sorted unique src ids, as gbn_model_tf_ids): the six edge counts the Fisher
 * tests of get_tests are built from - a segmented reduction over the by-TF edge list on the device:
 *   counts[d][tf][0..5] = { val != 0 (targets without evidence included, as pandas' NaN != 0), val == 0,
 *                           type +1 & val +1, type +1 & val -1, type -1 & val +1, type -1 & val -1 }
 * evid_val[d][i] is the value of target evid_trg[i] in contrast d: -1, 0, +1, or GBN_EVID_ABSENT.
 * Every edge row counts (duplicates and type 0 included), as the reference's group-by does.
 * tf_ids_out may be NULL; *n_tf_out receives the number of sources (call with counts == NULL for the size). */
#define GBN_EVID_ABSENT 127
int gbn_fisher_counts(int device, const gbn_network *net, uint32_t n_datasets, uint32_t n_evid,
                      const uint32_t *evid_trg, const int8_t *evid_val /* [n_datasets][n_evid] */,
                      uint32_t *n_tf_out, uint32_t *tf_ids_out /* [n_tf] */, uint32_t *counts /* [n_datasets][n_tf][6] */);
/* message of the last failed gbn_fisher_counts call on this thread (that entry point needs no model) */
const char *gbn_fisher_last_error(void);

/* ---- parity / test hooks (per LOCAL chain index: dataset-major, chain-minor) ---- */
/* state in random_nodes order: X[nX], T[nX] (absent if const_params), S[E, edge input order] */
int gbn_model_get_state(gbn_model *m, uint32_t chain, double *out /* [n_nodes] */);
int gbn_model_set_state(gbn_model *m, uint32_t chain, const double *in /* [n_nodes] */);
/* RVNode::get_blanket_loglikelihood (RVNode.cpp:94-106) for every candidate value of every node:
 * out[node*3 + v]; X fills v = 0,1; S fills v = 0,1,2; T fills v = 0 (current value); rest NaN */
int gbn_model_eval_conditionals(gbn_model *m, uint32_t chain, double *out /* [n_nodes*3] */);
/* HNode::get_own_likelihood (HNode.cpp:33-53) of every target, sorted id order */
int gbn_model_target_likelihoods(gbn_model *m, uint32_t chain, double *out /* [n_trg] */);
/* sum of get_own_loglikelihood over X, T, S and the latent H nodes */
int gbn_model_joint_logprob(gbn_model *m, uint32_t chain, double *out);
/* draws for the next n_sweeps sweeps of EVERY local chain instead of the Philox streams:
 * flat_u[chain][sweep][nX + E] (X then S in edge input order), beta_v / exp_v [chain][sweep][nX] */
int gbn_model_inject_draws(gbn_model *m, uint32_t n_sweeps, const double *flat_u,
                           const double *beta_v, const double *exp_v);
/* the draws the keyed Philox streams give chain `chain` for sweeps [sweep0, sweep0+n) */
int gbn_model_export_draws(gbn_model *m, uint32_t chain, uint32_t sweep0, uint32_t n,
                           double *flat_u, double *beta_v, double *exp_v);
/* per-chain running statistics: RVNode::mean / RVNode::variance (RVNode.cpp:115-123) */
int gbn_model_node_moments(gbn_model *m, uint32_t chain, double *mean /* [n_stats] */,
                           double *variance /* [n_stats] */);
/* Philox4x32-10 evaluated on the device (known-answer test hook) */
int gbn_philox_kat(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- host-side topology compiler (test hook; needs no device) ----
 * Compiles a network and one evidence set exactly as gbn_model_create / gbn_model_set_evidence do
 * (GraphORNOR::build_structure, GraphORNOR.cpp:13-265: ranks, edge order, evidence map, YPROB counts)
 * and copies the device layout out.  Call once with NULL arrays for the sizes; any array may be NULL. */
typedef struct gbn_topology_dims {
    uint32_t n_tf, n_trg, n_edge, n_edge_unique;
    uint32_t n_slots;            /* padded (target, parent) slots, 32-target groups stored transposed */
    uint32_t n_groups, max_indeg, max_outdeg;
} gbn_topology_dims;
int gbn_topology_probe(const gbn_network *net, const gbn_evidence *ev, int comp_yprob, gbn_topology_dims *dims,
                       uint32_t *slot_of_edge /* [n_edge] */, uint32_t *deg /* [n_trg], device target order */,
                       uint32_t *ref_of_dt /* [n_trg] device target -> rank of its id */,
                       uint32_t *gbase /* [n_groups + 1] first slot of each group */,
                       uint32_t *tf_ptr /* [n_tf + 1] */, uint32_t *tf_child /* [n_edge_unique] device targets */,
                       uint32_t *tf_slot /* [n_edge_unique] */, uint8_t *y /* [n_trg] device target order: evidence + 1 */,
                       double yprob[3]);

/* ---- checkpoint / resume (absent from the reference) ----
 * One flat blob with the chain states, statistics and sweep counters of every local chain; loading it
 * into a model built from the same inputs continues the chains bit for bit. */
int gbn_model_checkpoint_size(gbn_model *m, size_t *bytes);
int gbn_model_checkpoint_save(gbn_model *m, void *out, size_t bytes);
int gbn_model_checkpoint_load(gbn_model *m, const void *in, size_t bytes);

/* ---- measurement ---- */
/* device time (CUDA events on the model's stream) of the last sample_n / sample call, in ms */
int gbn_model_last_device_ms(const gbn_model *m, double *ms);
/* per-kernel device times of the fast path: when switched on, CUDA events are recorded on the
 * model's stream around the X, T and S kernels of every sweep; gbn_model_phase_ms returns the
 * accumulated milliseconds per kernel and the number of sweeps they cover */
int gbn_model_set_phase_timing(gbn_model *m, int on);
int gbn_model_phase_ms(const gbn_model *m, double ms_out[3], uint64_t *sweeps);
/* the fast path sweeps disjoint chain groups on separate streams so that one group's X kernel overlaps
 * another group's S kernel; 0 restores the default (3 for >= 192 chains), 1 serialises the kernels (per-kernel timing) */
int gbn_model_set_stream_groups(gbn_model *m, uint32_t groups);
/* number of kernels this library launched on behalf of the model since creation */
int gbn_model_launch_count(const gbn_model *m, uint64_t *n);
/* total sweeps drawn since creation (Philox counter word) */
int gbn_model_sweep_count(const gbn_model *m, uint64_t *n);
/* device time between two points chosen by the caller (CUDA events on the model's stream) */
int gbn_model_timer_start(gbn_model *m);
int gbn_model_timer_stop(gbn_model *m, double *ms);
/* diagnostics since the last call: [0..2] the X phase's speculative windows (rounds, TFs committed, rounds
 * that ended in a flip), [3] prior fallbacks of S draws (Multinomial.cpp:78-85), [4..6] cycles one thread of the
 * persistent small-network kernel spent in the X, T and S phases (summed over chains and sweeps), [7] S updates whose
 * parent TF was active (the only ones that look the TF's factor up), [8] X updates that took the exact product of their
 * children's likelihoods (the magnitude bound said it could underflow), [9] X updates that carried the bound, [10] exact products that needed a
 * second pass over the children (not predicted by the TF's previous update);
 * enable != 0 switches counting on, 0 off */
int gbn_model_debug_counters(gbn_model *m, int enable, uint64_t out[16]);
int gbn_model_synchronize(gbn_model *m);

#ifdef __cplusplus
}
#endif
#endif /* GBNET_B200_H */
