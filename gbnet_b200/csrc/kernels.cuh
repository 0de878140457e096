// Launchers of every kernel in the library (host-callable; defined in the .cu files).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "device_model.cuh"

namespace gbn {

// ---- eval_kernels.cu (parity hooks)
void launch_conditionals(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st);
void launch_target_likelihoods(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st);
void launch_joint_logprob(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st);
void launch_philox_kat(const uint32_t ctr[4], const uint32_t key[2], uint32_t *out, cudaStream_t st);
void launch_export_draws(const DevNet &net, uint32_t gchain, uint32_t sweep0, uint32_t n,
                         double *flat_u, double *beta_v, double *exp_v, cudaStream_t st);

// initial state of every local chain; cum[a][v] = cumulative prior likelihood of X (host libm)
// uf[d][k] = 1 when sum over TF k's children of lmin[y of the child] < threshold (the product of the children's
// likelihoods could underflow a double: capi.cu, upload_evidence)
void launch_tf_underflow_flags(const DevNet &net, uint8_t *uf, const double lmin[3], double threshold, cudaStream_t st);
void launch_init_chains(const DevNet &net, const DevChains &ch, const double cum[2][2], cudaStream_t st, bool skip_s_bytes = false);

// ---- sweep_strict.cu: n_sweeps full sweeps X -> T -> S of every local chain, one CTA per chain,
// every likelihood recomputed from the state in the reference's operand order.
// sweep0 = sweeps drawn since creation (Philox counter word), stat_n0 = sweeps since burn_stats.
// Returns the number of kernels launched, or -1 with *err set.
int launch_sweep_strict(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0,
                        uint32_t stat_n0, cudaStream_t st, const char **err);
size_t strict_smem_bytes(const DevNet &net);

// ---- sweep_fast.cu: the throughput path (cached per-target products, speculative X windows)
// phase_ev: NULL, or 3 * n_sweeps + 1 events recorded before the X, T, S kernels of every sweep and after the
// last kernel (per-kernel device times for the roofline report)
// chain0 / nchains: the local chains this call sweeps (the model layer runs disjoint chain groups on
// separate streams so that one group's X kernel overlaps another group's S kernel)
bool fast_t_beside_x(const DevNet &net);   // launch_sweep_fast leaves the current T in ch.T2 after an odd number of sweeps
bool fast_uses_x_stream(const DevNet &net);
bool fast_x1_supported(const DevNet &net);   // the X phase runs on the single-chain shared-memory kernel (needs ch.w0f / w2f / tq1)
uint32_t fast_x_bundle();    // chains per X-kernel CTA = interleave factor of the product cache
int launch_sweep_fast(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0,
                      uint32_t stat_n0, cudaStream_t st, const char **err, cudaEvent_t *phase_ev = nullptr,
                      uint32_t chain0 = 0, uint32_t nchains = UINT32_MAX,
                      cudaStream_t stx = nullptr, cudaEvent_t evx = nullptr, cudaEvent_t evs = nullptr);
// rebuild the packed S words, the by-TF copy, FXp and the product cache from X, T and the BYTE view of S
// (after creation, set_state, checkpoint load)
int launch_fast_refresh(const DevNet &net, const DevChains &ch, cudaStream_t st, const char **err, bool initial = false);
// the byte view of S (ch.S) from the packed words a fast sweep leaves behind
int launch_fast_unpack(const DevNet &net, const DevChains &ch, cudaStream_t st, const char **err);
// the S nodes' per-batch 8-bit outcome counts (ch.dC) added to the 32-bit totals (ch.cS); at most
// FAST_MAX_PENDING sweeps may be drawn between two folds
int launch_fast_fold(const DevNet &net, const DevChains &ch, cudaStream_t st, const char **err, bool overwrite = false);
constexpr uint32_t FAST_MAX_PENDING = 255;
bool fast_supported(const DevNet &net, const char **why);

// ---- sweep_small.cu: networks whose whole chain state fits one SM's shared memory - one persistent CTA per chain
// runs all the sweeps of a batch in ONE launch (state in at the start, state and per-batch counters out at the end)
bool small_supported(const DevNet &net, const char **why);
int launch_sweep_small(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0, uint32_t stat_n0,
                       cudaStream_t st, const char **err, uint32_t chain0 = 0, uint32_t nchains = UINT32_MAX);

// ---- sim_kernels.cu: simulation mode (GraphORNOR.cpp:28,58-63,89-106): n_sweeps sweeps H, Y -> T
int launch_sim_sweep(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0,
                     uint32_t stat_n0, cudaStream_t st, const char **err);
void launch_sim_conditionals(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st);
void launch_sim_export_draws(const DevNet &net, uint32_t gchain, uint32_t sweep0, uint32_t n,
                             double *flat_u, double *beta_v, double *exp_v, cudaStream_t st);

// ---- moments.cu (K5): cross-chain moments -> Gelman-Rubin and posterior summaries
// T nodes, pass 1: tsum[d][k] = { sum_k mean_k, sum_k var_k, sum of mean_k over local chains whose GLOBAL index is >= 1 }
void launch_moment_sums(const DevNet &net, const DevChains &ch, double N, double *tsum, cudaStream_t st);
// T nodes, pass 2: tdev[d][k] = { sum_k (mean_k - mu)^2, sum over global chains >= 1 of (mean_k - pmu)^2 }
void launch_moment_devs(const DevNet &net, const DevChains &ch, double N, const double *tsum, double *tdev, cudaStream_t st);
// discrete nodes: exact u64 sums of the STORED counts (layout in moments.cu; moment_isum_words() words per dataset);
// after the all-reduces, launch_expand_counts fills sums / psums / dev / pdev of every statistic
size_t moment_isum_words(const DevNet &net);
void launch_count_sums(const DevNet &net, const DevChains &ch, unsigned long long *isums, cudaStream_t st, bool with_pending = false,
                       bool totals_zero = false);
void launch_expand_counts(const DevNet &net, double N, const unsigned long long *isums, const double *tsum, const double *tdev,
                          double *sums, double *psums, double *dev, double *pdev, cudaStream_t st);
// finalize: Gelman-Rubin per node (random_nodes order) of dataset d, and its maximum
void launch_gelman_rubin(const DevNet &net, double N, const double *sums, const double *dev,
                         uint32_t dataset, double *gr_out /* [n_nodes] or NULL */, double *max_out /* [n_datasets] */,
                         cudaStream_t st);
// per-chain running moments in stat order (RVNode::mean / variance)
void launch_node_moments(const DevNet &net, const DevChains &ch, uint32_t c, double N, double *mean, double *var, cudaStream_t st);

// stat index helpers shared by host and device: per dataset, stats are laid out
//   X: 2 per TF (outcome 0, 1) | T: 1 per TF (absent if const_params) | S: 3 per edge in EDGE INPUT order
//   simulation mode: H_i, Y_i interleaved in reference target order (3 stats each) | T
__host__ __device__ inline uint32_t n_t_of(const DevNet &n) { return n.const_params ? 0 : n.nX; }
__host__ __device__ inline uint32_t n_stats_of(const DevNet &n) { return n.sim ? 6 * n.nH + n_t_of(n) : 2 * n.nX + n_t_of(n) + 3 * n.E; }
__host__ __device__ inline uint32_t n_nodes_of(const DevNet &n) { return n.sim ? 2 * n.nH + n_t_of(n) : n.nX + n_t_of(n) + n.E; }
// work items of the moment kernels: nodes in slot order, pad slots included (and skipped)
__host__ __device__ inline uint32_t n_work_of(const DevNet &n) { return n.sim ? 2 * n.nH + n_t_of(n) : n.nX + n_t_of(n) + n.Ep; }

}  // namespace gbn
