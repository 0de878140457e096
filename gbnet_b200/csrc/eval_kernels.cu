// K6 — parity hooks: per-node conditional log-likelihoods, target likelihoods and the joint
// log-probability of a chain's CURRENT state, evaluated with the reference's own expressions
// and operand order (no cached products).  These are what the 1e-12 relative fp64 parity check
// of BASELINE.json compares against the oracle; they are not on the sampling path.
#include "kernels.cuh"
#include "device_math.cuh"

namespace gbn {

namespace {

struct ChainPtrs {
    const uint8_t *X;
    const double *T;
    const uint8_t *S;
};

__device__ __forceinline__ ChainPtrs chain_ptrs(const DevNet &net, const DevChains &ch, uint32_t c)
{
    return ChainPtrs{ ch.X + (size_t) c * net.nX, ch.T + (size_t) c * net.nX, ch.S + (size_t) c * net.Ep };
}

// XNode/TNode::get_children_loglikelihood (XNode.cpp:16-22, TNode.cpp:15-25) with X_k or T_k
// overridden; children summed in ascending target order
__device__ double children_loglik(const DevNet &net, const ChainPtrs &p, uint32_t d, uint32_t k,
                                  bool ov_x, uint32_t xv, bool ov_t, double tv)
{
    double loglik = 0.;
    for (uint32_t c = net.tf_ptr[k]; c < net.tf_ptr[k + 1]; c++) {
        uint32_t i = net.tf_child[c];
        double L = target_likelihood(net, d, i,
            [&](uint32_t kk, uint32_t) -> uint32_t { return (ov_x && kk == k) ? xv : (uint32_t) p.X[kk]; },
            [&](uint32_t kk, uint32_t) -> double { return (ov_t && kk == k) ? tv : p.T[kk]; },
            [&](uint32_t q) -> uint32_t { return p.S[q]; });
        loglik += log(L);
    }
    return loglik;
}

// RVNode::get_blanket_loglikelihood (RVNode.cpp:94-106) for every candidate of every node
__global__ void k_conditionals(DevNet net, DevChains ch, uint32_t c, double *out)
{
    const uint32_t nT = net.const_params ? 0 : net.nX;
    const uint32_t n_nodes = net.nX + nT + net.E;
    const uint32_t d = chain_dataset(net, c);
    const ChainPtrs p = chain_ptrs(net, ch, c);
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x) {
        double o[3] = { NAN, NAN, NAN };
        if (n < net.nX) {
            uint32_t k = n;
            const double *prob = net.xprob[net.x_prior_active[k]];
            for (uint32_t v = 0; v < 2; v++) {
                double loglik = 0.;
                loglik += log(prob[v]);
                loglik += children_loglik(net, p, d, k, true, v, false, 0.);
                o[v] = loglik;
            }
        } else if (n < net.nX + nT) {
            uint32_t k = n - net.nX;
            double loglik = 0.;
            loglik += log(beta_pdf(p.T[k], net.t_a[k], net.t_b[k], net.t_lnB[k]));
            if (net.listen) loglik += children_loglik(net, p, d, k, false, 0, false, 0.);
            else loglik += 0.;
            o[0] = loglik;
        } else {
            uint32_t e = n - net.nX - nT;
            uint32_t q = net.slot_of_edge[e];
            uint32_t i = net.slot_trg[q];
            for (uint32_t v = 0; v < 3; v++) {
                double loglik = 0.;
                loglik += log(net.sprob[net.mor_idx[q]][v]);
                double L = target_likelihood(net, d, i,
                    [&](uint32_t kk, uint32_t) -> uint32_t { return p.X[kk]; },
                    [&](uint32_t kk, uint32_t) -> double { return p.T[kk]; },
                    [&](uint32_t qq) -> uint32_t { return qq == q ? v : (uint32_t) p.S[qq]; });
                loglik += log(L);
                o[v] = loglik;
            }
        }
        out[3 * (size_t) n + 0] = o[0];
        out[3 * (size_t) n + 1] = o[1];
        out[3 * (size_t) n + 2] = o[2];
    }
}

__global__ void k_target_likelihoods(DevNet net, DevChains ch, uint32_t c, double *out)
{
    const uint32_t d = chain_dataset(net, c);
    const ChainPtrs p = chain_ptrs(net, ch, c);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < net.nH; i += gridDim.x * blockDim.x)
        out[net.ref_of_dt[i]] = target_likelihood(net, d, i,
            [&](uint32_t kk, uint32_t) -> uint32_t { return p.X[kk]; },
            [&](uint32_t kk, uint32_t) -> double { return p.T[kk]; },
            [&](uint32_t q) -> uint32_t { return p.S[q]; });
}

// joint log-probability: sum of get_own_loglikelihood over X, T, S (edge input order) and the
// latent H nodes.  One block; each thread sums a strided subset, then a fixed-order tree.
__global__ void k_joint_logprob(DevNet net, DevChains ch, uint32_t c, double *out)
{
    __shared__ double scratch[64];
    const uint32_t d = chain_dataset(net, c);
    const ChainPtrs p = chain_ptrs(net, ch, c);
    double lp = 0., unused = 0.;
    for (uint32_t k = threadIdx.x; k < net.nX; k += blockDim.x)
        lp += log(net.xprob[net.x_prior_active[k]][p.X[k]]);
    if (!net.const_params)
        for (uint32_t k = threadIdx.x; k < net.nX; k += blockDim.x)
            lp += log(beta_pdf(p.T[k], net.t_a[k], net.t_b[k], net.t_lnB[k]));
    for (uint32_t q = threadIdx.x; q < net.Ep; q += blockDim.x)
        if (net.par_edge[q] != 0xffffffffu) lp += log(net.sprob[net.mor_idx[q]][p.S[q]]);
    for (uint32_t i = threadIdx.x; i < net.nH; i += blockDim.x) {
        double L = target_likelihood(net, d, i,
            [&](uint32_t kk, uint32_t) -> uint32_t { return p.X[kk]; },
            [&](uint32_t kk, uint32_t) -> double { return p.T[kk]; },
            [&](uint32_t q) -> uint32_t { return p.S[q]; });
        lp += log(L);
    }
    block_sum2(lp, unused, scratch);
    if (threadIdx.x == 0) *out = lp;
}

__global__ void k_philox_kat(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out)
{
    u32x4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// the draws the keyed streams give one chain (orn_export_draws in the oracle)
__global__ void k_export_draws(DevNet net, uint32_t gchain, uint32_t sweep0, uint32_t n,
                               double *flat_u, double *beta_v, double *exp_v)
{
    const uint32_t per = net.nX + net.E;
    const size_t total = (size_t) n * per;
    for (size_t idx = blockIdx.x * (size_t) blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x) {
        uint32_t i = (uint32_t) (idx / per), j = (uint32_t) (idx % per);
        uint32_t sw = sweep0 + i;
        if (j < net.nX) {
            flat_u[idx] = keyed_u(net.seed, gchain, sw, j, KIND_X);
            beta_v[(size_t) i * net.nX + j] = draw_beta(net.seed, gchain, sw, j, net.t_a[j], net.t_b[j]);
            exp_v[(size_t) i * net.nX + j] = draw_exp(net.seed, gchain, sw, j);
        } else {
            uint32_t e = j - net.nX;
            const uint32_t q = net.slot_of_edge[e], i = net.slot_trg[q];
            flat_u[idx] = keyed_s(net.seed, gchain, sw, net.ref_of_dt[i], (q - slot0(net, i)) >> 5);
        }
    }
}

// Multinomial ctor draw for X (Multinomial.cpp:42-49), T = prior mean, S = mor + 1
__global__ void k_init_chains(DevNet net, DevChains ch, double c00, double c01, double c10, double c11, int skip_s)
{
    const uint32_t c = blockIdx.y;
    const uint32_t gchain = chain_global_id(net, c);
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < net.nX) {
        const bool act = net.x_prior_active[j] != 0;
        const double c0 = act ? c10 : c00, c1 = act ? c11 : c01;
        const double u = keyed_u(net.seed, gchain, 0, j, KIND_INIT);
        const double dice = 0. * (1 - u) + c1 * u;
        // simulation mode: X is a constant, 1 for the TFs of active_tf_set (GraphORNOR.cpp:58-63)
        ch.X[(size_t) c * net.nX + j] = net.sim ? (act ? 1 : 0) : (dice < c0 ? 0 : 1);
        ch.T[(size_t) c * net.nX + j] = net.t_mean[j];
    }
    if (!skip_s && j < net.Ep) ch.S[(size_t) c * net.Ep + j] = net.mor_idx[j];     // pads: 1 = edge off
}

// one warp per (dataset, TF): the children counted by their evidence class, the bound from the three counts
__global__ void k_tf_underflow_flags(DevNet net, uint8_t *uf, double l0, double l1, double l2, double threshold)
{
    const uint32_t d = blockIdx.y;
    const uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= net.nX) return;
    const uint8_t *y = net.y + (size_t) d * net.nH;
    uint32_t n0 = 0, n2 = 0, n = 0;
    for (uint32_t c = net.tf_ptr[k] + lane; c < net.tf_ptr[k + 1]; c += 32) {
        const uint32_t v = y[net.tf_child[c]];
        n0 += v == 0; n2 += v == 2; n += 1;
    }
    n0 = __reduce_add_sync(0xffffffffu, n0); n2 = __reduce_add_sync(0xffffffffu, n2); n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) uf[(size_t) d * net.nX + k] = ((double) n0 * l0 + (double) (n - n0 - n2) * l1 + (double) n2 * l2 < threshold) ? 1 : 0;
}

}  // namespace

void launch_tf_underflow_flags(const DevNet &net, uint8_t *uf, const double lmin[3], double threshold, cudaStream_t st)
{
    k_tf_underflow_flags<<<dim3((net.nX + 7) / 8, net.n_datasets), 256, 0, st>>>(net, uf, lmin[0], lmin[1], lmin[2], threshold);
}

// skip_s_bytes: the fast path fills its packed views from the network's initial words (launch_fast_refresh, initial);
// the byte view of S is then unpacked on demand
void launch_init_chains(const DevNet &net, const DevChains &ch, const double cum[2][2], cudaStream_t st, bool skip_s_bytes)
{
    uint32_t n = skip_s_bytes ? net.nX : (net.nX > net.Ep ? net.nX : net.Ep);
    k_init_chains<<<dim3((n + 255) / 256, ch.n_chains), 256, 0, st>>>(net, ch, cum[0][0], cum[0][1], cum[1][0], cum[1][1], skip_s_bytes ? 1 : 0);
}

void launch_conditionals(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st)
{
    uint32_t n_nodes = net.nX + (net.const_params ? 0 : net.nX) + net.E;
    uint32_t blocks = (n_nodes + 127) / 128;
    k_conditionals<<<blocks, 128, 0, st>>>(net, ch, c, out);
}

void launch_target_likelihoods(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st)
{
    k_target_likelihoods<<<(net.nH + 127) / 128, 128, 0, st>>>(net, ch, c, out);
}

void launch_joint_logprob(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st)
{
    k_joint_logprob<<<1, 1024, 0, st>>>(net, ch, c, out);
}

void launch_philox_kat(const uint32_t ctr[4], const uint32_t key[2], uint32_t *out, cudaStream_t st)
{
    k_philox_kat<<<1, 1, 0, st>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

void launch_export_draws(const DevNet &net, uint32_t gchain, uint32_t sweep0, uint32_t n,
                         double *flat_u, double *beta_v, double *exp_v, cudaStream_t st)
{
    size_t total = (size_t) n * (net.nX + net.E);
    uint32_t blocks = (uint32_t) ((total + 127) / 128);
    if (blocks > 4096) blocks = 4096;
    if (blocks == 0) blocks = 1;
    k_export_draws<<<blocks, 128, 0, st>>>(net, gchain, sweep0, n, flat_u, beta_v, exp_v);
}

}  // namespace gbn
