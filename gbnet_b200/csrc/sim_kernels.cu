// Simulation mode: the reference's build_structure with EMPTY evidence
// (GraphORNOR.cpp:28) fixes X (1 for the TFs of active_tf_set, 0 otherwise: :58-63) and S (mor + 1,
// :247-248), and makes the per-target H and Y nodes random instead (:89-106, HNode.cpp:48-50):
//   H_i ~ Multinomial over h of HNodeORNOR::get_model_likelihood (no child term: RVNode.cpp:88-92)
//   Y_i ~ YNOISE[H_i][.]                                         (YDataNode.cpp:26-29)
// visited in the order H_0, Y_0, H_1, Y_1, ... then the T nodes.  Every target's noise parent is ZN,
// because the choice is made at construction while every Y still holds 1 (GraphORNOR.cpp:257-260).
// The sweep is a forward simulator of DE data consistent with the model; its statistics are the
// outcome frequencies of H and Y.  Expressions and operand order are the reference's (this path is
// not throughput-critical: one thread per (chain, target), parents rescanned).
#include "kernels.cuh"
#include "device_math.cuh"

namespace gbn {

namespace {

struct SimArgs {
    uint32_t sweep, stat_n, use_inj, inj_row;
};

__device__ __forceinline__ void sim_model_probs(const DevNet &net, const uint8_t *X, const double *T, const uint8_t *S,
                                                uint32_t d, uint32_t i, uint32_t ov_k, double ov_t, double p[3])
{
    ParentScan ps;
    const uint32_t deg = net.deg[i];
    uint32_t q = slot0(net, i);
    for (uint32_t j = 0; j < deg; j++, q += 32) {
        const uint32_t kk = net.par_tf[q];
        ps.add(kk == ov_k ? ov_t : T[kk], X[kk], S[q], net.z_n);
    }
    ps.model_probs(net.yprob + 3 * (size_t) d, p);
}

// H_i then Y_i of every target (they only depend on X, T, S: all targets are independent)
__global__ void k_sim_hy(DevNet net, DevChains ch, SimArgs a)
{
    const uint32_t c = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;       // device target
    if (i >= net.nH) return;
    const uint32_t d = chain_dataset(net, c);
    const uint32_t gchain = chain_global_id(net, c);
    const uint32_t ref_i = net.ref_of_dt[i];
    const double *yp = net.yprob + 3 * (size_t) d;
    const size_t o = (size_t) c * net.nH + i;
    double p[3], llik[3];
    sim_model_probs(net, ch.X + (size_t) c * net.nX, ch.T + (size_t) c * net.nX, ch.S + (size_t) c * net.Ep,
                    d, i, 0xffffffffu, 0., p);
    const double *inj = a.use_inj ? ch.inj_flat + ((size_t) c * ch.inj_sweeps + a.inj_row) * (2 * (size_t) net.nH) : nullptr;
    for (int h = 0; h < 3; h++) llik[h] = (0. + log(p[h])) + 0.;
    const double uh = inj ? inj[2 * ref_i] : keyed_u(net.seed, gchain, a.sweep, ref_i, KIND_H);
    const uint32_t h = multinomial_draw<3>(llik, yp, uh, ch.H[o]);
    ch.H[o] = (uint8_t) h;
    for (int y = 0; y < 3; y++) llik[y] = (0. + log(net.ynoise[h][y])) + 0.;
    const double uy = inj ? inj[2 * ref_i + 1] : keyed_u(net.seed, gchain, a.sweep, ref_i, KIND_Y);
    const uint32_t y = multinomial_draw<3>(llik, yp, uy, ch.Y[o]);
    ch.Y[o] = (uint8_t) y;
    uint2 cs = ch.cH[o]; cs.x += (h == 0); cs.y += (h == 2); ch.cH[o] = cs;
    cs = ch.cY[o]; cs.x += (y == 0); cs.y += (y == 2); ch.cY[o] = cs;
}

// T phase.  A TF with X = 0 never enters a target's products, so the blanket difference of its T is
// the prior's alone whether or not T listens to its children (the children terms are identical on
// both sides); the fixed active TFs of a listening model are walked in index order by one CTA.
constexpr int ST_NT = 256;
__global__ void __launch_bounds__(ST_NT) k_sim_t(DevNet net, DevChains ch, SimArgs a)
{
    __shared__ double scratch[2 * (ST_NT / 32)];
    const uint32_t c = blockIdx.x;
    const uint32_t tid = threadIdx.x;
    const uint32_t nX = net.nX;
    const uint32_t d = chain_dataset(net, c);
    const uint32_t gchain = chain_global_id(net, c);
    const uint8_t *X = ch.X + (size_t) c * nX;
    double *T = ch.T + (size_t) c * nX;
    const uint8_t *S = ch.S + (size_t) c * net.Ep;
    const uint8_t *H = ch.H + (size_t) c * net.nH;
    const double n_after = (double) a.stat_n;
    const size_t row = ((size_t) c * ch.inj_sweeps + a.inj_row) * nX;
    const bool seq = net.listen != 0;
    // independent updates (prior-only blanket)
    for (uint32_t k = tid; k < nX; k += ST_NT) {
        if (seq && X[k] != 0) continue;
        const double al = net.t_a[k], be = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
        const double prev = T[k];
        const double prev_ll = (0. + log(beta_pdf(prev, al, be, lnB))) + 0.;
        const double draw = a.use_inj ? ch.inj_beta[row + k] : draw_beta(net.seed, gchain, a.sweep, k, al, be);
        const double dx = draw - mean;
        const double prop = prev + dx;
        const double prop_ll = (0. + log(beta_pdf(prop, al, be, lnB))) + 0.;
        const double nt = mh_decide(prev, prop, prev_ll, prop_ll, mean, dx, al, be, lnB,
                                    a.use_inj != 0, a.use_inj ? ch.inj_exp[row + k] : 0., net.seed, gchain, a.sweep, k);
        T[k] = nt;
        const size_t o = (size_t) c * nX + k;
        double m = ch.tMean[o], m2 = ch.tM2[o];
        welford_add(m, m2, nt, n_after);
        ch.tMean[o] = m; ch.tM2[o] = m2;
    }
    __syncthreads();
    if (!seq) return;
    // listening T of the active TFs: TNode::get_children_loglikelihood (TNode.cpp:15-25) over log P_model(H_i)
    for (uint32_t k = 0; k < nX; k++) {
        if (X[k] == 0) continue;                              // uniform: X is the same for all threads
        const double al = net.t_a[k], be = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
        const double prev = T[k];
        const double draw = a.use_inj ? ch.inj_beta[row + k] : draw_beta(net.seed, gchain, a.sweep, k, al, be);
        const double dx = draw - mean;
        const double prop = prev + dx;
        double acc0 = 0., acc1 = 0.;
        for (uint32_t cc = net.tf_ptr[k] + tid; cc < net.tf_ptr[k + 1]; cc += ST_NT) {
            const uint32_t i = net.tf_child[cc];
            double p0[3], p1[3];
            sim_model_probs(net, X, T, S, d, i, k, prev, p0);
            sim_model_probs(net, X, T, S, d, i, k, prop, p1);
            acc0 += log(p0[H[i]]);
            acc1 += log(p1[H[i]]);
        }
        block_sum2(acc0, acc1, scratch);
        const double prev_ll = (0. + log(beta_pdf(prev, al, be, lnB))) + acc0;
        const double prop_ll = (0. + log(beta_pdf(prop, al, be, lnB))) + acc1;
        const double nt = mh_decide(prev, prop, prev_ll, prop_ll, mean, dx, al, be, lnB,
                                    a.use_inj != 0, a.use_inj ? ch.inj_exp[row + k] : 0., net.seed, gchain, a.sweep, k);
        __syncthreads();                                      // every thread has read T[k] = prev
        if (tid == 0) {
            T[k] = nt;
            const size_t o = (size_t) c * nX + k;
            double m = ch.tMean[o], m2 = ch.tM2[o];
            welford_add(m, m2, nt, n_after);
            ch.tMean[o] = m; ch.tM2[o] = m2;
        }
        __syncthreads();
    }
}

// RVNode::get_blanket_loglikelihood of every candidate of every random node, reference order:
// H_i (3), Y_i (3) interleaved by reference target, then T (current value)
__global__ void k_sim_conditionals(DevNet net, DevChains ch, uint32_t c, double *out)
{
    const uint32_t nT = n_t_of(net);
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t d = chain_dataset(net, c);
    const uint8_t *X = ch.X + (size_t) c * net.nX;
    const double *T = ch.T + (size_t) c * net.nX;
    const uint8_t *S = ch.S + (size_t) c * net.Ep;
    if (n < net.nH) {
        const uint32_t ref_i = net.ref_of_dt[n];
        double p[3];
        sim_model_probs(net, X, T, S, d, n, 0xffffffffu, 0., p);
        const uint32_t h = ch.H[(size_t) c * net.nH + n];
        for (int v = 0; v < 3; v++) {
            out[3 * (size_t) (2 * ref_i) + v] = (0. + log(p[v])) + 0.;
            out[3 * (size_t) (2 * ref_i + 1) + v] = (0. + log(net.ynoise[h][v])) + 0.;
        }
    } else if (n < net.nH + nT) {
        const uint32_t k = n - net.nH;
        double loglik = 0.;
        loglik += log(beta_pdf(T[k], net.t_a[k], net.t_b[k], net.t_lnB[k]));
        if (net.listen) {
            double acc = 0.;
            for (uint32_t cc = net.tf_ptr[k]; cc < net.tf_ptr[k + 1]; cc++) {
                const uint32_t i = net.tf_child[cc];
                double p[3];
                sim_model_probs(net, X, T, S, d, i, 0xffffffffu, 0., p);
                acc += log(p[ch.H[(size_t) c * net.nH + i]]);
            }
            loglik += acc;
        } else loglik += 0.;
        double *o = out + 3 * (size_t) (2 * net.nH + k);
        o[0] = loglik; o[1] = NAN; o[2] = NAN;
    }
}

__global__ void k_sim_export_draws(DevNet net, uint32_t gchain, uint32_t sweep0, uint32_t n,
                                   double *flat_u, double *beta_v, double *exp_v)
{
    const uint32_t per = 2 * net.nH > net.nX ? 2 * net.nH : net.nX;
    const size_t total = (size_t) n * per;
    for (size_t idx = blockIdx.x * (size_t) blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t) (idx / per), j = (uint32_t) (idx % per);
        const uint32_t sw = sweep0 + i;
        if (j < 2 * net.nH)
            flat_u[(size_t) i * 2 * net.nH + j] = keyed_u(net.seed, gchain, sw, j >> 1, (j & 1) ? KIND_Y : KIND_H);
        if (j < net.nX) {
            beta_v[(size_t) i * net.nX + j] = draw_beta(net.seed, gchain, sw, j, net.t_a[j], net.t_b[j]);
            exp_v[(size_t) i * net.nX + j] = draw_exp(net.seed, gchain, sw, j);
        }
    }
}

}  // namespace

int launch_sim_sweep(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0,
                     uint32_t stat_n0, cudaStream_t st, const char **err)
{
    int launches = 0;
    for (uint32_t sw = 0; sw < n_sweeps; sw++) {
        SimArgs a;
        a.sweep = sweep0 + sw;
        a.stat_n = stat_n0 + sw + 1;
        a.use_inj = (ch.inj_flat != nullptr && (ch.inj_pos + sw) < ch.inj_sweeps) ? 1 : 0;
        a.inj_row = ch.inj_pos + sw;
        k_sim_hy<<<dim3((net.nH + 127) / 128, ch.n_chains), 128, 0, st>>>(net, ch, a);
        launches++;
        if (!net.const_params) { k_sim_t<<<ch.n_chains, ST_NT, 0, st>>>(net, ch, a); launches++; }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return launches;
}

void launch_sim_conditionals(const DevNet &net, const DevChains &ch, uint32_t c, double *out, cudaStream_t st)
{
    const uint32_t n = net.nH + n_t_of(net);
    k_sim_conditionals<<<(n + 127) / 128, 128, 0, st>>>(net, ch, c, out);
}

void launch_sim_export_draws(const DevNet &net, uint32_t gchain, uint32_t sweep0, uint32_t n,
                             double *flat_u, double *beta_v, double *exp_v, cudaStream_t st)
{
    const uint32_t per = 2 * net.nH > net.nX ? 2 * net.nH : net.nX;
    size_t total = (size_t) n * per;
    uint32_t blocks = (uint32_t) ((total + 127) / 128);
    if (blocks > 4096) blocks = 4096;
    if (blocks == 0) blocks = 1;
    k_sim_export_draws<<<blocks, 128, 0, st>>>(net, gchain, sweep0, n, flat_u, beta_v, exp_v);
}

}  // namespace gbn
