// STRICT sweep kernel: one CTA per chain, the reference's sweep order and expressions.
//
// Replaces GraphBase::sample (GraphBase.cpp:22-27) for all chains at once (the OpenMP chain pool
// of ModelBase.cpp:131-136 becomes the grid).  Inside a chain the node order of random_nodes is
// kept: X[sorted id] -> T[sorted id] -> S[edge input order]; the parallelism used is exact:
//   * one X (or listening T) update: the sum over the TF's children is spread over the CTA;
//   * T updates that do not listen to children are independent of each other (TNode.cpp:19-22);
//   * S updates of different targets commute (SNode.cpp:15-21 reads one target only), so threads
//     own targets and walk each target's edges in input order.
// Every likelihood is recomputed from X, T, S by a scan over the target's parents in
// append_parent order (HNodeORNOR.cpp:40-106), so duplicate edges, listening T nodes and
// degenerate values all behave as in the reference.  This is the kernel parity is pinned on; the
// throughput path is sweep_fast.cu.
#include "kernels.cuh"
#include "device_math.cuh"

namespace gbn {

namespace {

template <int NT>
__global__ void __launch_bounds__(NT) k_sweep_strict(DevNet net, DevChains ch, uint32_t n_sweeps,
                                                     uint32_t sweep0, uint32_t stat_n0)
{
    extern __shared__ double smem[];
    double *sT = smem;                                   // [nX]
    double *scratch = sT + net.nX;                       // [2 * NT / 32]
    uint8_t *sX = (uint8_t *) (scratch + 2 * (NT / 32)); // [nX]

    const uint32_t c = blockIdx.x;
    const uint32_t tid = threadIdx.x;
    const uint32_t nX = net.nX, E = net.E, nH = net.nH;
    const uint32_t d = chain_dataset(net, c);
    const uint32_t gchain = chain_global_id(net, c);
    uint8_t *gX = ch.X + (size_t) c * nX;
    double *gT = ch.T + (size_t) c * nX;
    uint8_t *gS = ch.S + (size_t) c * net.Ep;
    uint32_t *cX1 = ch.cX1 + (size_t) c * nX;
    double *tMean = ch.tMean + (size_t) c * nX;
    double *tM2 = ch.tM2 + (size_t) c * nX;
    uint2 *cS = ch.cS + (size_t) c * net.Ep;
    const uint8_t *gy = net.y + (size_t) d * nH;
    const double *yp = net.yprob + 3 * (size_t) d;

    for (uint32_t k = tid; k < nX; k += NT) { sX[k] = gX[k]; sT[k] = gT[k]; }
    __syncthreads();

    for (uint32_t sw = 0; sw < n_sweeps; sw++) {
        const uint32_t sweep = sweep0 + sw;
        const double n_after = (double) (stat_n0 + sw + 1);
        const bool inj = ch.inj_flat != nullptr && (ch.inj_pos + sw) < ch.inj_sweeps;
        const size_t inj_row = (size_t) c * ch.inj_sweeps + ch.inj_pos + sw;
        const double *inj_flat = inj ? ch.inj_flat + inj_row * (nX + E) : nullptr;
        const double *inj_beta = inj ? ch.inj_beta + inj_row * nX : nullptr;
        const double *inj_exp = inj ? ch.inj_exp + inj_row * nX : nullptr;

        // ------------------------------------------------------------ X phase (sequential over TFs)
        for (uint32_t k = 0; k < nX; k++) {
            double acc0 = 0., acc1 = 0.;
            const uint32_t c1 = net.tf_ptr[k + 1];
            for (uint32_t cc = net.tf_ptr[k] + tid; cc < c1; cc += NT) {
                const uint32_t i = net.tf_child[cc];
                const uint32_t y = gy[i];
                const double z = (y != 1) ? net.z_y : net.z_n;
                ParentScan p0, p1;
                const uint32_t deg = net.deg[i];
                uint32_t q = slot0(net, i);
                for (uint32_t j = 0; j < deg; j++, q += 32) {
                    const uint32_t kk = net.par_tf[q];
                    const uint32_t s = gS[q];
                    const double t = sT[kk];
                    const uint32_t x = sX[kk];
                    p0.add(t, kk == k ? 0u : x, s, z);
                    p1.add(t, kk == k ? 1u : x, s, z);
                }
                const double col[3] = { net.ynoise[0][y], net.ynoise[1][y], net.ynoise[2][y] };
                acc0 += log(p0.likelihood(yp, col));
                acc1 += log(p1.likelihood(yp, col));
            }
            block_sum2(acc0, acc1, scratch);
            if (tid == 0) {
                const double *prob = net.xprob[net.x_prior_active[k]];
                double llik[2];
                llik[0] = (0. + log(prob[0])) + acc0;        // RVNode.cpp:94-106
                llik[1] = (0. + log(prob[1])) + acc1;
                const double u = inj ? inj_flat[k] : keyed_u(net.seed, gchain, sweep, k, KIND_X);
                const uint32_t nx = multinomial_draw<2>(llik, prob, u, sX[k]);
                sX[k] = (uint8_t) nx;
                cX1[k] += nx;                                // Multinomial.cpp:101-105 as a count
            }
            __syncthreads();
        }

        // ------------------------------------------------------------ T phase
        if (!net.const_params) {
            if (!net.listen) {
                for (uint32_t k = tid; k < nX; k += NT) {
                    const double a = net.t_a[k], b = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
                    const double prev = sT[k];
                    const double prev_ll = (0. + log(beta_pdf(prev, a, b, lnB))) + 0.;
                    const double draw = inj ? inj_beta[k] : draw_beta(net.seed, gchain, sweep, k, a, b);
                    const double dx = draw - mean;
                    const double prop = prev + dx;
                    const double prop_ll = (0. + log(beta_pdf(prop, a, b, lnB))) + 0.;
                    const double nt = mh_decide(prev, prop, prev_ll, prop_ll, mean, dx, a, b, lnB,
                                                inj, inj ? inj_exp[k] : 0., net.seed, gchain, sweep, k);
                    sT[k] = nt;
                    double m = tMean[k], m2 = tM2[k];
                    welford_add(m, m2, nt, n_after);          // Beta.cpp:98-101
                    tMean[k] = m; tM2[k] = m2;
                }
                __syncthreads();
            } else {
                for (uint32_t k = 0; k < nX; k++) {
                    const double a = net.t_a[k], b = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
                    const double prev = sT[k];
                    // every thread evaluates the same draw (uniform control flow, no broadcast needed)
                    const double draw = inj ? inj_beta[k] : draw_beta(net.seed, gchain, sweep, k, a, b);
                    const double dx = draw - mean;
                    const double prop = prev + dx;
                    double acc0 = 0., acc1 = 0.;
                    const uint32_t c1 = net.tf_ptr[k + 1];
                    for (uint32_t cc = net.tf_ptr[k] + tid; cc < c1; cc += NT) {
                        const uint32_t i = net.tf_child[cc];
                        const uint32_t y = gy[i];
                        const double z = (y != 1) ? net.z_y : net.z_n;
                        ParentScan p0, p1;
                        const uint32_t deg = net.deg[i];
                        uint32_t q = slot0(net, i);
                        for (uint32_t j = 0; j < deg; j++, q += 32) {
                            const uint32_t kk = net.par_tf[q];
                            const uint32_t s = gS[q];
                            const uint32_t x = sX[kk];
                            const double t = sT[kk];
                            p0.add(kk == k ? prev : t, x, s, z);
                            p1.add(kk == k ? prop : t, x, s, z);
                        }
                        const double col[3] = { net.ynoise[0][y], net.ynoise[1][y], net.ynoise[2][y] };
                        acc0 += log(p0.likelihood(yp, col));
                        acc1 += log(p1.likelihood(yp, col));
                    }
                    block_sum2(acc0, acc1, scratch);
                    const double prev_ll = (0. + log(beta_pdf(prev, a, b, lnB))) + acc0;
                    const double prop_ll = (0. + log(beta_pdf(prop, a, b, lnB))) + acc1;
                    const double nt = mh_decide(prev, prop, prev_ll, prop_ll, mean, dx, a, b, lnB,
                                                inj, inj ? inj_exp[k] : 0., net.seed, gchain, sweep, k);
                    if (tid == 0) {
                        sT[k] = nt;
                        double m = tMean[k], m2 = tM2[k];
                        welford_add(m, m2, nt, n_after);
                        tMean[k] = m; tM2[k] = m2;
                    }
                    __syncthreads();
                }
            }
        }

        // ------------------------------------------------------------ S phase (threads own targets)
        for (uint32_t i = tid; i < nH; i += NT) {
            const uint32_t y = gy[i];
            const double z = (y != 1) ? net.z_y : net.z_n;
            const double col[3] = { net.ynoise[0][y], net.ynoise[1][y], net.ynoise[2][y] };
            const uint32_t q0 = slot0(net, i), deg = net.deg[i];
            const uint32_t ref_i = net.ref_of_dt[i];
            for (uint32_t j = 0; j < deg; j++) {
                const uint32_t q = q0 + 32 * j;
                const uint32_t cur = gS[q];
                const uint32_t mi = net.mor_idx[q];
                double llik[3];
#pragma unroll
                for (uint32_t v = 0; v < 3; v++) {
                    ParentScan ps;
                    for (uint32_t jj = 0; jj < deg; jj++) {
                        const uint32_t qq = q0 + 32 * jj;
                        const uint32_t kk = net.par_tf[qq];
                        ps.add(sT[kk], sX[kk], jj == j ? v : (uint32_t) gS[qq], z);
                    }
                    llik[v] = (0. + log(net.sprob[mi][v])) + log(ps.likelihood(yp, col));
                }
                const double u = inj ? inj_flat[nX + net.par_edge[q]] : keyed_s(net.seed, gchain, sweep, ref_i, j);
                const uint32_t ns = multinomial_draw<3>(llik, net.sprob[mi], u, cur);
                gS[q] = (uint8_t) ns;
                { uint2 cs = cS[q]; cs.x += (ns == 0); cs.y += (ns == 2); cS[q] = cs; }
            }
        }
        __syncthreads();
    }

    for (uint32_t k = tid; k < nX; k += NT) { gX[k] = sX[k]; gT[k] = sT[k]; }
}

constexpr int STRICT_NT = 256;

}  // namespace

size_t strict_smem_bytes(const DevNet &net)
{
    return sizeof(double) * (net.nX + 2 * (STRICT_NT / 32)) + net.nX;
}

int launch_sweep_strict(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0,
                        uint32_t stat_n0, cudaStream_t st, const char **err)
{
    if (n_sweeps == 0 || ch.n_chains == 0) return 0;
    size_t smem = strict_smem_bytes(net);
    if (smem > 227 * 1024) { *err = "strict kernel: X/T state of one chain does not fit in shared memory (n_tf too large)"; return -1; }
    cudaError_t e = cudaFuncSetAttribute(k_sweep_strict<STRICT_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    k_sweep_strict<STRICT_NT><<<ch.n_chains, STRICT_NT, smem, st>>>(net, ch, n_sweeps, sweep0, stat_n0);
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

}  // namespace gbn
