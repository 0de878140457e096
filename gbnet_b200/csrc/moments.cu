// K5 — cross-chain moments: Gelman-Rubin (ModelBase.cpp:47-112) and posterior means / sdevs
// (ModelBase.cpp:160-209) from the per-chain counters, without ever materialising per-chain
// gsl_rstat workspaces.
//
// Per chain k and stat j the reference keeps mean_kj and variance_kj (RVNode.cpp:115-123).  For the
// discrete nodes these follow exactly from the outcome counts: mean = c/N and
// variance = sum (x - mean)^2 / (N-1) = c (N-c) / N / (N-1).  For T they are the Welford pair.
// The cross-chain reduction distributes over ranks.  Discrete nodes (X, S: all but nX of the
// statistics) use exact integer sums of the counts in one pass (k_count_sums / k_expand_counts: one
// u64 all-reduce); the continuous T nodes use two floating-point passes over 3 nX and 2 nX doubles:
//   pass 1  sum_k mean_kj, sum_k variance_kj                     (-> all-reduce, sum)
//   pass 2  sum_k (mean_kj - mu_j)^2                             (-> all-reduce, sum)
//   final   B = N * pass2 / (M-1), W = pass1.var / M, Vhat, R    (every rank, redundantly)
// The posterior getters skip GLOBAL chain 0 exactly as the reference does (ModelBase.cpp:172,197).
//
// Work items are "nodes in slot order": X_k (2 stats), T_k (1 stat), S at (padded) slot q (3 stats), so that
// for every chain the counter reads are coalesced; results are written at the stat / node index
// of the reference's random_nodes order (S in edge input order).
#include "kernels.cuh"
#include "device_math.cuh"

namespace gbn {

namespace {

struct NodeRef {
    uint32_t kind;      // 0 X, 1 T, 2 S
    uint32_t idx;       // k or slot q
    uint32_t n_stats;   // 2, 1, 3
    uint32_t stat0;     // first stat index (reference order)
    uint32_t node;      // node index (reference order)
};

__device__ __forceinline__ NodeRef node_ref(const DevNet &net, uint32_t n)
{
    const uint32_t nT = net.const_params ? 0 : net.nX;
    NodeRef r;
    if (net.sim) {                                          // H_i | Y_i | T  (sim_kernels.cu)
        if (n < 2 * net.nH) {
            const uint32_t isY = n >= net.nH, dt = n - (isY ? net.nH : 0), ref = net.ref_of_dt[dt];
            r.kind = 3 + isY; r.idx = dt; r.n_stats = 3; r.stat0 = 6 * ref + 3 * isY; r.node = 2 * ref + isY;
        } else {
            r.kind = 1; r.idx = n - 2 * net.nH; r.n_stats = 1; r.stat0 = 6 * net.nH + r.idx; r.node = n;
        }
        return r;
    }
    if (n < net.nX) { r.kind = 0; r.idx = n; r.n_stats = 2; r.stat0 = 2 * n; r.node = n; }
    else if (n < net.nX + nT) { r.kind = 1; r.idx = n - net.nX; r.n_stats = 1; r.stat0 = 2 * net.nX + r.idx; r.node = n; }
    else {
        r.kind = 2; r.idx = n - net.nX - nT;
        const uint32_t e = net.par_edge[r.idx];
        r.n_stats = e == 0xffffffffu ? 0 : 3;              // pad slot: nothing to report
        r.stat0 = 2 * net.nX + nT + 3 * e;
        r.node = net.nX + nT + e;
    }
    return r;
}

__device__ __forceinline__ void count_moment(double c, double N, double &mean, double &var)
{
    mean = c / N;
    var = N > 1. ? (c * (N - c) / N) / (N - 1.) : 0.;
}

// mean / variance of every stat of node r in local chain c
__device__ __forceinline__ void node_moments(const DevNet &net, const DevChains &ch, uint32_t c, const NodeRef &r,
                                             double N, double mean[3], double var[3])
{
    if (r.kind == 0) {
        double c1 = (double) ch.cX1[(size_t) c * net.nX + r.idx];
        count_moment(N - c1, N, mean[0], var[0]);
        count_moment(c1, N, mean[1], var[1]);
    } else if (r.kind == 1) {
        mean[0] = ch.tMean[(size_t) c * net.nX + r.idx];
        var[0] = N > 1. ? ch.tM2[(size_t) c * net.nX + r.idx] / (N - 1.) : 0.;
    } else {
        const uint2 cs = r.kind == 2 ? ch.cS[(size_t) c * net.Ep + r.idx]
                       : (r.kind == 3 ? ch.cH : ch.cY)[(size_t) c * net.nH + r.idx];
        double c0 = (double) cs.x;
        double c2 = (double) cs.y;
        count_moment(c0, N, mean[0], var[0]);
        count_moment(N - c0 - c2, N, mean[1], var[1]);
        count_moment(c2, N, mean[2], var[2]);
    }
}

// T nodes (one statistic each; the discrete nodes go through k_count_sums).  A block is 32 nodes x MS_SL
// chain slices: lane = node (consecutive addresses), each slice sums every MS_SL-th chain, and the slices
// are added up in a fixed order, so the result does not depend on the launch.
constexpr int MS_SL = 16;
// tsum[d][k] = { sum_k mean, sum_k var, sum of mean over local chains whose GLOBAL index is >= 1 }
__global__ void __launch_bounds__(32 * MS_SL) k_moment_sums(DevNet net, DevChains ch, double N, double *tsum, uint32_t n0, uint32_t n1)
{
    __shared__ double s_part[3][MS_SL][32];
    const uint32_t d = blockIdx.y;
    const uint32_t lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const uint32_t n = n0 + blockIdx.x * 32 + lane;
    const bool live = n < n1;
    const NodeRef r = node_ref(net, live ? n : n0);
    double sm = 0., sv = 0., ps = 0.;
    if (live)
        for (uint32_t l = sl; l < net.n_graphs_local; l += MS_SL) {
            double mean[3], var[3];
            node_moments(net, ch, d * net.n_graphs_local + l, r, N, mean, var);
            sm += mean[0]; sv += var[0];
            if ((net.chain_offset + l) >= 1) ps += mean[0];
        }
    s_part[0][sl][lane] = sm; s_part[1][sl][lane] = sv; s_part[2][sl][lane] = ps;
    __syncthreads();
    if (sl == 0 && live) {
        sm = sv = ps = 0.;
        for (int q = 0; q < MS_SL; q++) { sm += s_part[0][q][lane]; sv += s_part[1][q][lane]; ps += s_part[2][q][lane]; }
        double *o = tsum + 3 * ((size_t) d * (n1 - n0) + (n - n0));
        o[0] = sm; o[1] = sv; o[2] = ps;
    }
}

// tdev[d][k] = { sum_k (mean_k - mu)^2, the same over global chains >= 1 around their own mean }
__global__ void __launch_bounds__(32 * MS_SL) k_moment_devs(DevNet net, DevChains ch, double N, const double *tsum, double *tdev,
                                                            uint32_t n0, uint32_t n1)
{
    __shared__ double s_part[2][MS_SL][32];
    const uint32_t d = blockIdx.y;
    const uint32_t lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const uint32_t n = n0 + blockIdx.x * 32 + lane;
    const bool live = n < n1;
    const NodeRef r = node_ref(net, live ? n : n0);
    const double M = (double) net.n_graphs;
    const size_t o = (size_t) d * (n1 - n0) + ((live ? n : n0) - n0);
    const double mu = tsum[3 * o] / M;
    const double pmu = M > 1. ? tsum[3 * o + 2] / (M - 1.) : 0.;
    double dv = 0., pd = 0.;
    if (live)
        for (uint32_t l = sl; l < net.n_graphs_local; l += MS_SL) {
            double mean[3], var[3];
            node_moments(net, ch, d * net.n_graphs_local + l, r, N, mean, var);
            const double a = mean[0] - mu;
            dv += a * a;
            if ((net.chain_offset + l) >= 1) { const double b = mean[0] - pmu; pd += b * b; }
        }
    s_part[0][sl][lane] = dv; s_part[1][sl][lane] = pd;
    __syncthreads();
    if (sl == 0 && live) {
        dv = pd = 0.;
        for (int q = 0; q < MS_SL; q++) { dv += s_part[0][q][lane]; pd += s_part[1][q][lane]; }
        tdev[2 * o] = dv;
        tdev[2 * o + 1] = pd;
    }
}

// ---- discrete nodes: exact integer moments -------------------------------------------------------
// For an outcome with per-chain counts c_k (k = 1..M chains, N sweeps): sum_k mean_k = S1 / N,
// sum_k var_k = (N S1 - S2) / (N (N-1)), sum_k (mean_k - mu)^2 = (M S2 - S1^2) / (M N^2), with
// S1 = sum c_k, S2 = sum c_k^2 - integers, so ONE pass over the counters and ONE all-reduce give the
// between- and within-chain terms exactly (no second pass, no fp64 division per chain and slot).
//
// What crosses the ranks (the one exchange step of the path) is kept to what the
// STORED counters determine: an X node stores the count of outcome 1, an S (H, Y) node the counts a, b of
// outcomes 0 and 2; the complementary outcome follows from c = N - (stored).  Per dataset:
//     X_k :  { S1, S2, z }                      S1 = sum c, S2 = sum c^2 over the local chains
//     S_e :  { A1, A2, B1, B2, AB, z }          AB = sum a b (the cross term the complement's S2 needs)
// z carries the counts of GLOBAL chain 0 (zero on every rank but the one that owns it; a | b << 32 for S):
// the posterior getters skip that chain (ModelBase.cpp:172,197), and "all chains minus chain 0" replaces a
// second set of sums.  C3: 6 x 301,038 + 3 x 1,500 u64 = 14.5 MB per check (round 1: 28.9 MB + 36 MB of doubles).
constexpr uint32_t IS_X = 3, IS_S = 6;
__host__ __device__ inline size_t isum_words(const DevNet &n) { return n.sim ? (size_t) IS_S * 2 * n.nH : (size_t) IS_X * n.nX + (size_t) IS_S * n.E; }

// with_pending: the fast path's per-batch byte counters (ch.dC) have not been folded into the totals: add them here
// (the fold rewrites 2.5 MB of totals per chain; between folds the R-hat check only reads)
// totals_zero: the totals of the fast path's S nodes are logically zero and were never written (capi.cu, cs_zero)
__global__ void k_count_sums(DevNet net, DevChains ch, unsigned long long *isums, int with_pending, int totals_zero)
{
    const uint32_t d = blockIdx.y;
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;          // X_k for w < nX, then slots
    if (w >= (net.sim ? 2 * net.nH : net.nX + net.Ep)) return;
    unsigned long long *out = isums + (size_t) d * isum_words(net);
    if (!net.sim && w < net.nX) {
        unsigned long long s1 = 0, s2 = 0, z = 0;
        for (uint32_t l = 0; l < net.n_graphs_local; l++) {
            const unsigned long long c1 = ch.cX1[(size_t) (d * net.n_graphs_local + l) * net.nX + w];
            s1 += c1; s2 += c1 * c1;
            if (net.chain_offset + l == 0) z = c1;
        }
        out[IS_X * w] = s1; out[IS_X * w + 1] = s2; out[IS_X * w + 2] = z;
        return;
    }
    const uint2 *cnt;
    const uint2 *dC = nullptr;                                          // pending counts of this slot: byte sh of entry didx
    size_t stride, idx, o, didx = 0;
    uint32_t sh = 0;
    if (net.sim) {                                                      // H_dt for w < nH, then Y_dt
        const uint32_t isY = w >= net.nH, dt = w - (isY ? net.nH : 0);
        cnt = isY ? ch.cY : ch.cH; stride = net.nH; idx = dt;
        o = (size_t) IS_S * (2 * net.ref_of_dt[dt] + isY);
    } else {
        const uint32_t q = w - net.nX, e = net.par_edge[q];
        if (e == 0xffffffffu) return;                                   // pad slot
        cnt = ch.cS; stride = net.Ep; idx = q;
        o = (size_t) IS_X * net.nX + (size_t) IS_S * e;
        if (with_pending) {                                             // slot q = gbase[g] + 32 j + lane (topology.hpp)
            const uint32_t g = net.slot_trg[q] >> 5, j = (q - net.gbase[g]) >> 5;
            dC = ch.dC; didx = net.dbase[g] + 32 * (j >> 2) + (q & 31); sh = 8 * (j & 3);
        }
    }
    unsigned long long a1 = 0, a2 = 0, b1 = 0, b2 = 0, ab = 0, z = 0;
    for (uint32_t l = 0; l < net.n_graphs_local; l++) {
        const uint2 cs = totals_zero ? make_uint2(0u, 0u) : cnt[(size_t) (d * net.n_graphs_local + l) * stride + idx];
        unsigned long long a = cs.x, b = cs.y;
        if (dC) {
            const uint2 dv = dC[(size_t) (d * net.n_graphs_local + l) * net.nD + didx];
            a += (dv.x >> sh) & 0xffu; b += (dv.y >> sh) & 0xffu;
        }
        a1 += a; a2 += a * a; b1 += b; b2 += b * b; ab += a * b;
        if (net.chain_offset + l == 0) z = a | (b << 32);
    }
    out[o] = a1; out[o + 1] = a2; out[o + 2] = b1; out[o + 3] = b2; out[o + 4] = ab; out[o + 5] = z;
}

// globally reduced sums -> per statistic: sums = {sum mean, sum var}, psums = sum of means over chains >= 1,
// dev / pdev = sums of squared deviations of the chain means (all chains / chains >= 1)
__global__ void k_expand_counts(DevNet net, double Nd, const unsigned long long *isums, const double *tsum, const double *tdev,
                                double *sums, double *psums, double *dev, double *pdev)
{
    typedef unsigned __int128 u128;
    const uint32_t n_stats = n_stats_of(net);
    const uint32_t nT = net.const_params ? 0 : net.nX;
    const uint32_t d = blockIdx.y;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_stats) return;
    const size_t o = (size_t) d * n_stats + j;
    const uint32_t t0 = net.sim ? 6 * net.nH : 2 * net.nX;
    if (j >= t0 && j < t0 + nT) {                                       // T statistics: the two floating-point passes
        const size_t q = (size_t) d * nT + (j - t0);
        sums[2 * o] = tsum[3 * q]; sums[2 * o + 1] = tsum[3 * q + 1]; psums[o] = tsum[3 * q + 2];
        dev[o] = tdev[2 * q]; pdev[o] = tdev[2 * q + 1];
        return;
    }
    const unsigned long long *in = isums + (size_t) d * isum_words(net);
    const u128 N = (u128) (unsigned long long) Nd, M = net.n_graphs;
    u128 S1, S2, c0;                                                    // over all chains; c0 = the count in global chain 0
    if (!net.sim && j < 2 * net.nX) {
        const unsigned long long *a = in + (size_t) IS_X * (j >> 1);
        S1 = a[0]; S2 = a[1]; c0 = a[2];
        if ((j & 1) == 0) { S2 = M * N * N - 2 * N * S1 + S2; S1 = M * N - S1; c0 = N - c0; }       // outcome 0: c = N - c1
    } else {
        const uint32_t js = net.sim ? j : j - 2 * net.nX - nT, node = js / 3, v = js - 3 * node;
        const unsigned long long *a = in + (net.sim ? 0 : (size_t) IS_X * net.nX) + (size_t) IS_S * node;
        const u128 A1 = a[0], A2 = a[1], B1 = a[2], B2 = a[3], AB = a[4], za = a[5] & 0xffffffffull, zb = a[5] >> 32;
        if (v == 0) { S1 = A1; S2 = A2; c0 = za; }
        else if (v == 2) { S1 = B1; S2 = B2; c0 = zb; }
        else { S1 = M * N - A1 - B1; S2 = M * N * N - 2 * N * (A1 + B1) + A2 + B2 + 2 * AB; c0 = N - za - zb; }   // c = N - a - b
    }
    const u128 P1 = S1 - c0, P2 = S2 - c0 * c0;                         // chains 1 .. M-1
    const double M_d = (double) net.n_graphs, n_d = M_d - 1.;
    sums[2 * o] = (double) S1 / Nd;
    sums[2 * o + 1] = Nd > 1. ? (double) (N * S1 - S2) / (Nd * (Nd - 1.)) : 0.;
    psums[o] = (double) P1 / Nd;
    dev[o] = (double) (M * S2 - S1 * S1) / (M_d * Nd * Nd);
    pdev[o] = n_d >= 1. ? (double) ((M - 1) * P2 - P1 * P1) / (n_d * Nd * Nd) : 0.;
}

// ModelBase::get_gelman_rubin (ModelBase.cpp:72-92) from the globally reduced sums
__global__ void k_gelman_rubin(DevNet net, double N, const double *sums, const double *dev, uint32_t d0, uint32_t nd,
                               double *gr_out, unsigned long long *max_out)
{
    const uint32_t n_nodes = n_nodes_of(net), n_stats = n_stats_of(net);
    const uint32_t d = d0 + blockIdx.y;
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    double max_gr = 0.;
    const NodeRef r = (n < n_work_of(net)) ? node_ref(net, n) : NodeRef{ 0, 0, 0, 0, 0 };
    if (r.n_stats > 0 && blockIdx.y < nd) {
        const double M = (double) net.n_graphs;
        for (uint32_t j = 0; j < r.n_stats; j++) {
            size_t o = (size_t) d * n_stats + r.stat0 + j;
            double B = N * (M > 1. ? dev[o] / (M - 1.) : 0.);
            double W = sums[2 * o + 1] / M;
            double Vhat = W * (N - 1.) / N + B * (M + 1.) / (M * N);
            double gr;
            if (W > 0. && Vhat > 0.) gr = sqrt((N + 3.) / (N + 1.) * Vhat / W);
            else if (Vhat > 0.) gr = INFINITY;
            else gr = 1.;
            max_gr = fmax(gr, max_gr);
        }
        if (gr_out) gr_out[(size_t) blockIdx.y * n_nodes + r.node] = max_gr;
    }
    // block maximum -> one atomic per block (non-negative doubles order like their bit patterns)
    for (int o = 16; o > 0; o >>= 1) max_gr = fmax(max_gr, __shfl_xor_sync(0xffffffffu, max_gr, o));
    __shared__ double wmax[32];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = max_gr;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.;
        for (uint32_t w = 0; w < (blockDim.x >> 5); w++) m = fmax(m, wmax[w]);
        atomicMax(max_out + d, (unsigned long long) __double_as_longlong(m));
    }
}

__global__ void k_node_moments(DevNet net, DevChains ch, uint32_t c, double N, double *mean_out, double *var_out)
{
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_work_of(net)) return;
    const NodeRef r = node_ref(net, n);
    if (r.n_stats == 0) return;
    double mean[3], var[3];
    node_moments(net, ch, c, r, N, mean, var);
    for (uint32_t j = 0; j < r.n_stats; j++) { mean_out[r.stat0 + j] = mean[j]; var_out[r.stat0 + j] = var[j]; }
}

}  // namespace

void launch_moment_sums(const DevNet &net, const DevChains &ch, double N, double *tsum, cudaStream_t st)
{
    const uint32_t nT = net.const_params ? 0 : net.nX;                  // T nodes only: the rest is k_count_sums
    if (nT == 0) return;
    dim3 grid((nT + 31) / 32, net.n_datasets);
    const uint32_t n0 = net.sim ? 2 * net.nH : net.nX;                  // first T work item
    k_moment_sums<<<grid, 32 * MS_SL, 0, st>>>(net, ch, N, tsum, n0, n0 + nT);
}

size_t moment_isum_words(const DevNet &net) { return isum_words(net); }

void launch_count_sums(const DevNet &net, const DevChains &ch, unsigned long long *isums, cudaStream_t st, bool with_pending,
                       bool totals_zero)
{
    dim3 grid(((net.sim ? 2 * net.nH : net.nX + net.Ep) + 127) / 128, net.n_datasets);
    const bool fast = !net.sim && ch.dC != nullptr;
    k_count_sums<<<grid, 128, 0, st>>>(net, ch, isums, with_pending && fast ? 1 : 0, totals_zero && fast ? 1 : 0);
}

void launch_expand_counts(const DevNet &net, double N, const unsigned long long *isums, const double *tsum, const double *tdev,
                          double *sums, double *psums, double *dev, double *pdev, cudaStream_t st)
{
    dim3 grid((n_stats_of(net) + 127) / 128, net.n_datasets);
    k_expand_counts<<<grid, 128, 0, st>>>(net, N, isums, tsum, tdev, sums, psums, dev, pdev);
}

void launch_moment_devs(const DevNet &net, const DevChains &ch, double N, const double *tsum, double *tdev, cudaStream_t st)
{
    const uint32_t nT = net.const_params ? 0 : net.nX;                  // T nodes only
    if (nT == 0) return;
    dim3 grid((nT + 31) / 32, net.n_datasets);
    const uint32_t n0 = net.sim ? 2 * net.nH : net.nX;
    k_moment_devs<<<grid, 32 * MS_SL, 0, st>>>(net, ch, N, tsum, tdev, n0, n0 + nT);
}

void launch_gelman_rubin(const DevNet &net, double N, const double *sums, const double *dev,
                         uint32_t dataset, double *gr_out, double *max_out, cudaStream_t st)
{
    // dataset == UINT32_MAX: all datasets, maxima only
    uint32_t d0 = dataset == UINT32_MAX ? 0 : dataset;
    uint32_t nd = dataset == UINT32_MAX ? net.n_datasets : 1;
    dim3 grid((n_work_of(net) + 127) / 128, nd);
    k_gelman_rubin<<<grid, 128, 0, st>>>(net, N, sums, dev, d0, nd, gr_out, (unsigned long long *) max_out);
}

void launch_node_moments(const DevNet &net, const DevChains &ch, uint32_t c, double N, double *mean, double *var, cudaStream_t st)
{
    k_node_moments<<<(n_work_of(net) + 127) / 128, 128, 0, st>>>(net, ch, c, N, mean, var);
}

}  // namespace gbn
