// FAST sweep path: the same Gibbs sweep as sweep_strict.cu (same node order, same keyed draws,
// same conditionals up to floating-point rounding), organised for throughput.
//
// What changes against the reference's loop (GraphBase.cpp:22-27 -> Multinomial.cpp:56-108 ->
// XNode.cpp:16-22 / SNode.cpp:15-21 -> HNode.cpp:33-53 -> HNodeORNOR.cpp:40-106):
//
//  * Per target and chain a 32-byte cache {pr0, pr2, zc, L} holds the two parent products, the
//    (1-z)^n term and the current likelihood.  The reference rescans all parents of a target for
//    each of the 3 latent states and each candidate value (15-21 x sum indeg^2 parent visits per
//    sweep); here an X candidate costs one cache line and ~20 flops, an S candidate ~15 flops.
//  * X phase (k_fast_x, one CTA per chain): X updates are sequential in the reference and do not
//    commute when TFs share a target.  Each warp of the CTA evaluates one TF of a WINDOW of
//    consecutive TFs against the same state; draws are taken for all of them; the window commits
//    up to and including the first TF whose value changed, applies that flip to the cache, and the
//    next window starts behind it.  A TF's conditional is only used if no earlier TF of its window
//    flipped, i.e. exactly when it equals the conditional the sequential sweep would have seen, so
//    the chain is the sequential chain, not an approximation of it.  X values rarely change
//    (prior 0.01 / 0.99), so a 16-wide window advances ~14 TFs per round.
//  * The sum of child log-likelihoods becomes a product with explicit exponent tracking (no log,
//    no exp per child); the underflow fallback of Multinomial.cpp:78-85 is kept: the scaled
//    product underflows to zero where exp(sum log) does.
//  * T phase (k_fast_t, thread per (chain, TF)): T nodes that do not listen to their children are
//    mutually independent (TNode.cpp:19-22).  Also refreshes FX = 1 - t*x and 1/FX.
//  * S phase (k_fast_s, thread per (chain, target)): S updates of different targets commute; a
//    thread walks its target's edges in input order with leave-one-out products (multiply by the
//    precomputed reciprocal factor), no transcendental: exp(log p + log L) = p * L.
//    Targets are assigned to lanes in descending in-degree order so a warp's lanes finish together.
//
// Not handled here (the model layer routes these to the strict kernel): T nodes that listen to
// their children, duplicate (TF, target) edges.
#include "kernels.cuh"
#include "device_math.cuh"

namespace gbn {

namespace {

constexpr int FX_NT = 512;           // X kernel: 16 warps = 16-TF speculative window
constexpr int FS_NT = 128;           // S kernel
constexpr int FT_NT = 128;           // T kernel

struct SweepArgs {
    uint32_t sweep;                  // absolute sweep index (Philox counter word)
    uint32_t stat_n;                 // statistics count AFTER this sweep
    uint32_t use_inj;                // 1: injected draws, row inj_row
    uint32_t inj_row;                // sweep row inside the injection arrays
};

__device__ __forceinline__ void normalise(double &m, int &e)
{
    int ex;
    m = frexp(m, &ex);
    e += ex;
}

__device__ __forceinline__ TargetCache load_cache(const TargetCache *p)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = q[0], b = q[1];
    return TargetCache{ a.x, a.y, b.x, b.y };
}

__device__ __forceinline__ void store_cache(TargetCache *p, double pr0, double pr2, double zc, double L)
{
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(pr0, pr2);
    q[1] = make_double2(zc, L);
}

// ---------------------------------------------------------------------------------- X phase
template <int NT>
__global__ void __launch_bounds__(NT, 2) k_fast_x(DevNet net, DevChains ch, SweepArgs a)
{
    constexpr int NW = NT / 32;
    extern __shared__ uint8_t sX[];              // [nX]
    __shared__ int s_dec[NW];

    const uint32_t c = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nX = net.nX, E = net.E, nH = net.nH;
    const uint32_t d = chain_dataset(net, c);
    const uint32_t gchain = chain_global_id(net, c);
    uint8_t *gX = ch.X + (size_t) c * nX;
    const double *gT = ch.T + (size_t) c * nX;
    const uint8_t *gStf = ch.Stf + (size_t) c * E;
    TargetCache *tcache = ch.tcache + (size_t) c * nH;
    uint32_t *cX1 = ch.cX1 + (size_t) c * nX;
    const uint8_t *gy = net.y + (size_t) d * nH;
    const double yp[3] = { net.yprob[3 * d], net.yprob[3 * d + 1], net.yprob[3 * d + 2] };
    const double *inj_flat = a.use_inj ? ch.inj_flat + ((size_t) c * ch.inj_sweeps + a.inj_row) * (nX + E) : nullptr;

    for (uint32_t k = tid; k < nX; k += NT) sX[k] = gX[k];
    __syncthreads();

    uint32_t k0 = 0;
    while (k0 < nX) {
        const uint32_t kw = k0 + warp;
        int dec = 0;                              // bit 0: drawn value, bit 1: value changed
        double fac = 1.;
        uint32_t cbeg = 0, cend = 0;
        if (kw < nX) {
            const uint32_t xcur = sX[kw];
            const double g = 1. - gT[kw];         // (1 - t*x) at x = 1   (HNodeORNOR.cpp:59)
            fac = xcur ? 1. / g : g;              // factor that turns the product into the flipped one
            cbeg = net.tf_ptr[kw]; cend = net.tf_ptr[kw + 1];
            double mc = 1., mf = 1.;
            int ec = 0, ef = 0, cnt = 0;
            for (uint32_t cc = cbeg + lane; cc < cend; cc += 32) {
                const uint32_t i = net.tf_child[cc];
                const uint32_t s = gStf[cc];
                const TargetCache tc = load_cache(tcache + i);
                double Lf = tc.L;
                if (s != 1) {
                    const uint32_t y = gy[i];
                    const double col[3] = { net.ynoise[0][y], net.ynoise[1][y], net.ynoise[2][y] };
                    const double pr0 = s == 0 ? tc.pr0 * fac : tc.pr0;
                    const double pr2 = s == 2 ? tc.pr2 * fac : tc.pr2;
                    Lf = lik_from_products(pr0, pr2, tc.zc, yp, col);
                }
                mc *= tc.L;
                mf *= Lf;
                if (++cnt == 16) { normalise(mc, ec); normalise(mf, ef); cnt = 0; }
            }
            normalise(mc, ec);
            normalise(mf, ef);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mc *= __shfl_xor_sync(0xffffffffu, mc, o);
                mf *= __shfl_xor_sync(0xffffffffu, mf, o);
                ec += __shfl_xor_sync(0xffffffffu, ec, o);
                ef += __shfl_xor_sync(0xffffffffu, ef, o);
            }
            const double *prob = net.xprob[net.x_prior_active[kw]];
            const double lik_cur = ldexp(prob[xcur] * mc, ec);
            const double lik_flip = ldexp(prob[1 - xcur] * mf, ef);
            const double l0 = xcur ? lik_flip : lik_cur, l1 = xcur ? lik_cur : lik_flip;
            double cum0 = 0. + l0, cum1 = cum0 + l1;
            if (!(cum1 > 0.)) { cum0 = 0. + prob[0]; cum1 = cum0 + prob[1]; }     // Multinomial.cpp:78-85
            const double u = a.use_inj ? inj_flat[kw] : keyed_u(net.seed, gchain, a.sweep, kw, KIND_X);
            const double dice = 0. * (1 - u) + cum1 * u;
            const uint32_t nx = dice < cum0 ? 0u : (dice < cum1 ? 1u : xcur);
            dec = (int) nx | ((nx != xcur) ? 2 : 0);
        }
        if (lane == 0) s_dec[warp] = dec;
        __syncthreads();
        uint32_t first = NW;
#pragma unroll
        for (int w = NW - 1; w >= 0; w--) if (s_dec[w] & 2) first = (uint32_t) w;
        const uint32_t ncommit = first < NW ? first + 1 : NW;
        if (warp < ncommit && kw < nX) {
            if (lane == 0) { sX[kw] = (uint8_t) (dec & 1); cX1[kw] += (uint32_t) (dec & 1); }
            if (dec & 2) {
                // apply the flip of X_kw to the cache of each child that depends on it
                for (uint32_t cc = cbeg + lane; cc < cend; cc += 32) {
                    const uint32_t s = gStf[cc];
                    if (s == 1) continue;
                    const uint32_t i = net.tf_child[cc];
                    const TargetCache tc = load_cache(tcache + i);
                    const uint32_t y = gy[i];
                    const double col[3] = { net.ynoise[0][y], net.ynoise[1][y], net.ynoise[2][y] };
                    const double pr0 = s == 0 ? tc.pr0 * fac : tc.pr0;
                    const double pr2 = s == 2 ? tc.pr2 * fac : tc.pr2;
                    store_cache(tcache + i, pr0, pr2, tc.zc, lik_from_products(pr0, pr2, tc.zc, yp, col));
                }
            }
        }
        __syncthreads();
        k0 += ncommit;
    }
    for (uint32_t k = tid; k < nX; k += NT) gX[k] = sX[k];
}

// ---------------------------------------------------------------------------------- T phase
// Beta::sample (Beta.cpp:60-104) for T nodes that do not listen to their children: the blanket is
// the Beta prior alone.  Also publishes FX = 1 - t*x (HNodeORNOR.cpp:59) and its reciprocal for
// the S phase.
__global__ void __launch_bounds__(FT_NT) k_fast_t(DevNet net, DevChains ch, SweepArgs a, int update_t)
{
    const uint32_t nX = net.nX;
    const uint32_t c = blockIdx.y;
    const uint32_t k = blockIdx.x * FT_NT + threadIdx.x;
    if (k >= nX) return;
    const size_t o = (size_t) c * nX + k;
    double t = ch.T[o];
    if (update_t) {
        const uint32_t gchain = chain_global_id(net, c);
        const double al = net.t_a[k], be = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
        const size_t row = ((size_t) c * ch.inj_sweeps + a.inj_row) * nX + k;
        const double prev = t;
        const double prev_ll = (0. + log(beta_pdf(prev, al, be, lnB))) + 0.;
        const double draw = a.use_inj ? ch.inj_beta[row] : draw_beta(net.seed, gchain, a.sweep, k, al, be);
        const double dx = draw - mean;
        const double prop = prev + dx;
        const double prop_ll = (0. + log(beta_pdf(prop, al, be, lnB))) + 0.;
        t = mh_decide(prev, prop, prev_ll, prop_ll, mean, dx, al, be, lnB,
                      a.use_inj != 0, a.use_inj ? ch.inj_exp[row] : 0., net.seed, gchain, a.sweep, k);
        ch.T[o] = t;
        double m = ch.tMean[o], m2 = ch.tM2[o];
        welford_add(m, m2, t, (double) a.stat_n);
        ch.tMean[o] = m; ch.tM2[o] = m2;
    }
    const double fx = 1. - t * (double) ch.X[o];
    ch.FX[o] = fx;
    ch.FXinv[o] = 1. / fx;
}

// ---------------------------------------------------------------------------------- S phase
// mode 0: sweep (draw every S of the target, count, refresh the cache)
// mode 1: refresh only (rebuild the cache from X, T, S; no draws, no counts)
template <int MODE>
__global__ void __launch_bounds__(FS_NT) k_fast_s(DevNet net, DevChains ch, SweepArgs a)
{
    const uint32_t nX = net.nX, E = net.E, nH = net.nH;
    const uint32_t c = blockIdx.y;
    const uint32_t idx = blockIdx.x * FS_NT + threadIdx.x;
    if (idx >= nH) return;
    const uint32_t i = net.trg_order[idx];
    const uint32_t d = chain_dataset(net, c);
    const uint32_t gchain = chain_global_id(net, c);
    uint8_t *gS = ch.S + (size_t) c * E;
    uint8_t *gStf = ch.Stf + (size_t) c * E;
    const double *FX = ch.FX + (size_t) c * nX;
    const double *FXinv = ch.FXinv + (size_t) c * nX;
    uint32_t *cS0 = ch.cS0 + (size_t) c * E;
    uint32_t *cS2 = ch.cS2 + (size_t) c * E;
    const double yp[3] = { net.yprob[3 * d], net.yprob[3 * d + 1], net.yprob[3 * d + 2] };
    const uint32_t y = net.y[(size_t) d * nH + i];
    const double col[3] = { net.ynoise[0][y], net.ynoise[1][y], net.ynoise[2][y] };
    const uint32_t cls = (y != 1) ? 0 : 1;                     // ZY for DE targets, ZN otherwise
    const double z = cls ? net.z_n : net.z_y;
    const double rz = 1. / z;
    const double *zct = net.zctab + (size_t) cls * (net.max_indeg + 2);
    const uint32_t q0 = net.trg_ptr[i], q1 = net.trg_ptr[i + 1];

    // products over the parents in append_parent order (HNodeORNOR.cpp:55-62, 72-81)
    double P0 = 1., P2 = 1.;
    uint32_t n = 0;
    for (uint32_t q = q0; q < q1; q++) {
        const uint32_t s = gS[q];
        if (s != 1) {
            const double f = FX[net.par_tf[q]] * z;
            n++;
            if (s == 0) P0 *= f; else P2 *= f;
        }
    }
    double Lfin = lik_from_products(P0, P2, zct[n], yp, col);
    if (MODE == 0) {
        const double *inj_flat = a.use_inj ? ch.inj_flat + ((size_t) c * ch.inj_sweeps + a.inj_row) * (nX + E) + nX : nullptr;
        uint32_t blk = 0xffffffffu;
        u32x4 rnd = { 0, 0, 0, 0 };
        for (uint32_t q = q0; q < q1; q++) {
            const uint32_t kk = net.par_tf[q];
            const uint32_t s = gS[q];
            const double *sp = net.sprob[net.mor_idx[q]];
            const double f = FX[kk] * z;
            const double finv = FXinv[kk] * rz;
            const double R0 = s == 0 ? P0 * finv : P0;          // products without this edge
            const double R2 = s == 2 ? P2 * finv : P2;
            const uint32_t nR = n - (s != 1);
            const double zc0 = zct[nR], zc1 = zct[nR + 1];
            double L[3];
            L[0] = lik_from_products(R0 * f, R2, zc1, yp, col);
            L[1] = lik_from_products(R0, R2, zc0, yp, col);
            L[2] = lik_from_products(R0, R2 * f, zc1, yp, col);
            // exp(log prior + log L) of Multinomial.cpp:66-75 as a product
            double cum0 = 0. + sp[0] * L[0];
            double cum1 = cum0 + sp[1] * L[1];
            double cum2 = cum1 + sp[2] * L[2];
            if (!(cum2 > 0.)) { cum0 = 0. + sp[0]; cum1 = cum0 + sp[1]; cum2 = cum1 + sp[2]; }
            double u;
            if (a.use_inj) u = inj_flat[net.par_edge[q]];
            else {
                if ((q >> 2) != blk) {
                    blk = q >> 2;
                    rnd = philox4x32_10(blk, a.sweep, gchain, KIND_S, (uint32_t) net.seed, (uint32_t) (net.seed >> 32));
                }
                u = unit_from_word(pick_word(rnd, q & 3));
            }
            const double dice = 0. * (1 - u) + cum2 * u;
            const uint32_t ns = dice < cum0 ? 0u : (dice < cum1 ? 1u : (dice < cum2 ? 2u : s));
            if (ns != s) {
                gS[q] = (uint8_t) ns;
                gStf[net.tfpos_of_slot[q]] = (uint8_t) ns;
                P0 = ns == 0 ? R0 * f : R0;
                P2 = ns == 2 ? R2 * f : R2;
                n = nR + (ns != 1);
            }
            cS0[q] += (ns == 0);
            cS2[q] += (ns == 2);
            Lfin = L[ns];
        }
    }
    store_cache(ch.tcache + (size_t) c * nH + i, P0, P2, zct[n], Lfin);
}

// Stf from S (refresh only)
__global__ void k_fast_stf(DevNet net, DevChains ch)
{
    const uint32_t c = blockIdx.y;
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= net.E) return;
    ch.Stf[(size_t) c * net.E + net.tfpos_of_slot[q]] = ch.S[(size_t) c * net.E + q];
}

}  // namespace

bool fast_supported(const DevNet &net, const char **why)
{
    if (net.Eu != net.E) { *why = "duplicate (TF, target) edges"; return false; }
    if (net.listen && !net.const_params) { *why = "T nodes listen to their children"; return false; }
    if ((size_t) net.nX + 64 > 200 * 1024) { *why = "n_tf too large for the X window kernel"; return false; }
    return true;
}

static int check(cudaError_t e, const char **err)
{
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 0;
}

int launch_fast_refresh(const DevNet &net, const DevChains &ch, cudaStream_t st, const char **err)
{
    if (ch.n_chains == 0) return 0;
    SweepArgs a{ 0, 0, 0, 0 };
    k_fast_stf<<<dim3((net.E + 255) / 256, ch.n_chains), 256, 0, st>>>(net, ch);
    k_fast_t<<<dim3((net.nX + FT_NT - 1) / FT_NT, ch.n_chains), FT_NT, 0, st>>>(net, ch, a, 0);
    k_fast_s<1><<<dim3((net.nH + FS_NT - 1) / FS_NT, ch.n_chains), FS_NT, 0, st>>>(net, ch, a);
    if (check(cudaGetLastError(), err)) return -1;
    return 3;
}

int launch_sweep_fast(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0,
                      uint32_t stat_n0, cudaStream_t st, const char **err)
{
    if (n_sweeps == 0 || ch.n_chains == 0) return 0;
    static bool attr_set = false;
    const size_t smem_x = net.nX;
    if (!attr_set || smem_x > 48 * 1024) {
        if (check(cudaFuncSetAttribute(k_fast_x<FX_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) (smem_x > 48 * 1024 ? smem_x : 48 * 1024)), err)) return -1;
        attr_set = true;
    }
    int launches = 0;
    for (uint32_t sw = 0; sw < n_sweeps; sw++) {
        SweepArgs a;
        a.sweep = sweep0 + sw;
        a.stat_n = stat_n0 + sw + 1;
        a.use_inj = (ch.inj_flat != nullptr && (ch.inj_pos + sw) < ch.inj_sweeps) ? 1 : 0;
        a.inj_row = ch.inj_pos + sw;
        k_fast_x<FX_NT><<<ch.n_chains, FX_NT, smem_x, st>>>(net, ch, a);
        k_fast_t<<<dim3((net.nX + FT_NT - 1) / FT_NT, ch.n_chains), FT_NT, 0, st>>>(net, ch, a, net.const_params ? 0 : 1);
        k_fast_s<0><<<dim3((net.nH + FS_NT - 1) / FS_NT, ch.n_chains), FS_NT, 0, st>>>(net, ch, a);
        launches += 3;
    }
    if (check(cudaGetLastError(), err)) return -1;
    return launches;
}

}  // namespace gbn
