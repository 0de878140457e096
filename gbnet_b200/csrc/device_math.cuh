// Device-side arithmetic shared by every kernel: keyed Philox draws, the Beta/Exp samplers of
// the T update, gsl_ran_beta_pdf, and the OR-NOR target likelihood.
//
// Compiled with -fmad=false: the reference's x86-64 build contains no FMA, so
// a*b+c must round twice here as well for products/sums to match it bit for bit.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace gbn {

// ---------------------------------------------------------------- keyed draws
// Replaces the per-chain MT19937 of GraphBase.cpp:8-9 / ModelORNOR.cpp:24 (seeded from
// time(NULL), not reproducible) by counter-based Philox4x32-10 (Salmon et al., SC'11):
//   key     = 64-bit model seed
//   counter = (node group, sweep, global chain id, stream kind [| block << 8])
// so any (chain, sweep, node) draw can be regenerated independently of every other.
enum : uint32_t { KIND_X = 0, KIND_S = 1, KIND_TBETA = 2, KIND_TEXP = 3, KIND_INIT = 4, KIND_H = 5, KIND_Y = 6 };

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1)
{
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t) (p0 >> 32), lo0 = (uint32_t) p0;
        uint32_t hi1 = (uint32_t) (p1 >> 32), lo1 = (uint32_t) p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return u32x4{ c0, c1, c2, c3 };
}

__device__ __forceinline__ uint32_t pick_word(const u32x4 &v, uint32_t i)
{
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// u = w / 2^32 in [0,1): the resolution gsl_rng_uniform has for mt19937
__device__ __forceinline__ double unit_from_word(uint32_t w) { return (double) w * (1.0 / 4294967296.0); }

// one 32-bit uniform per discrete node; four neighbouring nodes share one Philox block
__device__ __forceinline__ double keyed_u(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t node, uint32_t kind)
{
    u32x4 r = philox4x32_10(node >> 2, sweep, chain, kind, (uint32_t) seed, (uint32_t) (seed >> 32));
    return unit_from_word(pick_word(r, node & 3));
}

// S nodes: the uniform of the j-th parent edge (append_parent order) of target i; four consecutive
// edges of one target share one Philox block, so a thread walking a target's edges refreshes its
// block every fourth step, in step with the other lanes of its warp
__device__ __forceinline__ u32x4 keyed_s_block(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t target, uint32_t j)
{
    return philox4x32_10(target, sweep, chain, KIND_S | ((j >> 2) << 8), (uint32_t) seed, (uint32_t) (seed >> 32));
}
__device__ __forceinline__ double keyed_s(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t target, uint32_t j)
{
    return unit_from_word(pick_word(keyed_s_block(seed, chain, sweep, target, j), j & 3));
}

// sequential word stream private to one T update
struct WordStream {
    uint32_t k0, k1, tf, sweep, chain, block, have;
    u32x4 buf;
    __device__ __forceinline__ WordStream(uint64_t seed, uint32_t chain_, uint32_t sweep_, uint32_t tf_)
        : k0((uint32_t) seed), k1((uint32_t) (seed >> 32)), tf(tf_), sweep(sweep_), chain(chain_), block(0), have(0) {}
    __device__ __forceinline__ double uniform_pos()
    {
        for (;;) {
            if (have == 0) {
                buf = philox4x32_10(tf, sweep, chain, KIND_TBETA | (block << 8), k0, k1);
                block++;
                have = 4;
            }
            uint32_t w = pick_word(buf, 4 - have);
            have--;
            if (w != 0) return unit_from_word(w);
        }
    }
    // polar Box-Muller, one deviate per call
    __device__ __forceinline__ double gaussian()
    {
        double x, y, r2;
        do {
            x = -1.0 + 2.0 * uniform_pos();
            y = -1.0 + 2.0 * uniform_pos();
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        return y * sqrt(-2.0 * log(r2) / r2);
    }
    // Marsaglia & Tsang (2000); a < 1 through Gamma(a+1) * U^(1/a)
    __device__ __forceinline__ double gamma(double a)
    {
        double boost = 1.0;
        if (a < 1.0) {
            double u = uniform_pos();
            boost = pow(u, 1.0 / a);
            a = a + 1.0;
        }
        double d = a - 1.0 / 3.0;
        double c = (1.0 / 3.0) / sqrt(d);
        double x, v, u;
        for (;;) {
            do {
                x = gaussian();
                v = 1.0 + c * x;
            } while (v <= 0.0);
            v = v * v * v;
            u = uniform_pos();
            if (u < 1.0 - 0.0331 * (x * x) * (x * x)) break;
            if (log(u) < 0.5 * (x * x) + d * (1.0 - v + log(v))) break;
        }
        return boost * (d * v);
    }
};

// replaces gsl_ran_beta (Beta.cpp:56)
__device__ __forceinline__ double draw_beta(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t tf, double a, double b)
{
    WordStream ws(seed, chain, sweep, tf);
    double x1 = ws.gamma(a);
    double x2 = ws.gamma(b);
    return x1 / (x1 + x2);
}

// replaces gsl_ran_exponential(rng, 1.0) (Beta.cpp:87)
__device__ __forceinline__ double draw_exp(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t tf)
{
    u32x4 r = philox4x32_10(tf, sweep, chain, KIND_TEXP, (uint32_t) seed, (uint32_t) (seed >> 32));
    return -log1p(-unit_from_word(r.x));
}

// ---------------------------------------------------------------- gsl_ran_beta_pdf
// GSL 2.x randist/beta.c (call sites Beta.cpp:50,82,83); lnB = lgamma(a+b)-lgamma(a)-lgamma(b)
// is evaluated on the host (topology.hpp)
__device__ __forceinline__ double beta_pdf(double x, double a, double b, double lnB)
{
    if (x < 0 || x > 1) return 0;
    if (x == 0.0 || x == 1.0) {
        if (a > 1.0 && b > 1.0) return 0.0;
        return exp(lnB) * pow(x, a - 1) * pow(1 - x, b - 1);
    }
    return exp(lnB + log(x) * (a - 1) + log1p(-x) * (b - 1));
}

// ---------------------------------------------------------------- OR-NOR target likelihood
// Running products of one scan over a target's parents, kept exactly as
// HNodeORNOR::get_model_likelihood keeps them (HNodeORNOR.cpp:40-106):
//   pr0 = prod_{s=0} f, pr2 = prod_{s=2} f, pr1 = prod_{s!=1} f (interleaved, parent order),
//   zc  = prod_{s!=1} (1 - z),      f = (1 - t*x) * z
struct ParentScan {
    double pr0, pr2, pr1, zc;
    __device__ __forceinline__ ParentScan() : pr0(1.), pr2(1.), pr1(1.), zc(1.) {}
    __device__ __forceinline__ void add(double t, uint32_t x, uint32_t s, double z)
    {
        if (s != 1) {
            double f = (1. - t * (double) x) * z;
            zc *= (1. - z);
            pr1 *= f;
            if (s == 0) pr0 *= f; else pr2 *= f;
        }
    }
    // P_model(h) (HNodeORNOR.cpp:64,82,96) and HNode::get_own_likelihood (HNode.cpp:33-53):
    // L = sum_h P(h) * YNOISE[h][y]
    __device__ __forceinline__ double likelihood(const double *yprob, const double *ynoise_col /* YNOISE[.][y] */) const
    {
        double p0 = (1. - pr0) * (1. - zc) + zc * yprob[0];
        double p1 = pr1 * (1. - zc) + zc * yprob[1];
        double p2 = (pr0 - pr2 * pr0) * (1. - zc) + zc * yprob[2];
        double l = 0.;
        l += p0 * ynoise_col[0];
        l += p1 * ynoise_col[1];
        l += p2 * ynoise_col[2];
        return l;
    }
    // HNodeORNOR::get_model_likelihood (HNodeORNOR.cpp:40-106) for h = 0, 1, 2
    __device__ __forceinline__ void model_probs(const double *yprob, double p[3]) const
    {
        p[0] = (1. - pr0) * (1. - zc) + zc * yprob[0];
        p[1] = pr1 * (1. - zc) + zc * yprob[1];
        p[2] = (pr0 - pr2 * pr0) * (1. - zc) + zc * yprob[2];
    }
};

// Multinomial::compute_outcome_likelihood + the draw of Multinomial::sample
// (Multinomial.cpp:56-99): lik = exp(llik), prior fallback when the sum is not > 0,
// dice = gsl_ran_flat(0, cum[last]) = 0*(1-u) + cum[last]*u, first v with dice < cum[v]
template <int N>
__device__ __forceinline__ uint32_t multinomial_draw(const double (&llik)[N], const double *prob, double u, uint32_t current)
{
    double cum[N], sum_lik = 0.;
#pragma unroll
    for (int i = 0; i < N; i++) { sum_lik += exp(llik[i]); cum[i] = sum_lik; }
    if (!(sum_lik > 0.)) {
        sum_lik = 0.;
#pragma unroll
        for (int i = 0; i < N; i++) { sum_lik += prob[i]; cum[i] = sum_lik; }
    }
    double dice = 0. * (1 - u) + cum[N - 1] * u;
    uint32_t out = current;
    bool done = false;
#pragma unroll
    for (int i = 0; i < N; i++)
        if (!done && dice < cum[i]) { out = (uint32_t) i; done = true; }
    return out;
}

}  // namespace gbn

// ======================================================================================
// Helpers shared by the kernels (appended: generic target likelihood, reductions)
// ======================================================================================
#include "device_model.cuh"

namespace gbn {

// HNode::get_own_likelihood (HNode.cpp:33-53) of DEVICE target i: one scan over the parents in
// append_parent order.  xf(kk, q), tf(kk, q), sf(q) give X, T of parent TF kk and S of slot q,
// so callers can override one node (candidate evaluation) without touching memory.
template <class XF, class TF, class SF>
__device__ __forceinline__ double target_likelihood(const DevNet &net, uint32_t dataset, uint32_t i,
                                                    XF xf, TF tf, SF sf)
{
    const uint32_t y = net.y[(size_t) dataset * net.nH + i];
    const double z = (y != 1) ? net.z_y : net.z_n;                    // GraphORNOR.cpp:257-260
    ParentScan ps;
    const uint32_t deg = net.deg[i];
    uint32_t q = slot0(net, i);
    for (uint32_t j = 0; j < deg; j++, q += 32) {
        const uint32_t kk = net.par_tf[q];
        ps.add(tf(kk, q), xf(kk, q), sf(q), z);
    }
    const double *yp = net.yprob + 3 * (size_t) dataset;
    double col[3] = { net.ynoise[0][y], net.ynoise[1][y], net.ynoise[2][y] };
    return ps.likelihood(yp, col);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block-wide sums of two values; every thread receives the totals.
// scratch must hold 2 * (blockDim.x / 32) doubles.  Contains two __syncthreads().
__device__ __forceinline__ void block_sum2(double &a, double &b, double *scratch)
{
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();                       // scratch free from the previous use
    if (lane == 0) { scratch[2 * w] = a; scratch[2 * w + 1] = b; }
    __syncthreads();
    double ta = 0., tb = 0.;
    for (uint32_t i = 0; i < nw; i++) { ta += scratch[2 * i]; tb += scratch[2 * i + 1]; }
    a = ta; b = tb;
}

// gsl_rstat_add (GSL rstat/rstat.c) for the mean / M2 pair; n is the count AFTER this add
__device__ __forceinline__ void welford_add(double &mean, double &M2, double x, double n)
{
    double delta = x - mean;
    double delta_n = delta / n;
    double term1 = delta * delta_n * (n - 1.0);
    mean += delta_n;
    M2 += term1;
}

// Beta::metropolis_hastings_with_prior_proposal (Beta.cpp:60-91) given the two blanket
// log-likelihoods; returns the value the node keeps.
__device__ __forceinline__ double mh_decide(double prev, double proposal, double prev_loglik, double prop_loglik,
                                            double prior_mean, double sample_dx, double a, double b, double lnB,
                                            bool have_expo, double expo,
                                            uint64_t seed, uint32_t gchain, uint32_t sweep, uint32_t k)
{
    if (!(prop_loglik > -INFINITY)) return prev;
    if (prev_loglik > -INFINITY) {
        double q10 = beta_pdf(prior_mean + sample_dx, a, b, lnB);
        double q01 = beta_pdf(prior_mean - sample_dx, a, b, lnB);
        double log_q = log(q01 / q10);
        double logratio = prop_loglik - prev_loglik + log_q;
        bool accept = logratio >= 0.;
        if (!accept) {
            if (!have_expo) expo = draw_exp(seed, gchain, sweep, k);
            accept = logratio > -expo;
        }
        if (!accept) return prev;
    }
    return proposal;
}

}  // namespace gbn
