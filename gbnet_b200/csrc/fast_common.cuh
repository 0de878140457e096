// Helpers shared by the fast-path kernels (sweep_fast.cu: networks of any size, state in HBM; sweep_small.cu:
// networks that fit one SM's shared memory, one persistent CTA per chain).
#pragma once

#include <cstdio>
#include <cstdint>

#include "device_model.cuh"
#include "device_math.cuh"
#include "topology.hpp"

// Bounds checks of every data-dependent index, compiled in by the "chk" build variant
// (GBN_LIB_VARIANT=chk; compute-sanitizer is closed on this pool): a violation traps the kernel.
#ifdef GBN_BOUNDS
#define GBN_CHECK(cond) do { if (!(cond)) { printf("GBN_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define GBN_CHECK(cond) do { } while (0)
#endif

namespace gbn {

namespace {

struct SweepArgs {
    uint32_t sweep;                  // absolute sweep index (Philox counter word)
    uint32_t stat_n;                 // statistics count AFTER this sweep
    uint32_t use_inj;                // 1: injected draws, row inj_row
    uint32_t inj_row;                // sweep row inside the injection arrays
    uint32_t chain0;                 // first local chain of this launch (chains are split over streams)
    uint32_t x_adapt;                // X kernel: adaptive window on (GBN_X_ADAPT=2 switches it off)
    uint32_t x1_force;               // single-chain X kernel: resolve every x1_force-th TF in fp64 whatever its draw (tests of
                                     // the resolve path: GBN_X1=3 every TF, GBN_X1=4 every eighth)
    uint32_t rk0[10], rk1[10];       // Philox round keys (seed words bumped by the Weyl constants): as kernel
                                     // parameters they are constant-bank operands of the round's XOR
};

inline SweepArgs sweep_args(const DevNet &net)
{
    SweepArgs a{};
    uint32_t k0 = (uint32_t) net.seed, k1 = (uint32_t) (net.seed >> 32);
    for (int r = 0; r < 10; r++) { a.rk0[r] = k0; a.rk1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    a.x_adapt = env_u32("GBN_X_ADAPT", 1) != 2;
    a.x1_force = env_u32("GBN_X1", 1) == 3 ? 1 : (env_u32("GBN_X1", 1) == 4 ? 8 : 0);
    return a;
}

// philox4x32_10 (device_math.cuh) with the key schedule taken from the kernel parameters
__device__ __forceinline__ u32x4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const SweepArgs &a)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ a.rk0[r];
        const uint32_t n2 = hi0 ^ c3 ^ a.rk1[r];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return u32x4{ c0, c1, c2, c3 };
}

// n += k under a predicate as ONE predicated instruction (the compiler's own choice for `if (p) n += k;` is an
// add and a select; for a double multiply it insists on the select, so there is no mul_if)
__device__ __forceinline__ void add_if(uint32_t &a, uint32_t k, bool p)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q add.u32 %0, %0, %1;\n\t}" : "+r"(a) : "r"(k), "r"((int) p));
}

// log(gsl_ran_beta_pdf(x, a, b)) (Beta.cpp:50,82,83 under the log of Beta.cpp:69-78).  Inside (0, 1) the density is
// exp(lnB + (a-1) log x + (b-1) log1p(-x)) and the reference takes the log of that exp: here the exponent is
// used directly (one exp and one log less per density - four densities per T update - at a difference of an
// ulp or two in the log-ratio, the size of the libm differences the fast path already has; the strict kernel
// keeps the reference's form).  Outside and at the end points the density's own special cases decide.
__device__ __forceinline__ double beta_logpdf(double x, double a, double b, double lnB)
{
    if (x > 0. && x < 1.) return lnB + log(x) * (a - 1) + log1p(-x) * (b - 1);
    return log(beta_pdf(x, a, b, lnB));
}

// Beta::metropolis_hastings_with_prior_proposal (Beta.cpp:60-91) on log densities; returns the value the node keeps
__device__ __forceinline__ double mh_decide_log(double prev, double proposal, double prev_loglik, double prop_loglik,
                                                double prior_mean, double sample_dx, double a, double b, double lnB,
                                                bool have_expo, double expo,
                                                uint64_t seed, uint32_t gchain, uint32_t sweep, uint32_t k)
{
    if (!(prop_loglik > -INFINITY)) return prev;
    if (prev_loglik > -INFINITY) {
        // log(q01 / q10): a zero density on either side keeps the reference's quotient (0 / q, q / 0, 0 / 0)
        const double lq01 = beta_logpdf(prior_mean - sample_dx, a, b, lnB), lq10 = beta_logpdf(prior_mean + sample_dx, a, b, lnB);
        const double log_q = (lq01 > -INFINITY && lq10 > -INFINITY) ? lq01 - lq10 : log(exp(lq01) / exp(lq10));
        const double logratio = prop_loglik - prev_loglik + log_q;
        bool accept = logratio >= 0.;
        if (!accept) {
            if (!have_expo) expo = draw_exp(seed, gchain, sweep, k);
            accept = logratio > -expo;
        }
        if (!accept) return prev;
    }
    return proposal;
}

// One update of a T node whose blanket is its Beta prior alone (Beta::metropolis_hastings_with_prior_proposal,
// Beta.cpp:60-91): prev -> the value the node keeps, given the prior draw of the proposal.  With the four density
// arguments inside (0, 1) every density is positive and the log of the acceptance ratio is
//   (a-1) log[prop q01 / (prev q10)] + (b-1) log[(1-prop)(1-q01) / ((1-prev)(1-q10))]
// (lnB cancels): two logs and two divisions instead of the eight logs of four separate log-densities, at a difference
// of a few ulps in the log-ratio (the size of the libm differences the fast path already has; the strict kernel keeps
// the reference's form).  A proposal outside [0, 1], or a reverse-move argument q01 outside it, is rejected without
// evaluating anything.  Anything else - an argument at an end point, or so small that a product could underflow, or a
// current value with density zero - takes the general path with the densities' own special cases.
__device__ __forceinline__ double mh_prior_step(double prev, double draw, double mean, double al, double be, double lnB,
                                                bool have_expo, double expo_in,
                                                uint64_t seed, uint32_t gchain, uint32_t sweep, uint32_t k)
{
    const double dx = draw - mean;
    const double prop = prev + dx;
    const double q01 = mean - dx, q10 = mean + dx;
    constexpr double LO = 1e-150;
    const bool in_prev = prev > LO && prev < 1., in_prop = prop > LO && prop < 1., in_q10 = q10 > LO && q10 < 1.;
    // the proposal has density zero (gsl_ran_beta_pdf outside [0, 1]): kept at prev, Beta.cpp:66-68
    if (in_prev && (prop < 0. || prop > 1.)) return prev;
    // q01 = 0 with the other three densities positive: log(q01 / q10) = -inf, the proposal is rejected (Beta.cpp:80-90)
    if (in_prev && in_prop && in_q10 && (q01 < 0. || q01 > 1.)) return prev;
    if (in_prev && in_prop && in_q10 && q01 > LO && q01 < 1.) {
        const double r1 = (prop * q01) / (prev * q10);
        const double r2 = ((1. - prop) * (1. - q01)) / ((1. - prev) * (1. - q10));
        const double logratio = log(r1) * (al - 1) + log(r2) * (be - 1);
        bool accept = logratio >= 0.;
        if (!accept) {
            const double expo = have_expo ? expo_in : draw_exp(seed, gchain, sweep, k);
            accept = logratio > -expo;
        }
        return accept ? prop : prev;
    }
    const double prev_ll = (0. + beta_logpdf(prev, al, be, lnB)) + 0.;
    const double prop_ll = (0. + beta_logpdf(prop, al, be, lnB)) + 0.;
    return mh_decide_log(prev, prop, prev_ll, prop_ll, mean, dx, al, be, lnB, have_expo, expo_in, seed, gchain, sweep, k);
}

// 2-bit S code at position pos of a packed array (16 codes per word)
__device__ __forceinline__ uint32_t code_at(const uint32_t *w, uint32_t pos) { return (w[pos >> 4] >> (2 * (pos & 15))) & 3u; }

// m * 2^e -> mantissa in [0.5, 1) and the exponent moved into e (frexp).  Likelihood products are
// positive normal numbers, so the exponent field is rewritten directly; zero / subnormal / inf / nan
// take the library path.
// out of line on purpose: inlined, the library path's instructions are issued (predicated off) by every call
__device__ __noinline__ double frexp_rare(double m, int *ex) { return frexp(m, ex); }

__device__ __forceinline__ void normalise(double &m, int &e)
{
    const int hi = __double2hiint(m);
    const int ef = (hi >> 20) & 0x7ff;
    if (ef != 0 && ef != 0x7ff) {
        e += ef - 1022;
        m = __hiloint2double((hi & 0x800fffff) | (1022 << 20), __double2loint(m));
    } else {
        int ex;
        m = frexp_rare(m, &ex);
        e += ex;
    }
}

// a * 2^e for e < -900 with ONE rounding (what exp() of the reference delivers near the bottom of the double range:
// subnormal, or zero): the first product is exact, the second rounds into the subnormal range
__device__ __forceinline__ double scale_down(double a, int e)
{
    if (e < -1080) return 0.;                                               // a < 2: below half the smallest subnormal
    const double t = a * __hiloint2double((1023 + (e + 1000)) << 20, 0);    // e + 1000 in [-80, 40]: exact
    return t * __hiloint2double((1023 - 1000) << 20, 0);
}

// the two outcome likelihoods a * 2^ea and b * 2^eb (a, b in [0.25, 2)) brought to one scale.  Only
// their ratio matters for the draw, except that the reference falls back to the prior when both
// underflow (Multinomial.cpp:78-85) and rounds subnormals: near the bottom of the double range the
// exact ldexp is used, elsewhere the smaller one is scaled by 2^(difference) with a bit pattern.
__device__ __forceinline__ void common_scale(double a, int ea, double b, int eb, double &la, double &lb)
{
    const int emax = ea > eb ? ea : eb;
    if (emax < -960) { la = scale_down(a, ea); lb = scale_down(b, eb); return; }
    const int da = ea - emax, db = eb - emax;                       // one is 0, the other <= 0
    la = da < -1000 ? 0. : a * __hiloint2double((1023 + da) << 20, 0);
    lb = db < -1000 ? 0. : b * __hiloint2double((1023 + db) << 20, 0);
}

}  // namespace

}  // namespace gbn
