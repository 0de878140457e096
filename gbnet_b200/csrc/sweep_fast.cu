// FAST sweep path: the same Gibbs sweep as sweep_strict.cu (same node order, same keyed draws,
// same conditionals up to floating-point rounding), organised for throughput.
//
// What changes against the reference's loop (GraphBase.cpp:22-27 -> Multinomial.cpp:56-108 ->
// XNode.cpp:16-22 / SNode.cpp:15-21 -> HNode.cpp:33-53 -> HNodeORNOR.cpp:40-106):
//
//  * The reference rescans all parents of a target for each of the 3 latent states and each
//    candidate value (15-21 x sum indeg^2 parent visits per sweep).  Here the target likelihood is
//    written as  L = c0[y,n] + pr0 * omz[n] * (a_y + pr2 * b_y)  (pr0 / pr2 = products over the
//    parents with S = 0 / S = 2, n = number of parents with S != 1; algebra at k_fast_s) and a
//    16-byte per-(target, chain) cache {beta, gamma} + a 16-bit table index carry it from the S
//    phase to the next X phase: an X candidate costs one gather and ~6 flops, an S candidate ~5.
//  * X phase (k_fast_x, one CTA per PAIR of chains): X updates are sequential in the reference and do
//    not commute when TFs share a target.  Each half-warp (a quarter or a whole warp for other degree
//    profiles) evaluates one TF of a WINDOW of 64 consecutive TFs against the same state; draws are
//    taken for all of them; the window commits up to and including the first TF whose value changed
//    (in either chain), applies that flip to the cache, and the next window starts behind it.  A TF's
//    conditional is only used if no earlier TF of its window flipped, i.e. exactly when it equals the
//    conditional the sequential sweep would have seen, so the chain is the sequential chain, not an
//    approximation of it.  X values rarely change (measured at 1,500 TFs: 27 rounds per sweep, 5 of
//    them ended by a flip).  The cache is interleaved by chain pair, so the lanes of one child read one line.
//  * The sum of child log-likelihoods becomes a product with explicit exponent tracking (no log,
//    no exp per child); the underflow fallback of Multinomial.cpp:78-85 is kept.
//  * T phase (k_fast_t, thread per (chain, TF)): T nodes that do not listen to their children are
//    mutually independent (TNode.cpp:19-22) and need nothing from the sweep's X phase, so they are drawn
//    beside it into an alternate buffer; k_fast_fx then publishes FXp = {1 - t*x, 1 / (1 - t*x)}.
//  * S phase (k_fast_s, one warp per (group of 32 targets, chain)): S updates of different targets
//    commute; a lane walks its target's edges in input order with leave-one-out values of
//    A = a pr0 and Bq = b pr0 pr2 (multiply by the precomputed reciprocal factor), no transcendental:
//    exp(log p + log L) = p * L.
//    Device targets are sorted by in-degree and slots are stored transposed (topology.hpp), so the
//    lanes of a warp finish together and every S / parent-id / counter access is coalesced.
//  * Chains are independent, so the model layer runs disjoint chain groups on separate stream pairs
//    (capi.cu): one group's latency-bound X kernel overlaps another group's S kernel and its own T draws.
//
//  * Listening T nodes (TNode.cpp:15-25): only ACTIVE TFs have a child-dependent blanket; they are
//    walked sequentially per chain by k_fast_tl using the same cache, the rest stay parallel.
//
// Not handled here (the model layer routes these to the strict kernel): duplicate (TF, target) edges.
#include <cstdio>
#include "kernels.cuh"
#include "device_math.cuh"
#include <type_traits>
#include <utility>

#include "topology.hpp"
#include "fast_common.cuh"

#ifndef GBN_S2_PRIOR_SMEM
#define GBN_S2_PRIOR_SMEM 1   // k_fast_s2: prior rows from shared memory (0: from the constant bank)
#endif
#ifndef GBN_S2_DC_PREFETCH
#define GBN_S2_DC_PREFETCH 1  // k_fast_s2: the counters' lines are pulled into L2 before the barrier
#endif
#ifndef GBN_X1_AHEAD
#define GBN_X1_AHEAD 3        // k_fast_x1: trips of look-ahead of the child-chunk loads (2 or 3)
#endif
#ifndef GBN_XEXP
#define GBN_XEXP 0        // experiments (profiles/r2_summary.md)
#endif

namespace gbn {

namespace {

#ifndef GBN_X_NT
#define GBN_X_NT 1024
#endif
constexpr int FX_NT = GBN_X_NT;      // X kernel: one warp per TF of the speculative window
#ifndef GBN_XQ
#define GBN_XQ 2
#endif
#ifndef GBN_XU
#define GBN_XU 4
#endif
constexpr int XU = GBN_XU;           // X kernel: independent gathers per lane and pipeline stage
constexpr int XQ = GBN_XQ;           // X kernel: chains per CTA = chains interleaved in the product cache (power of two)
#ifndef GBN_S_NT
#define GBN_S_NT 256
#endif
constexpr int FS_NT = GBN_S_NT;      // S kernel: one warp = one group of 32 targets
constexpr int FT_NT = 128;           // T kernel

// q with L >= 2^(-q / 32), read off the bits of L (a likelihood: 0.005 <= L < 1, so q <= 245): L = m 2^e with
// 1 <= m < 2, and LOG2_TAB[i] = floor(32 log2(1 + i / 256)) is a lower bound of 32 log2 m for the eight leading
// mantissa bits i.  The bound overshoots -32 log2 L by less than 1.3, i.e. 0.04 bits per child.
__constant__ uint8_t LOG2_TAB[256] = {
    0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 5, 5,
    5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 10, 10,
    10, 10, 10, 10, 10, 11, 11, 11, 11, 11, 11, 11, 12, 12, 12, 12, 12, 12, 12, 12, 13, 13, 13, 13, 13, 13, 13, 14, 14, 14, 14, 14,
    14, 14, 14, 15, 15, 15, 15, 15, 15, 15, 15, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18,
    18, 18, 18, 19, 19, 19, 19, 19, 19, 19, 19, 20, 20, 20, 20, 20, 20, 20, 20, 20, 21, 21, 21, 21, 21, 21, 21, 21, 21, 22, 22, 22,
    22, 22, 22, 22, 22, 22, 23, 23, 23, 23, 23, 23, 23, 23, 23, 24, 24, 24, 24, 24, 24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 25,
    25, 25, 26, 26, 26, 26, 26, 26, 26, 26, 26, 26, 27, 27, 27, 27, 27, 27, 27, 27, 27, 27, 28, 28, 28, 28, 28, 28, 28, 28, 28, 28,
    29, 29, 29, 29, 29, 29, 29, 29, 29, 29, 29, 30, 30, 30, 30, 30, 30, 30, 30, 30, 30, 30, 31, 31, 31, 31, 31, 31, 31, 31, 31, 31 };
__device__ __forceinline__ uint8_t l_bound_bits(double L)
{
    const int hi = __double2hiint(L);
    const int q = 32 * (1023 - ((hi >> 20) & 0x7ff)) - (int) LOG2_TAB[(hi >> 12) & 255];
    return (uint8_t) min(max(q, 0), 255);
}

// One edge factor f = (1 - t x) z of a target is replaced by q f (phi = q - 1): the target's flip weights and
// likelihood after the change.  With rho = 1 + phi w (w = the weight of the edge's code):
//   S = 0 (f in both terms):  L' = L rho,  w0' = q w0 / rho,             w2' = q w2 / rho
//   S = 2 (f in gamma only):  L' = L rho,  w0' = (w0 + phi w2) / rho,    w2' = q w2 / rho
__device__ __forceinline__ void flip_weights(double2 &w, double &L, uint32_t s, double phi)
{
    const double q = 1. + phi;
    const double rho = __fma_rn(phi, s == 0 ? w.x : w.y, 1.);
    const double ir = 1. / rho;
    w.x = (s == 0 ? q * w.x : __fma_rn(phi, w.y, w.x)) * ir;
    w.y = q * w.y * ir;
    L *= rho;
}

// ---------------------------------------------------------------------------------- X phase
// One CTA sweeps the X nodes of XQ chains (a "bundle": chains XQ p .. XQ p + XQ - 1).
//
// Flip weights.  The S phase leaves per (target, chain) two numbers (tc2) and the likelihood itself (tL):
//     w0 = pr0 omz (a + pr2 b) / L,     w2 = pr0 pr2 omz b / L,     L = c0[n] + pr0 omz (a + pr2 b)
// (algebra at k_fast_s).  An edge with S = 0 carries its factor f = (1 - t x) z in both terms of the numerator of
// w0, an edge with S = 2 in that of w2, so replacing f by q f turns L into L (1 + (q - 1) w), w = w0 or w2 by the
// edge's code (an edge that is off: w = 0).  The conditional of X_k (Multinomial.cpp:56-86 with
// XNode::get_children_loglikelihood, XNode.cpp:16-22) needs the product of L over the children for both values
// of X_k; their RATIO is
//     prod_children (1 + phi_k w),      phi_k = -t_k  (X_k = 0 -> 1),   t_k / (1 - t_k)  (X_k = 1 -> 0),
// one 8-byte gather and one fused multiply-add per child - no table lookup, no second product.  Only the
// underflow rule of Multinomial.cpp:78-85 (both outcome likelihoods round to zero: draw from the prior) needs the
// magnitude prod L itself; it is carried along (one more 8-byte gather per child) for exactly the (dataset, TF)
// pairs whose children COULD underflow a double: sum_children log min_h YNOISE[h][y] < -600 (net.tf_uf, computed
// on the host when the evidence is set).
//
// The kernel is bound by L1 wavefronts and instruction issue: a warp-wide gather costs one wavefront per distinct
// 128-byte line, and the children of a TF are scattered over the cache.  The cache is therefore interleaved by
// bundle - tc2[bundle][target][XQ] - and a warp's lanes are (32 / XQ children) x (XQ chains): the XQ lanes of one
// child read one sector.  The chains of a bundle share the window position: a window commits up to the first TF
// that changed in ANY of them (flips are rare, so the shorter commit costs little), and every committed TF still
// saw exactly the state the sequential sweep of its own chain would have seen.
// LPT lanes of a warp work on one TF (32: a warp per TF; 16 / 8: two / four TFs per warp - the per-TF
// instruction overhead is issued by all warps at once, and a longer window means fewer rounds).
// U children per lane and trip, the ids / S codes of the NEXT trip requested before the current trip's weights
// are used; a lane past the end of its TF's list reads the sentinel position E (the dummy target nH: weights 0,
// likelihood 1) instead of branching.
template <int NT, int LPT, int U, bool Q_SMEM>
__global__ void __launch_bounds__(NT, 1) k_fast_x(DevNet net, DevChains ch, SweepArgs a, uint32_t nchains)
{
    constexpr int NW = NT / 32, NS = 32 / LPT, W = NW * NS;                  // warps, TFs per warp, TFs per window
    constexpr uint32_t Q = XQ, CPW = LPT / XQ;                               // chains per CTA, children per TF and gather
    static_assert(LPT >= 2 * XQ, "a TF needs at least two child slots");
    extern __shared__ __align__(16) unsigned char smem_x[];
    __shared__ uint32_t s_dec[W];                                            // per window position: chains of the bundle that flipped
    __shared__ double s_phi[W * XQ];                                         // per window position and chain: phi of the flip
    const uint32_t nX = net.nX, nH = net.nH;
    const uint32_t nXr = (nX + 3) & ~3u, nStfW = n_stf_words(net);
    uint32_t *sPtr = reinterpret_cast<uint32_t *>(smem_x);                   // [nX + 1 rounded up to 4] tf_ptr
    uint8_t *sX = reinterpret_cast<uint8_t *>(sPtr + ((nX + 4) & ~3u));      // [Q][nX rounded up to 4] value | prior class << 1 | may underflow << 2
    uint8_t *sQ = sX + Q * nXr;                                              // [nH + 1][Q] q with L >= 2^-q per (target, chain) (Q_SMEM)

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sg = lane / LPT, sl = lane % LPT;                         // this lane's TF of the warp, lane within it
    const uint32_t h = sl % Q, cs = sl / Q;                                  // this lane's chain and child slot
    const uint32_t pos = warp * NS + sg;                                     // window position of this lane's TF
    const uint32_t cbase = a.chain0 + Q * blockIdx.x;                        // a multiple of Q by construction
    const uint32_t nvalid = min(Q, nchains - Q * blockIdx.x);                // a last, partial bundle recomputes its last chain
    const uint32_t cmine = cbase + min(h, nvalid - 1);
    double2 *tcb = ch.tc2 + (size_t) (cbase / Q) * (nH + 1) * Q;             // bundle base: entry (i, chain) at Q i + chain
    double *tlb = ch.tL + (size_t) (cbase / Q) * (nH + 1) * Q;

    // everything a round needs per TF is staged once: a round must not wait on global memory for
    // anything but the children's ids and weights
    for (uint32_t q = 0; q < Q; q++) {
        const uint32_t c = cbase + min(q, nvalid - 1);
        const uint32_t d = chain_dataset(net, c), gchain = chain_global_id(net, c);
        const uint8_t *gX = ch.X + (size_t) c * nX;
        const uint8_t *uf = net.tf_uf + (size_t) d * nX;
        for (uint32_t k = tid; k < nX; k += NT) sX[q * nX + k] = (uint8_t) (gX[k] | (net.x_prior_active[k] << 1) | (uf[k] << 2));
        if (!a.use_inj) {   // every X uniform of this chain and sweep: one Philox block per four TFs, computed once
            uint4 *gU = reinterpret_cast<uint4 *>(ch.xU + (size_t) c * nXr);
            for (uint32_t blk = tid; blk < nXr / 4; blk += NT) {
                const u32x4 r = philox_rk(blk, a.sweep, gchain, KIND_X, a);
                gU[blk] = make_uint4(r.x, r.y, r.z, r.w);
            }
        }
    }
    for (uint32_t k = tid; k <= nX; k += NT) sPtr[k] = net.fp_ptr[k];
    const uint8_t *gQ = ch.tq + (size_t) (cbase / Q) * (nH + 1) * Q;         // this bundle's bounds, interleaved like the weights
    if (Q_SMEM) {
        static_assert(XQ % 2 == 0, "the bounds are copied two bytes at a time");
        const uint32_t nq = (nH + 1) * Q / 2;
        for (uint32_t i = tid; i < nq; i += NT) reinterpret_cast<uint16_t *>(sQ)[i] = reinterpret_cast<const uint16_t *>(gQ)[i];
    }
    __syncthreads();
    const uint8_t *qh = (Q_SMEM ? sQ : gQ) + h;
    // The TFs that will want the likelihoods of their children (below: their bound came close at the last update -
    // an active hub or two per chain) gather them from an array nothing else reads: pull those entries into L2
    // now, or each trip of such a TF waits for DRAM while the rest of the CTA waits at the round's barrier.
    {
        __shared__ uint32_t s_hot[32], s_nhot;
        if (tid == 0) s_nhot = 0;
        __syncthreads();
        for (uint32_t q = 0; q < nvalid; q++)
            for (uint32_t k = tid; k < nX; k += NT)
                if ((sX[q * nX + k] & 4u) && ch.xqp[(cbase + q) * nX + k] >= 900) {
                    const uint32_t slot = atomicAdd(&s_nhot, 1u);
                    if (slot < 32) s_hot[slot] = k | (q << 31);
                }
        __syncthreads();
        const uint32_t nhot = min(s_nhot, 32u);
        for (uint32_t i = 0; i < nhot; i++) {
            const uint32_t k = s_hot[i] & 0x7fffffffu, q = s_hot[i] >> 31;
            const double *base = tlb + q;
            for (uint32_t c2 = sPtr[k] + tid; c2 < sPtr[k + 1]; c2 += NT)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t) net.fp_child[c2] * Q));
        }
    }

    const double *gT = ch.T + (size_t) cmine * nX;
    const uint32_t *gU = ch.xU + (size_t) cmine * nXr;
    const double *inj = a.use_inj ? ch.inj_flat + ((size_t) cmine * ch.inj_sweeps + a.inj_row) * (nX + net.E) : nullptr;
    // this lane's chain inside the bundle-interleaved cache, as 32-bit element offsets: w0 of target i at
    // tw_all[w_off + 2 Q i], w2 one further; L at ch.tL[l_off + Q i]; the chain's by-TF S words from stf_off
    const double *tw_all = reinterpret_cast<const double *>(ch.tc2);
    const uint32_t l_off = (cbase / Q) * (nH + 1) * Q + h, w_off = 2 * l_off, stf_off = cmine * nStfW;
    const uint32_t E = net.Efp;                     // the sentinel position of the padded by-TF lists

    // Adaptive window.  A round that ends in a flip throws away the evaluations behind the flip.  With the CLI
    // defaults 18 % of the rounds end that way and the full window is right; when X values move often (a diffuse
    // posterior, many active TFs: most rounds end in a flip after ~20 TFs) the window follows the recent commit
    // length: after two flip-ended rounds in a row it is twice the last committed length (at least 8), and a round
    // that commits its whole window doubles it again.  Every thread derives the same length from the same votes;
    // the visiting order is untouched (profiles/r2_summary.md: flip regimes).
    uint32_t wlen = W, streak = 0;
    uint32_t k0 = 0;
    while (k0 < nX) {
        const uint32_t kw = k0 + pos;
        const bool live = kw < nX && pos < wlen;
        int dec = 0;                              // bit 0: drawn value, bit 1: value changed (this lane's chain)
        double phi = 0.;
        if (NS > 1 || live) {                     // warp-uniform when a warp holds one TF; else dead TFs ride along empty
            const uint32_t kr = live ? kw : 0u;
            const uint32_t xs = sX[h * nX + kr], xcur = xs & 1u;
            // per-TF scalars come from global memory (written by this CTA or an earlier kernel): the loads
            // are issued with the first child ids and are off the critical path
            const double t = gT[kr];
            const uint32_t uword = a.use_inj ? 0u : gU[kr];
            uint32_t cbeg = 0, cend = 0;
            if (live) { cbeg = sPtr[kw]; cend = sPtr[kw + 1]; }
            GBN_CHECK(cbeg <= cend && cend <= E);
            uint32_t cc = cbeg + cs;
            uint32_t id[U], sh[U];
            // (32-bit element offsets off the kernel-parameter base pointers: one IMAD.WIDE per address, no
            // 64-bit pointer kept or rebuilt in the loop)
            auto load_ids = [&](uint32_t c0) {
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t cr = c0 + u * CPW;
                    const uint32_t c = cr < cend ? cr : E;
                    id[u] = __ldg(net.fp_child + c);
                    sh[u] = (__ldg(ch.Stfp + (stf_off + (c >> 4))) >> (2 * (c & 15))) & 3u;
                }
            };
            load_ids(cc);
            uint32_t trips = (cend - cbeg + U * CPW - 1) / (U * CPW);        // trip count, uniform over the warp
            if (NS > 1) trips = __reduce_max_sync(0xffffffffu, trips);
            // f -> q f with q = 1 - t (X: 0 -> 1) or 1 / (1 - t) (X: 1 -> 0); phi = q - 1
            phi = -t;
            if (xcur) phi = t / (1. - t);
            // Could prod L underflow a double for this TF?  Never for most (tf_uf = 0, decided on the host from the
            // evidence).  For the rest a cheap bound rides along: q_i with L_i >= 2^(-q_i / 32) (a byte per target and
            // chain, staged in shared memory) summed over the children; only when the sum says "maybe" is the exact
            // product taken (the slow path below).
            const bool track = (xs & 4u) != 0;
            // ... and a TF whose bound came close last time gathers the likelihoods in the same pass (active hubs:
            // about three updates per chain and sweep at C3), so that the exact product rarely needs a second one
#if GBN_XEXP == 2 || GBN_XEXP == 3
            const bool want_l = false;
#else
            const bool want_l = track && ch.xqp[cmine * nX + kr] >= 900;
#endif
            double mr = 1., ml = 1.;                                         // ratio flipped / current; prod L (if gathered)
            int er = 0, el = 0;
            uint32_t qs = 0;
            auto child_loop = [&](auto with_q, auto with_l) {
                constexpr bool WITH_Q = decltype(with_q)::value, WITH_L = decltype(with_l)::value;
                int cnt = 0;
                for (; trips > 0; trips--) {
                    double w[U], lc[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        GBN_CHECK(id[u] <= nH && sh[u] <= 2);
                        const uint32_t iw = sh[u] == 1 ? nH : id[u];         // an edge that is off: the dummy's zero weight
                        w[u] = tw_all[w_off + 2 * Q * iw + (sh[u] >> 1)];
#if GBN_XEXP == 1
                        if (WITH_Q) qs += id[u] & 1u;
#else
                        if (WITH_Q) qs += qh[Q * id[u]];
#endif
                        if (WITH_L) { lc[u] = 1.; if (want_l) lc[u] = ch.tL[l_off + Q * id[u]]; }
                    }
                    cc += U * CPW;
                    load_ids(cc);                                             // next trip
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        mr *= __fma_rn(phi, w[u], 1.);
                        if (WITH_L) ml *= lc[u];
                    }
                    if ((cnt += U) >= 16) { normalise(mr, er); if (WITH_L) normalise(ml, el); cnt = 0; }
                }
            };
            const bool any_l = __any_sync(0xffffffffu, want_l);
            if (any_l) child_loop(std::true_type{}, std::true_type{});
            else if (__any_sync(0xffffffffu, track)) child_loop(std::true_type{}, std::false_type{});
            else child_loop(std::false_type{}, std::false_type{});
            normalise(mr, er);
#pragma unroll
            for (int o = LPT / 2; o >= (int) Q; o >>= 1) {
                mr *= __shfl_xor_sync(0xffffffffu, mr, o);
                er += __shfl_xor_sync(0xffffffffu, er, o);
                qs += __shfl_xor_sync(0xffffffffu, qs, o);
            }
            normalise(mr, er);
            if (any_l) {                                                     // warp-uniform
                normalise(ml, el);
#pragma unroll
                for (int o = LPT / 2; o >= (int) Q; o >>= 1) {
                    ml *= __shfl_xor_sync(0xffffffffu, ml, o);
                    el += __shfl_xor_sync(0xffffffffu, el, o);
                }
                normalise(ml, el);
            }
            if (track && cs == 0 && live && h < nvalid) ch.xqp[cmine * nX + kr] = (uint16_t) min(qs >> 5, 65535u);
#if GBN_XEXP == 3
            const bool exact = false;
#else
            const bool exact = track && qs >= 32 * 990;                      // prior (>= 2^-7) x prod L could leave the normal range
#endif
            if (!exact) { ml = 1.; el = 0; }                                 // far from underflow: only the ratio matters
            const bool second = exact && !want_l;
            const uint32_t xmask = __ballot_sync(0xffffffffu, second);       // the lanes of one (TF, chain) agree
            if (ch.dbg && cs == 0 && live && track) atomicAdd(ch.dbg + 9, 1ull);
            if (ch.dbg && cs == 0 && exact) atomicAdd(ch.dbg + 8, 1ull);
            if (ch.dbg && cs == 0 && want_l) atomicAdd(ch.dbg + 11, 1ull);
            if (second) {
                // not predicted (the bound jumped since the TF's last update): a second pass over all the children
                if (ch.dbg && cs == 0) atomicAdd(ch.dbg + 10, 1ull);
                uint32_t hm = 0;
#pragma unroll
                for (uint32_t b = 0; b < 32; b += Q) hm |= 1u << b;
                const uint32_t tmask = xmask & (hm << h) & (LPT == 32 ? 0xffffffffu : (((1u << LPT) - 1u) << (sg * LPT)));
                ml = 1.; el = 0;
                for (uint32_t c2 = cbeg + cs; c2 < cend; c2 += 4 * CPW) {        // four independent gathers in flight
                    double l4[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t cr = c2 + u * CPW;
                        l4[u] = ch.tL[l_off + Q * __ldg(net.fp_child + (cr < cend ? cr : E))];   // past the end: the dummy's 1
                    }
                    ml *= (l4[0] * l4[1]) * (l4[2] * l4[3]);
                    normalise(ml, el);
                }
#pragma unroll
                for (int o = LPT / 2; o >= (int) Q; o >>= 1) {
                    ml *= __shfl_xor_sync(tmask, ml, o);
                    el += __shfl_xor_sync(tmask, el, o);
                }
                normalise(ml, el);
            }
#if GBN_XEXP == 4
            ml = 1.; el = 0;
#endif
            // outcome likelihoods exp(log prior + sum log L) of Multinomial.cpp:66-75: current = p prod L, flipped = p' prod L ratio
            const double *prob = net.xprob[(xs >> 1) & 1u];
            double mf = ml * mr;
            int ef = el + er;
            normalise(mf, ef);
            double lik_cur, lik_flip;
            common_scale(prob[xcur] * ml, el, prob[1 - xcur] * mf, ef, lik_cur, lik_flip);
            const double l0 = xcur ? lik_flip : lik_cur, l1 = xcur ? lik_cur : lik_flip;
            double cum0 = 0. + l0, cum1 = cum0 + l1;
            if (!(cum1 > 0.)) { cum0 = 0. + prob[0]; cum1 = cum0 + prob[1]; }     // Multinomial.cpp:78-85
            const double u = a.use_inj ? inj[kr] : unit_from_word(uword);
            const double dice = 0. * (1 - u) + cum1 * u;
            const uint32_t nx = dice < cum0 ? 0u : (dice < cum1 ? 1u : xcur);
            dec = (live && h < nvalid) ? ((int) nx | ((nx != xcur) ? 2 : 0)) : 0;
        }
        // bit q of a TF's word: chain q of the bundle flipped
        const uint32_t fb = __ballot_sync(0xffffffffu, (dec & 2) && cs == 0);
        const uint32_t flipped = (fb >> (sg * LPT)) & ((1u << Q) - 1u);
        if (sl == 0) s_dec[pos] = flipped;
        if (cs == 0) s_phi[pos * Q + h] = phi;
        __syncthreads();
        uint32_t first = W;
#pragma unroll
        for (uint32_t w0 = 0; w0 < (uint32_t) W; w0 += 32) {
            const uint32_t f = __ballot_sync(0xffffffffu, w0 + lane < (uint32_t) W && s_dec[(w0 + lane) % W] != 0);
            if (f && first == (uint32_t) W) first = w0 + (uint32_t) __ffs(f) - 1;
        }
        const uint32_t ncommit = first < W ? first + 1 : wlen;
        if (pos < ncommit && live && cs == 0) sX[h * nX + kw] = (uint8_t) ((sX[h * nX + kw] & 6u) | (dec & 1));
        if (a.x_adapt) {
            if (first < (uint32_t) W) { if (++streak >= 2) wlen = min((uint32_t) W, max(8u, 2 * ncommit)); }
            else { streak = 0; wlen = min((uint32_t) W, 2 * wlen); }
        }
        if (first < (uint32_t) W) {
            // The window ends at the one TF that flipped (in one or more chains of the bundle): apply the flip
            // to the weights of each child that depends on it.  The whole CTA does this - every other
            // warp would only wait at the barrier, and a hub's children walked by one warp cost tens of
            // microseconds of dependent loads.
            const uint32_t kf = k0 + first, fb0 = sPtr[kf], fe0 = sPtr[kf + 1], fmask = s_dec[first];
            for (uint32_t q = 0; q < Q; q++) {
                if (!((fmask >> q) & 1u)) continue;
                const double ph = s_phi[first * Q + q];
                const uint32_t *gS = ch.Stfp + (size_t) (cbase + q) * nStfW;
                uint32_t *gA = ch.Aw + (size_t) (cbase + q) * net.nSw;
                for (uint32_t c2 = fb0 + tid; c2 < fe0; c2 += NT) {
                    // the S kernel's "parent is active" bit of this edge follows the flip (whatever the edge's code)
                    const uint32_t ab = net.fp_abit[c2];
                    if (ab == UINT32_MAX) continue;                           // a pad of the by-TF list
                    GBN_CHECK((ab >> 4) < net.nSw);
                    atomicXor(gA + (ab >> 4), 1u << (ab & 15));
                    const uint32_t s = code_at(gS, c2);
                    if (s == 1) continue;
                    const size_t i = (size_t) net.fp_child[c2] * Q + q;
                    flip_weights(tcb[i], tlb[i], s, ph);
                    if (Q_SMEM) sQ[i] = l_bound_bits(tlb[i]);                 // later windows of this sweep read the bound here
                    else ch.tq[(size_t) (cbase / Q) * (nH + 1) * Q + i] = l_bound_bits(tlb[i]);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && ch.dbg) {
            atomicAdd(ch.dbg + 0, 1ull);
            atomicAdd(ch.dbg + 1, (unsigned long long) (k0 + ncommit <= nX ? ncommit : nX - k0));
            atomicAdd(ch.dbg + 2, first < W ? 1ull : 0ull);
        }
        k0 += ncommit;
    }
    // every TF was committed exactly once: write X back and count (Multinomial.cpp:101-105 as a count)
    for (uint32_t q = 0; q < nvalid; q++) {
        uint8_t *gX = ch.X + (size_t) (cbase + q) * nX;
        uint32_t *cX1 = ch.cX1 + (size_t) (cbase + q) * nX;
        for (uint32_t k = tid; k < nX; k += NT) { const uint32_t x = sX[q * nX + k] & 1u; gX[k] = (uint8_t) x; cX1[k] += x; }
    }
}

// ---------------------------------------------------------------------------------- X phase, one chain per CTA
// k_fast_x (above) gathers 8 bytes per (child, chain) from the flip-weight cache in L2: one L1 wavefront per
// distinct line, 16 lines per warp-wide gather - the kernel is bound by those wavefronts.  Here the flip weights of
// ONE chain, rounded to fp32 by the S kernel (2 x 4 bytes per target: 160 KB at 20,000 targets), and the bound
// bytes are staged in shared memory by bulk asynchronous copies (cp.async.bulk -> mbarrier), and the children's
// gathers become shared-memory loads (a warp-wide gather of 32 random words costs ~3.5 bank-conflict cycles).
//
// The draw must not depend on the rounding.  The fp32 product R~ of the factors 1 + phi w carries a RIGOROUS
// relative error bound B (below); the draw is decided from R~ only if R~ (1 - 2B) and R~ (1 + 2B) give the same
// value - i.e. the uniform is not within the error of the boundary (probability ~ B p0 p1, around 1e-4 per
// update).  Otherwise, and for the TFs whose children's product may underflow a double (the underflow rule of
// Multinomial.cpp:78-85 needs the exact magnitude), the WHOLE CTA evaluates the TF in fp64 from the global cache:
// the arithmetic of k_fast_x, one trip even for a hub.  Only undecided TFs in front of the window's first flip
// are resolved; the others are evaluated again by a later window anyway.
//
// Error bound.  Per factor f = 1 + phi w computed as fmaf(fl(phi), fl(w), 1): |f~ - f| <= |phi w| 2^-23 + |f| 2^-24
// (+ second-order terms), and |phi w| = |f - 1| <= f + 1, so the relative error is <= 2^-23 (1.5 + 1 / f).  The
// running fp32 product of U factors and its conversion add 2^-24 per factor; the fp64 accumulation is exact at
// this scale.  With n factors and m = the smallest computed factor:  B = 1.05 n 2^-23 (2 + 1 / m).  Lanes past the
// end of a list multiply by exactly 1 (the dummy's weight is 0) and are counted in n anyway.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// the draw of Multinomial.cpp:66-99 for a binary node from the likelihood ratio flipped / current = mr 2^er and, where
// the magnitude matters, the current product ml 2^el (1, 0 otherwise); returns the drawn value
__device__ __forceinline__ uint32_t x_draw(const double *prob, uint32_t xcur, double ml, int el, double mr, int er, double u)
{
    double mf = ml * mr;
    int ef = el + er;
    normalise(mf, ef);
    double lik_cur, lik_flip;
    common_scale(prob[xcur] * ml, el, prob[1 - xcur] * mf, ef, lik_cur, lik_flip);
    const double l0 = xcur ? lik_flip : lik_cur, l1 = xcur ? lik_cur : lik_flip;
    double cum0 = 0. + l0, cum1 = cum0 + l1;
    if (!(cum1 > 0.)) { cum0 = 0. + prob[0]; cum1 = cum0 + prob[1]; }     // Multinomial.cpp:78-85
    const double dice = 0. * (1 - u) + cum1 * u;
    return dice < cum0 ? 0u : (dice < cum1 ? 1u : xcur);
}

constexpr uint32_t X1_BULK = 16384;          // bytes per bulk copy

// shared-memory loads at 32-bit addresses (volatile: the flip patch rewrites the entries between rounds)
__device__ __forceinline__ float lds_f32_volatile(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8_volatile(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x)
{
    uint32_t v;
    asm volatile("mov.u32 %0, %1;" : "=r"(v) : "r"(x));
    return v;
}
__device__ __forceinline__ uint32_t smem_addr_opaque(const void *p)
{
    unsigned long long a;
    asm("cvta.to.shared.u64 %0, %1;" : "=l"(a) : "l"(p));
    return (uint32_t) a;
}

template <int NT, int LPT, int U>
__global__ void __launch_bounds__(NT, NT >= 1024 ? 1 : (NT >= 512 ? 2 : 4)) k_fast_x1(DevNet net, DevChains ch, SweepArgs a)
{
    constexpr int NW = NT / 32, NS = 32 / LPT, W = NW * NS;                  // warps, TFs per warp, TFs per window
    constexpr uint32_t Q = XQ;                                               // interleave of the fp64 cache (the fallback reads it)
    static_assert(NW <= 32, "the block-wide reduction takes one partial per lane");
    extern __shared__ __align__(16) unsigned char smem_x[];
    __shared__ uint32_t s_dec[W];                                            // per window position: 0 kept, 1 flipped, 2 undecided
    __shared__ double s_phi[W];                                              // per window position: phi of the flip
    __shared__ double s_rm[NW], s_rl[NW];                                    // block-wide reduction of a resolved TF
    __shared__ int s_re[NW], s_rel[NW];
    __shared__ uint32_t s_rq[NW];
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t nX = net.nX, nH = net.nH, nHp = n_w32(net), nHq = n_q8(net);
    const uint32_t nXr = (nX + 3) & ~3u, nStfW = n_stf_words(net);
    float *sW0 = reinterpret_cast<float *>(smem_x);                          // [nHp] w0 per target (entry nH: 0)
    float *sW2 = sW0 + nHp;                                                  // [nHp] w2
    uint8_t *sQ = reinterpret_cast<uint8_t *>(sW2 + nHp);                    // [nHq] q with L >= 2^(-q / 32)
    uint32_t *sPtr = reinterpret_cast<uint32_t *>(sQ + nHq);                 // [nX + 1 rounded up to 4] tf_ptr
    uint8_t *sX = reinterpret_cast<uint8_t *>(sPtr + ((nX + 4) & ~3u));      // [nX] value | prior class << 1 | may underflow << 2

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sg = lane / LPT, cs = lane % LPT;                         // this lane's TF of the warp, child slot within it
    const uint32_t pos = warp * NS + sg;                                     // window position of this lane's TF
    const uint32_t c = a.chain0 + blockIdx.x;
    const uint32_t d = chain_dataset(net, c), gchain = chain_global_id(net, c);

    // ---- stage the chain: weights and bounds by bulk copies, the rest by the threads while the copies fly
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t bw = 4 * nHp, bq = nHq;
        mbar_expect_tx(&s_bar, 2 * bw + bq);
        const unsigned char *g0 = reinterpret_cast<const unsigned char *>(ch.w0f + (size_t) c * nHp);
        const unsigned char *g2 = reinterpret_cast<const unsigned char *>(ch.w2f + (size_t) c * nHp);
        const unsigned char *gq = ch.tq1 + (size_t) c * nHq;
        for (uint32_t o = 0; o < bw; o += X1_BULK) bulk_g2s(reinterpret_cast<unsigned char *>(sW0) + o, g0 + o, min(X1_BULK, bw - o), &s_bar);
        for (uint32_t o = 0; o < bw; o += X1_BULK) bulk_g2s(reinterpret_cast<unsigned char *>(sW2) + o, g2 + o, min(X1_BULK, bw - o), &s_bar);
        for (uint32_t o = 0; o < bq; o += X1_BULK) bulk_g2s(sQ + o, gq + o, min(X1_BULK, bq - o), &s_bar);
    }
    {
        const uint8_t *gX = ch.X + (size_t) c * nX;
        const uint8_t *uf = net.tf_uf + (size_t) d * nX;
        for (uint32_t k = tid; k < nX; k += NT) sX[k] = (uint8_t) (gX[k] | (net.x_prior_active[k] << 1) | (uf[k] << 2));
        if (!a.use_inj) {   // every X uniform of this chain and sweep: one Philox block per four TFs, computed once
            uint4 *gU = reinterpret_cast<uint4 *>(ch.xU + (size_t) c * nXr);
            for (uint32_t blk = tid; blk < nXr / 4; blk += NT) {
                const u32x4 r = philox_rk(blk, a.sweep, gchain, KIND_X, a);
                gU[blk] = make_uint4(r.x, r.y, r.z, r.w);
            }
        }
        for (uint32_t k = tid; k <= nX; k += NT) sPtr[k] = net.fp_ptr[k];
    }
    __syncthreads();
    // this chain inside the bundle-interleaved fp64 cache, as 32-bit element offsets: w0 of target i at
    // tw_all[w_off + 2 Q i], w2 one further; L at ch.tL[l_off + Q i]
    const double *tw_all = reinterpret_cast<const double *>(ch.tc2);
    const uint32_t l_off = (c / Q) * (nH + 1) * Q + (c % Q), w_off = 2 * l_off, stf_off = c * nStfW;
    const uint32_t E = net.Efp;                     // the sentinel position of the padded by-TF lists
    // The TFs whose magnitude bound reached the exact-product threshold at their last update (an active hub per chain at
    // C3) will most likely be resolved in fp64 again, from cache lines nothing else touches: pull them into L2 now.
    {
        __shared__ uint32_t s_hot[32], s_nhot;
        if (tid == 0) s_nhot = 0;
        __syncthreads();
        for (uint32_t k = tid; k < nX; k += NT)
            if ((sX[k] & 4u) && ch.xqp[c * nX + k] >= 985) {
                const uint32_t slot = atomicAdd(&s_nhot, 1u);
                if (slot < 32) s_hot[slot] = k;
            }
        __syncthreads();
        const uint32_t nhot = min(s_nhot, 32u);
        for (uint32_t i = 0; i < nhot; i++) {
            const uint32_t k = s_hot[i];
            for (uint32_t c2 = sPtr[k] + tid; c2 < sPtr[k + 1]; c2 += NT) {
                const uint32_t t = net.fp_child[c2];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ch.tL + (l_off + Q * t)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(tw_all + (w_off + 2 * Q * t)));
            }
        }
    }
    while (!mbar_try_wait(&s_bar, 0)) { }

    const double *gT = ch.T + (size_t) c * nX;
    const uint32_t *gU = ch.xU + (size_t) c * nXr;
    const double *inj = a.use_inj ? ch.inj_flat + ((size_t) c * ch.inj_sweeps + a.inj_row) * (nX + net.E) : nullptr;
    uint16_t *xqp = ch.xqp + (size_t) c * nX;
    const uint4 *child4 = reinterpret_cast<const uint4 *>(net.fp_child);                                  // chunk q: children 4q .. 4q+3
    const uint8_t *code_base = reinterpret_cast<const uint8_t *>(ch.Stfp);                                // byte code_off + q: their S codes
    const uint32_t code_off = opaque_u32(4 * stf_off);                       // (kept in a register: left to the compiler, the offset is
                                                                             // rebuilt from the CTA index and the network size in every trip)
    const uint32_t sW0_s = smem_addr_opaque(sW0), sW2_s = smem_addr_opaque(sW2), sQ_s = smem_addr_opaque(sQ);

    uint32_t wlen = W, streak = 0;                                           // adaptive window: see k_fast_x
    uint32_t k0 = 0;
    while (k0 < nX) {
        const uint32_t kw = k0 + pos;
        const bool live = kw < nX && pos < wlen;
        uint32_t state = 0;                        // this lane's TF: 0 kept, 1 flipped, 2 undecided
        uint32_t nxv = 0;
        double phi = 0.;
        if (NS > 1 || live) {                     // warp-uniform when a warp holds one TF; else dead TFs ride along empty
            const uint32_t kr = live ? kw : 0u;
            const uint32_t xs = sX[kr], xcur = xs & 1u;
            const double t = gT[kr];
            const uint32_t uword = a.use_inj ? 0u : gU[kr];
            const bool track = live && (xs & 4u) != 0;
            // The by-TF lists are padded to multiples of four per TF (topology.hpp): a lane takes CHUNKS of four consecutive
            // children - one 16-byte load of the ids, one byte load of their four S codes - LPT chunks apart, and requests
            // the chunks of the next three trips before it touches this trip's weights (a trip's shared-memory gathers are
            // short; with one trip of look-ahead every trip waited for an L2 round trip).
            uint32_t q4 = 0, q4e = 0;                                        // this lane's chunk; the end of the TF's chunks
            if (live) { q4 = sPtr[kw] >> 2; q4e = sPtr[kw + 1] >> 2; }
            GBN_CHECK(q4 <= q4e && 4 * q4e <= E);
            uint32_t trips = (q4e - q4 + LPT - 1) / LPT;                     // trip count, uniform over the warp
            if (NS > 1) trips = __reduce_max_sync(0xffffffffu, trips);
            const uint32_t nfac = trips * 4 * LPT;                           // factors of this TF's product, idle lanes included
            q4 += cs;
            struct Chunk { uint4 id; uint32_t cd; };
            auto load_chunk = [&](uint32_t q) {
                const uint32_t qq = q < q4e ? q : (E >> 2);                  // past the end: the sentinel chunk (dummy target, code 1)
                Chunk r;
                r.id = __ldg(child4 + qq);
                r.cd = __ldg(code_base + (code_off + qq));                   // a 32-bit offset off the parameter's pointer
                return r;
            };
            Chunk cb[GBN_X1_AHEAD + 1];                                      // the chunk in use and those requested ahead
#pragma unroll
            for (int i = 0; i < GBN_X1_AHEAD; i++) cb[i] = load_chunk(q4 + i * LPT);
            phi = -t;                                                        // f -> q f, q = 1 - t (X: 0 -> 1) or 1 / (1 - t) (1 -> 0)
            if (xcur) phi = t / (1. - t);
            const float phf = (float) phi;
            double mr = 1.;
            int er = 0;
            float minf = 1.f;
            uint32_t qs = 0;
            auto child_loop = [&](auto with_q) {
                constexpr bool WITH_Q = decltype(with_q)::value;
                int cnt = 0;
                // one trip: the chunk three trips ahead is requested into the buffer the previous trip used, then this
                // trip's four children are gathered from shared memory (32-bit addresses: behind generic pointers the
                // shared window base is rebuilt in every trip)
                auto trip = [&](const Chunk &cur, Chunk &ahead) {
                    ahead = load_chunk(q4 + GBN_X1_AHEAD * LPT);
                    const uint32_t id[4] = { cur.id.x, cur.id.y, cur.id.z, cur.id.w };
                    float w[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t sh = (cur.cd >> (2 * u)) & 3u;
                        GBN_CHECK(id[u] <= nH && sh <= 2);
                        const uint32_t base = (sh & 2u) ? sW2_s : sW0_s;
                        w[u] = lds_f32_volatile(base + 4 * (sh == 1 ? nH : id[u]));   // an edge that is off: the dummy's zero weight
                        if (WITH_Q) qs += lds_u8_volatile(sQ_s + id[u]);
                    }
                    float pm = 1.f;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const float f = __fmaf_rn(phf, w[u], 1.f);
                        minf = fminf(minf, f);
                        pm *= f;
                    }
                    mr *= (double) pm;
                    if ((cnt += 4) >= 16) { normalise(mr, er); cnt = 0; }
                    q4 += LPT;
                    return --trips == 0;
                };
                // the chunk buffers take turns (no register moves between trips)
                if (trips == 0) return;
                for (;;) {
#pragma unroll
                    for (int i = 0; i <= GBN_X1_AHEAD; i++)
                        if (trip(cb[i], cb[(i + GBN_X1_AHEAD) % (GBN_X1_AHEAD + 1)])) return;
                }
            };
            if (__any_sync(0xffffffffu, track)) child_loop(std::true_type{}); else child_loop(std::false_type{});
            normalise(mr, er);
#pragma unroll
            for (int o = LPT / 2; o >= 1; o >>= 1) {
                mr *= __shfl_xor_sync(0xffffffffu, mr, o);
                er += __shfl_xor_sync(0xffffffffu, er, o);
                qs += __shfl_xor_sync(0xffffffffu, qs, o);
                minf = fminf(minf, __shfl_xor_sync(0xffffffffu, minf, o));
            }
            normalise(mr, er);
            if (live) {
                const bool exact = track && qs >= 32 * 990;                  // prior (>= 2^-7) x prod L could leave the normal range
                if (track && cs == 0) xqp[kr] = (uint16_t) min(qs >> 5, 65535u);
                // (a NaN or non-positive factor fails the first test; phi beyond 2^20 only when t is within 1e-6 of 1)
                // (a NaN or non-positive factor fails the first test; phi beyond 2^20 only when t is within 1e-6 of 1)
                const bool r_ok = minf >= 0x1p-10f && fabsf(phf) <= 0x1p20f;
                // twice the bound (rounded up: the reciprocal is an approximation within 2 ulp)
                const double B2 = (double) (2.2f * 0x1p-23f * (float) nfac * (2.f + __frcp_ru(fmaxf(minf, 0x1p-10f))));
                bool sure = !exact && r_ok && B2 < 0.125;
                const double u = a.use_inj ? inj[kr] : unit_from_word(uword);
                const double *prob = net.xprob[(xs >> 1) & 1u];
                if (sure) {
                    // the draw of x_draw() from R~, and its distance from the boundary: a relative change delta of the
                    // flipped outcome's likelihood moves dice - cum0 by delta l_flip u (flipped outcome = 1) or
                    // delta l_flip (u - 1) (flipped outcome = 0)
                    double lik_cur, lik_flip;
                    common_scale(prob[xcur], 0, prob[1 - xcur] * mr, er, lik_cur, lik_flip);
                    const double l0 = xcur ? lik_flip : lik_cur, l1 = xcur ? lik_cur : lik_flip;
                    const double cum0 = 0. + l0, cum1 = cum0 + l1;
                    const double dice = 0. * (1 - u) + cum1 * u;
                    nxv = dice < cum0 ? 0u : 1u;
                    sure = fabs(dice - cum0) > B2 * lik_flip * (xcur ? 1. - u : u) && dice < cum1;
                } else if (exact && r_ok && B2 < 0.125) {
                    // Far below the double range: the bytes also bound the product from above, L < 2^(-(q - 1.3) / 32), so
                    // log2 prod L < ub, and log2 of the flipped outcome's product < ub + er + 1.  When both are below -1090
                    // both outcome likelihoods are exactly zero in fp64 (as the reference's exp() delivers them), and the
                    // draw is from the prior (Multinomial.cpp:78-85) - no exact product needed.
                    const float ub = (1.3f * (float) nfac - (float) qs) * (1.f / 32.f);
                    if (ub < -1090.f && ub + (float) er + 1.f < -1090.f) {
                        const double cum0 = 0. + prob[0], cum1 = cum0 + prob[1];
                        const double dice = 0. * (1 - u) + cum1 * u;
                        nxv = dice < cum0 ? 0u : (dice < cum1 ? 1u : xcur);
                        sure = true;
                        if (ch.dbg && cs == 0) atomicAdd(ch.dbg + 11, 1ull);
                    }
                }
                if (a.x1_force && kw % a.x1_force == 0) sure = false;
                state = sure ? (nxv != xcur ? 1u : 0u) : 2u;
                if (ch.dbg && cs == 0 && track) atomicAdd(ch.dbg + 9, 1ull);
            }
        }
        if (cs == 0) { s_dec[pos] = state; s_phi[pos] = phi; }
        __syncthreads();
        // Undecided TFs in front of the window's first decided flip are resolved in fp64 from the global cache (the
        // arithmetic of k_fast_x), all at once: G = NW / (their number) warps per TF - the whole CTA for the usual single
        // hub, one warp each when there are many.  Then the first TF of the window that flipped is known.
        uint32_t first;
        {
            constexpr int NB = (W + 31) / 32;
            uint32_t um[NB];                                                 // undecided positions
            uint32_t fs = W, n_u = 0;                                        // first decided flip
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t v = 32 * b + lane < (uint32_t) W ? s_dec[(32 * b + lane) % W] : 0u;
                const uint32_t f1 = __ballot_sync(0xffffffffu, v == 1u);
                um[b] = __ballot_sync(0xffffffffu, v == 2u);
                if (f1 && fs == (uint32_t) W) fs = 32 * b + (uint32_t) __ffs(f1) - 1;
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
                if ((uint32_t) (32 * b) >= fs) um[b] = 0;
                else if (fs < (uint32_t) (32 * b + 32)) um[b] &= (1u << (fs - 32 * b)) - 1u;
                n_u += __popc(um[b]);
            }
            if (n_u) {                                                       // uniform over the CTA
                uint32_t G = NW;
                while (G > 1 && NW / G < n_u) G >>= 1;
                const uint32_t P = NW / G, wg = warp / G, wl = warp % G;
                for (uint32_t j0 = 0; j0 < n_u; j0 += P) {
                    const uint32_t j = j0 + wg;
                    const bool act = j < n_u;                                // warp-uniform
                    uint32_t up = 0;                                         // the j-th undecided position
                    {
                        uint32_t r = j;
                        bool found = !act;
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            const uint32_t cb = __popc(um[b]);
                            if (!found && r < cb) { up = 32 * b + __fns(um[b], 0, r + 1); found = true; }
                            if (!found) r -= cb;
                        }
                    }
                    const uint32_t kf = k0 + up, xs = sX[kf], xcur = xs & 1u;
                    const bool track = (xs & 4u) != 0;
                    const double ph = s_phi[up];
                    double mr = 1., ml = 1.;
                    int er = 0, el = 0;
                    uint32_t qs = 0;
                    if (act) {
                        const uint32_t fb0 = sPtr[kf], fe0 = sPtr[kf + 1];
                        int cnt = 0;
                        for (uint32_t c2 = fb0 + 32 * wl + lane; c2 < fe0; c2 += 32 * G) {
                            const uint32_t i = net.fp_child[c2];
                            const uint32_t sc = (ch.Stfp[stf_off + (c2 >> 4)] >> (2 * (c2 & 15))) & 3u;
                            GBN_CHECK(i <= nH && sc <= 2 && (i < nH || sc == 1));
                            const double w = tw_all[w_off + 2 * Q * (sc == 1 ? nH : i) + (sc >> 1)];
                            mr *= __fma_rn(ph, w, 1.);
                            if (track) { ml *= ch.tL[l_off + Q * i]; qs += sQ[i]; }
                            if (++cnt == 8) { normalise(mr, er); normalise(ml, el); cnt = 0; }
                        }
                        normalise(mr, er); normalise(ml, el);
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) {
                            mr *= __shfl_xor_sync(0xffffffffu, mr, o); er += __shfl_xor_sync(0xffffffffu, er, o);
                            ml *= __shfl_xor_sync(0xffffffffu, ml, o); el += __shfl_xor_sync(0xffffffffu, el, o);
                            qs += __shfl_xor_sync(0xffffffffu, qs, o);
                            normalise(mr, er); normalise(ml, el);
                        }
                    }
                    if (lane == 0) { s_rm[warp] = mr; s_re[warp] = er; s_rl[warp] = ml; s_rel[warp] = el; s_rq[warp] = qs; }
                    __syncthreads();
                    if (act) {
                        // every warp of the group folds the group's G partials the same way
                        const bool in = lane < G;
                        const uint32_t src = wg * G + (in ? lane : 0u);
                        mr = in ? s_rm[src] : 1.; er = in ? s_re[src] : 0;
                        ml = in ? s_rl[src] : 1.; el = in ? s_rel[src] : 0;
                        qs = in ? s_rq[src] : 0u;
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) {
                            mr *= __shfl_xor_sync(0xffffffffu, mr, o); er += __shfl_xor_sync(0xffffffffu, er, o);
                            ml *= __shfl_xor_sync(0xffffffffu, ml, o); el += __shfl_xor_sync(0xffffffffu, el, o);
                            qs += __shfl_xor_sync(0xffffffffu, qs, o);
                            normalise(mr, er); normalise(ml, el);
                        }
                        const bool exact = track && qs >= 32 * 990;
                        if (!exact) { ml = 1.; el = 0; }                     // far from underflow: only the ratio matters
                        const double u = a.use_inj ? inj[kf] : unit_from_word(gU[kf]);
                        const uint32_t nx = x_draw(net.xprob[(xs >> 1) & 1u], xcur, ml, el, mr, er, u);
                        if (wl == 0 && lane == 0) {
                            s_dec[up] = nx != xcur ? 1u : 0u;
                            if (ch.dbg) { atomicAdd(ch.dbg + 10, 1ull); if (exact) atomicAdd(ch.dbg + 8, 1ull); }
                        }
                    }
                    __syncthreads();
                }
            }
            first = W;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t f1 = __ballot_sync(0xffffffffu, 32 * b + lane < (uint32_t) W && s_dec[(32 * b + lane) % W] == 1u);
                if (f1 && first == (uint32_t) W) first = 32 * b + (uint32_t) __ffs(f1) - 1;
            }
        }
        const uint32_t ncommit = first < W ? first + 1 : wlen;
        if (pos < ncommit && live && cs == 0) sX[kw] = (uint8_t) (sX[kw] ^ (s_dec[pos] & 1u));    // decided: 0 kept, 1 flipped
        if (a.x_adapt) {
            if (first < (uint32_t) W) { if (++streak >= 2) wlen = min((uint32_t) W, max(8u, 2 * ncommit)); }
            else { streak = 0; wlen = min((uint32_t) W, 2 * wlen); }
        }
        if (first < (uint32_t) W) {
            // The window ends at the TF that flipped: apply the flip to the weights of each child that depends on it -
            // the fp64 cache (what a later resolve reads), its fp32 image and the bound byte in shared memory, and the
            // S kernel's "parent is active" bit of the edge.  The whole CTA walks the children.
            const uint32_t kf = k0 + first, fb0 = sPtr[kf], fe0 = sPtr[kf + 1];
            const double ph = s_phi[first];
            const uint32_t *gS = ch.Stfp + (size_t) c * nStfW;
            uint32_t *gA = ch.Aw + (size_t) c * net.nSw;
            double2 *tcb = ch.tc2 + (size_t) (c / Q) * (nH + 1) * Q + (c % Q);
            double *tlb = ch.tL + (size_t) (c / Q) * (nH + 1) * Q + (c % Q);
            for (uint32_t c2 = fb0 + tid; c2 < fe0; c2 += NT) {
                const uint32_t ab = net.fp_abit[c2];
                if (ab == UINT32_MAX) continue;                               // a pad of the by-TF list
                GBN_CHECK((ab >> 4) < net.nSw);
                atomicXor(gA + (ab >> 4), 1u << (ab & 15));
                const uint32_t s = code_at(gS, c2);
                if (s == 1) continue;
                const uint32_t i = net.fp_child[c2];
                double2 w = tcb[(size_t) Q * i];
                double L = tlb[(size_t) Q * i];
                flip_weights(w, L, s, ph);
                tcb[(size_t) Q * i] = w;
                tlb[(size_t) Q * i] = L;
                sW0[i] = (float) w.x;
                sW2[i] = (float) w.y;
                sQ[i] = l_bound_bits(L);
            }
        }
        __syncthreads();
        if (tid == 0 && ch.dbg) {
            atomicAdd(ch.dbg + 0, 1ull);
            atomicAdd(ch.dbg + 1, (unsigned long long) (k0 + ncommit <= nX ? ncommit : nX - k0));
            atomicAdd(ch.dbg + 2, first < W ? 1ull : 0ull);
        }
        k0 += ncommit;
    }
    // every TF was committed exactly once: write X back and count (Multinomial.cpp:101-105 as a count)
    {
        uint8_t *gX = ch.X + (size_t) c * nX;
        uint32_t *cX1 = ch.cX1 + (size_t) c * nX;
        for (uint32_t k = tid; k < nX; k += NT) { const uint32_t x = sX[k] & 1u; gX[k] = (uint8_t) x; cX1[k] += x; }
    }
}

// ---------------------------------------------------------------------------------- T phase
// Beta::sample (Beta.cpp:60-104) for T nodes that do not listen to their children: the blanket is
// the Beta prior alone.  Also publishes FXp = {1 - t*x, 1 / (1 - t*x)} (HNodeORNOR.cpp:59) for
// the S phase.
// Tout: where the new values go (ch.T, or the alternate buffer when the draws run beside the X phase,
// which still reads the old ones); publish: write FXp as well.
// update_t: 0 = publish FXp only (refresh), 1 = update every T, 2 = update the T of inactive TFs only
// (listening T nodes: an inactive TF's children do not depend on its T - the factor is (1 - t*0) z -
// so its blanket is still the prior alone; the active ones are k_fast_tl's)
__global__ void __launch_bounds__(FT_NT) k_fast_t(DevNet net, DevChains ch, SweepArgs a, int update_t, double *Tout, int publish)
{
    const uint32_t nX = net.nX;
    const uint32_t c = blockIdx.y + a.chain0;
    const uint32_t k = blockIdx.x * FT_NT + threadIdx.x;
    if (k >= nX) return;
    const size_t o = (size_t) c * nX + k;
    double t = ch.T[o];
    if (update_t == 2 && ch.X[o] != 0) return;
    if (update_t) {
        const uint32_t gchain = chain_global_id(net, c);
        const double al = net.t_a[k], be = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
        const size_t row = ((size_t) c * ch.inj_sweeps + a.inj_row) * nX + k;
        const double draw = a.use_inj ? ch.inj_beta[row] : draw_beta(net.seed, gchain, a.sweep, k, al, be);
        t = mh_prior_step(t, draw, mean, al, be, lnB, a.use_inj != 0, a.use_inj ? ch.inj_exp[row] : 0.,
                          net.seed, gchain, a.sweep, k);
        Tout[o] = t;
        if (t == 1.) *ch.flags = 1u;
        double m = ch.tMean[o], m2 = ch.tM2[o];
        welford_add(m, m2, t, (double) a.stat_n);
        ch.tMean[o] = m; ch.tM2[o] = m2;
    }
    if (publish) {
        const double fx = 1. - t * (double) ch.X[o];
        ch.FXp[o] = make_double2(fx, 1. / fx);
    }
}

// FXp = {1 - t*x, 1 / (1 - t*x)} from T and the X the X phase just drew (HNodeORNOR.cpp:59): the
// publishing half of k_fast_t when the T draws ran beside the X phase
__global__ void __launch_bounds__(FT_NT) k_fast_fx(DevNet net, DevChains ch, SweepArgs a, const double *T)
{
    const uint32_t k = blockIdx.x * FT_NT + threadIdx.x;
    if (k >= net.nX) return;
    const size_t o = (size_t) (blockIdx.y + a.chain0) * net.nX + k;
    const double fx = 1. - T[o] * (double) ch.X[o];
    ch.FXp[o] = make_double2(fx, 1. / fx);
}

// T nodes that LISTEN to their children (TNode.cpp:15-25; the pyx default), active TFs only: one CTA
// per chain walks the chain's active TFs in index order (the reference's order; T updates of TFs
// that share a target do not commute).  The blanket difference  sum_children log L(t') - log L(t)
// comes from the flip weights (see k_fast_x): the edge's factor (1 - t) z changes by r = (1 - t') / (1 - t),
// so L(t') / L(t) = 1 + (r - 1) w; the product of the ratios is kept as mantissa x 2^e and one log is taken.
constexpr int FTL_NT = 256;
__global__ void __launch_bounds__(FTL_NT) k_fast_tl(DevNet net, DevChains ch, SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_tl[];
    __shared__ uint32_t s_count;
    __shared__ double s_m[FTL_NT / 32];
    __shared__ int s_e[FTL_NT / 32];
    const uint32_t nX = net.nX, nH = net.nH;
    uint32_t *list = reinterpret_cast<uint32_t *>(smem_tl);          // [nX] active TFs, ascending
    const uint32_t c = blockIdx.x + a.chain0;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gchain = chain_global_id(net, c);
    const uint8_t *gX = ch.X + (size_t) c * nX;
    const uint32_t *gStf = ch.Stfp + (size_t) c * n_stf_words(net);
    double2 *tc2 = ch.tc2 + (size_t) (c / XQ) * (nH + 1) * XQ + (c % XQ);    // bundle-interleaved: entry i at XQ i
    double *tL = ch.tL + (size_t) (c / XQ) * (nH + 1) * XQ + (c % XQ);

    if (warp == 0) {
        uint32_t count = 0;
        for (uint32_t base = 0; base < nX; base += 32) {
            const bool act = base + lane < nX && gX[base + lane] != 0;
            const uint32_t mask = __ballot_sync(0xffffffffu, act);
            if (act) list[count + __popc(mask & ((1u << lane) - 1))] = base + lane;
            count += __popc(mask);
        }
        if (lane == 0) s_count = count;
    }
    __syncthreads();
    const uint32_t count = s_count;
    // Everything of an update that does not depend on the children - the proposal, the prior densities,
    // the proposal-density ratio, the exponential draw of the acceptance test - is computed for a batch of
    // active TFs at once, one thread per TF; only the children's likelihood ratio, the decision and the
    // cache update stay in the sequential walk.  (A TF's own T is not touched by the other TFs' updates.)
    __shared__ double s_prop[FTL_NT], s_r[FTL_NT], s_prev_ll[FTL_NT], s_lpp[FTL_NT], s_logq[FTL_NT], s_expo[FTL_NT];
    for (uint32_t b0 = 0; b0 < count; b0 += FTL_NT) {
        const uint32_t nb = min(count - b0, (uint32_t) FTL_NT);
        if (tid < nb) {
            const uint32_t k = list[b0 + tid];
            const size_t o = (size_t) c * nX + k;
            const double al = net.t_a[k], be = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
            const size_t row = ((size_t) c * ch.inj_sweeps + a.inj_row) * nX + k;
            const double prev = ch.T[o];
            const double draw = a.use_inj ? ch.inj_beta[row] : draw_beta(net.seed, gchain, a.sweep, k, al, be);
            const double dx = draw - mean;
            const double prop = prev + dx;
            s_prop[tid] = prop;
            s_r[tid] = (1. - prop) / (1. - prev);
            s_prev_ll[tid] = (0. + beta_logpdf(prev, al, be, lnB)) + 0.;
            s_lpp[tid] = 0. + beta_logpdf(prop, al, be, lnB);   // -inf for an out-of-range proposal
            {                                                   // Beta.cpp:74-78
                const double lq01 = beta_logpdf(mean - dx, al, be, lnB), lq10 = beta_logpdf(mean + dx, al, be, lnB);
                s_logq[tid] = (lq01 > -INFINITY && lq10 > -INFINITY) ? lq01 - lq10 : log(exp(lq01) / exp(lq10));
            }
            s_expo[tid] = a.use_inj ? ch.inj_exp[row] : draw_exp(net.seed, gchain, a.sweep, k);
        }
        __syncthreads();
        for (uint32_t idx = 0; idx < nb; idx++) {
            const uint32_t k = list[b0 + idx];
            const size_t o = (size_t) c * nX + k;
            const double prev = ch.T[o], prop = s_prop[idx], r = s_r[idx];
            const uint32_t cbeg = net.fp_ptr[k], cend = net.fp_ptr[k + 1];
            double dll = 0.;
            if (s_lpp[idx] > -INFINITY) {               // out-of-range proposals are rejected before the children matter
                // sum_children log L(t') - log L(t) = log prod (1 + (r - 1) w), w = the flip weight of the edge's code
                double mr = 1.;
                int er = 0, cnt = 0;
                for (uint32_t cc = cbeg + tid; cc < cend; cc += FTL_NT) {
                    const uint32_t s = code_at(gStf, cc);
                    if (s == 1) continue;
                    const uint32_t i = net.fp_child[cc];
                    GBN_CHECK(i < nH);                      // (pads carry code 1 and were skipped)
                    const double2 w = tc2[XQ * (size_t) i];
                    mr *= __fma_rn(r - 1., s == 0 ? w.x : w.y, 1.);
                    if (++cnt == 8) { normalise(mr, er); cnt = 0; }
                }
                normalise(mr, er);
#pragma unroll
                for (int w = 16; w > 0; w >>= 1) {
                    mr *= __shfl_xor_sync(0xffffffffu, mr, w);
                    er += __shfl_xor_sync(0xffffffffu, er, w);
                }
                normalise(mr, er);
                if (lane == 0) { s_m[warp] = mr; s_e[warp] = er; }
                __syncthreads();
                mr = 1.; er = 0;
                for (int w = 0; w < FTL_NT / 32; w++) { mr *= s_m[w]; er += s_e[w]; }
                dll = log(mr) + (double) er * 0.69314718055994530942;
                __syncthreads();                        // s_m / s_e free for the next TF
            }
            // Beta::metropolis_hastings_with_prior_proposal (Beta.cpp:60-91), as mh_decide with its pieces precomputed
            const double prev_ll = s_prev_ll[idx];
            const double prop_ll = s_lpp[idx] + dll;
            double t = prev;
            if (prop_ll > -INFINITY) {
                t = prop;
                if (prev_ll > -INFINITY) {
                    const double logratio = prop_ll - prev_ll + s_logq[idx];
                    if (!(logratio >= 0.) && !(logratio > -s_expo[idx])) t = prev;
                }
            }
            if (t != prev) {
                // accepted: the edge factors of this TF change by the ratio r inside its children's weights
                for (uint32_t cc = cbeg + tid; cc < cend; cc += FTL_NT) {
                    const uint32_t s = code_at(gStf, cc);
                    if (s == 1) continue;
                    const size_t i = XQ * (size_t) net.fp_child[cc];
                    flip_weights(tc2[i], tL[i], s, r - 1.);
                }
            }
            if (tid == 0) {
                ch.T[o] = t;
                if (t == 1.) *ch.flags = 1u;
                double m = ch.tMean[o], m2 = ch.tM2[o];
                welford_add(m, m2, t, (double) a.stat_n);
                ch.tMean[o] = m; ch.tM2[o] = m2;
                const double fx = 1. - t;
                ch.FXp[o] = make_double2(fx, 1. / fx);
            }
            __syncthreads();                            // cache updates visible to the next active TF
        }
    }
}

// ---------------------------------------------------------------------------------- S phase
// One WARP per (group of 32 device targets, chain); lane = target.  Device targets are sorted by
// in-degree and a group's steps are stored transposed (topology.hpp), so when the lanes visit their
// j-th parent edge together every access below is one run of consecutive addresses.  No shared-memory
// staging of the slots, no block-level barrier after the start, no write-back pass.
//
// Packed views (round 2).  A lane reads the S codes of 16 consecutive steps as ONE
// 32-bit word (2 bits per code) and keeps them in a register for both passes, the prior rows of those
// steps as one more (static) word, the parent TFs of four steps as 4 x 16 bits, and counts outcomes in a
// per-batch pair of 4 x 8-bit words per four steps (dC) that the model layer folds into the 32-bit totals
// before anything reads them.  Per edge and sweep the kernel moves ~5 bytes (round 1: 17) and issues ~1/4
// of the round-1 load / store instructions.
//
// Likelihood algebra.  With N_h = YNOISE[h][y], K = sum_h YPROB[h] N_h, zc = (1-z)^n, omz = 1 - zc
// (HNodeORNOR.cpp:64,82,96 and HNode.cpp:40-46 expanded):
//     L(pr0, pr2, n) = c0[n] + omz[n] * pr0 * (a + pr2 * b),
//     c0[n] = omz[n] N_0 + zc[n] K,   a = N_2 - N_0,   b = N_1 - N_2.
// c0 / omz / zc are per-dataset tables over n (net.stab, built on the host with zc by repeated
// multiplication as in the reference).  The state carried along a target's edges is A = a pr0,
// Bq = b pr0 pr2 and n.  At an edge the factor f = (1 - t x) z is first taken OUT of the state (multiply by
// the reciprocal from FXp), the three candidates are evaluated from the state without the edge, the draw is
// taken, and the factor of the drawn code is put back in - predicated multiplies, no selects, no branch.
// Pad steps carry S = 1 and the prior row {0, 1, 0}: they draw 1 again and count nothing, so the loop does
// not test for them.
// MODE 0: sweep (draw every S of the target, count, refresh the cache)
// MODE 1: refresh only (rebuild the cache from X, T, S; no draws, no counts)
#ifndef GBN_S_MINB
#define GBN_S_MINB 4          // 64 registers: 4 CTAs (32 warps) per SM measured best in round 1 (profiles/r1_summary.md)
#endif
template <int MODE, bool INJ, bool PF>
__global__ void __launch_bounds__(FS_NT, GBN_S_MINB) k_fast_s(DevNet net, DevChains ch, SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t nX = net.nX, nH = net.nH, D = net.stab_D;
    double2 *sTab = reinterpret_cast<double2 *>(smem_raw);                      // [3][D] {c0[y][n], omz[class of y][n]}
    double *sZp = reinterpret_cast<double *>(sTab + 3 * D);                     // [2][D] z^m by noise class

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t c = blockIdx.y + a.chain0;
    const uint32_t d = chain_dataset(net, c);
    const double2 *gFX = ch.FXp + (size_t) c * nX;
    {
        const double *tab = net.stab + (size_t) d * STAB_ROWS * D;
        for (uint32_t j = tid; j < 3 * D; j += FS_NT) {
            const uint32_t yy = j / D, nn = j - yy * D;
            sTab[j] = make_double2(tab[j], tab[(3 + (yy != 1 ? 0 : 1)) * D + nn]);
        }
        for (uint32_t j = tid; j < 2 * D; j += FS_NT) sZp[j] = tab[13 * D + j];
    }
    __syncthreads();

    const uint32_t g = blockIdx.x * (FS_NT / 32) + (tid >> 5);
    if (g >= net.n_groups) return;
    const uint32_t dt = 32 * g + lane;
    const bool valid = dt < nH;
    const uint32_t gb = net.gbase[g];
    const uint32_t W = (net.gbase[g + 1] - gb) >> 5;                 // group width = largest in-degree
    const uint32_t y = valid ? net.y[(size_t) d * nH + dt] : 1u;
    const uint32_t cls = (y != 1) ? 0 : 1;                          // ZY for DE targets, ZN otherwise
    const double z = cls ? net.z_n : net.z_y;
    const double ca = net.ynoise[2][y] - net.ynoise[0][y];          // a
    const double cb = net.ynoise[1][y] - net.ynoise[2][y];          // b
    const double2 *cot = sTab + y * D;                              // {c0, omz} over n for this target's y
    // 32-bit element offsets off the kernel-parameter base pointers (one IMAD.WIDE per address; 64-bit pointers
    // kept in registers get rebuilt from the constant bank in every block under the 64-register cap)
    const uint32_t sb = net.sbase[g] + lane, db = net.dbase[g] + lane;
    const uint32_t sw_off = c * net.nSw + sb, dc_off = c * net.nD + db;
    const uint16_t *par16 = reinterpret_cast<const uint16_t *>(net.par16);   // step j of this lane: [4 (db + 32 (j / 4)) + j % 4]
    const uint32_t W4 = (W + 3) >> 2;                                // 4-step blocks (the last may hold pads)

    if (MODE == 0) {
        // the counters are first touched by the second pass: start pulling their lines into L2 now
        for (uint32_t b = 0; b < W4; b++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ch.dC + (dc_off + 32 * b)));
    }
    // Products over the parents (HNodeORNOR.cpp:55-62, 72-81): A = a pr0, Bq = b pr0 pr2, so that
    // L = c0[n] + omz[n] (A + Bq) and the a, b constants stay out of the edge loop.  An inactive TF contributes
    // the constant factor z (1 - t*0 = 1 exactly), so the products are z^(number of edges with that code) - two
    // population counts per word of 16 steps - times the factors 1 - t of the ACTIVE parents, which the mask
    // words name (about 1 % of the edges); pads hold S = 1.
    double A = ca, Bq = cb;
    uint32_t n = 0;
    {
        uint32_t n0 = 0;
        for (uint32_t w = 0; 16 * w < W; w++) {
            const uint32_t sw = ch.Sw[sw_off + 32 * w];
            GBN_CHECK(net.sbase[g] + lane + 32 * w < net.nSw);
            const uint32_t x = sw ^ 0x55555555u;
            n += __popc((x | (x >> 1)) & 0x55555555u);               // codes != 1
            n0 += __popc(~(sw | (sw >> 1)) & 0x55555555u);           // codes == 0
            for (uint32_t act = ch.Aw[sw_off + 32 * w] & 0xffffu; act; act &= act - 1) {
                const uint32_t u = __ffs(act) - 1, s = (sw >> (2 * u)) & 3u, j = 16 * w + u;
                const uint32_t kk = par16[4 * (db + 32 * (j >> 2)) + (j & 3)];
                GBN_CHECK(kk < nX);
                const double fx = __ldg(&gFX[kk].x);
                if (s == 0) A *= fx;
                if (s != 1) Bq *= fx;
            }
        }
        GBN_CHECK(n < D && n0 <= n);
        A *= sZp[cls * D + n0];
        Bq *= sZp[cls * D + n];
    }
    if (MODE == 0) {
        const double rz = 1. / z;
        const uint32_t gchain = chain_global_id(net, c);
        const uint32_t ref_i = valid ? net.ref_of_dt[dt] : 0u;
        const double *inj_flat = INJ ? ch.inj_flat + ((size_t) c * ch.inj_sweeps + a.inj_row) * (nX + net.E) + nX : nullptr;
        const uint32_t *pEdge = net.par_edge + gb + lane;
        const uint32_t stf_off = c * n_stf_words(net);
        // blocks of four steps; every fourth block starts a new word of S codes / prior rows / activity bits.
        // PF: the block's counters and by-TF positions are requested one block ahead
        uint32_t sw0 = 0, sw = 0, mw = 0, aw = 0, chg = 0;
        uint4 tp4n = make_uint4(0, 0, 0, 0);
        uint2 cwn = make_uint2(0, 0);
        if (PF) { tp4n = net.tfpos4[db]; cwn = __ldcs(ch.dC + dc_off); }
        for (uint32_t bb = 0; bb < W4; bb++) {
            {
                if ((bb & 3u) == 0) { sw0 = ch.Sw[sw_off + 8 * bb]; sw = sw0; mw = net.mor2[sb + 8 * bb]; aw = ch.Aw[sw_off + 8 * bb]; chg = 0; }
                const uint32_t j0 = 4 * bb;
                uint4 tp4;
                uint2 cw;
                if (PF) {
                    tp4 = tp4n; cw = cwn;
                    if (bb + 1 < W4) { tp4n = net.tfpos4[db + 32 * (bb + 1)]; cwn = __ldcs(ch.dC + (dc_off + 32 * (bb + 1))); }
                } else {
                    tp4 = net.tfpos4[db + 32 * bb]; cw = __ldcs(ch.dC + (dc_off + 32 * bb));
                }
                u32x4 rnd = { 0, 0, 0, 0 };
                if (!INJ) rnd = philox_rk(ref_i, a.sweep, gchain, KIND_S | (bb << 8), a);
                uint32_t chg4 = 0;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (j0 + u < W) {                                // warp-uniform
                        const uint32_t s = (sw >> (2 * u)) & 3u, r = (mw >> (2 * u)) & 3u;
                        GBN_CHECK(s <= 2 && n + 1 < D);
                        // f = (1 - t x) z and its reciprocal: z and 1 / z unless the parent TF is active
                        // (about 1 % of the edges.  The diagnostic count keeps the block a branch: without it the
                        // compiler predicates the lookup's 14 instructions into every step of every lane)
                        double f = z, finv = rz;
                        if ((aw >> u) & 1u) {
                            if (ch.dbg) atomicAdd(ch.dbg + 7, 1ull);
                            const uint32_t kk = par16[4 * (db + 32 * bb) + u];
                            GBN_CHECK(kk < nX);
                            const double2 fx = __ldg(gFX + kk);
                            f = fx.x * z; finv = fx.y * rz;
                        }
                        // the state without this edge
                        const double Af = A * finv, Bf = Bq * finv;
                        const double Ar = s == 0 ? Af : A, Br = s != 1 ? Bf : Bq;
                        const uint32_t nR = n - (s != 1);
                        const double2 t0 = cot[nR], t1 = cot[nR + 1];
                        const double *sp = net.sprob[r];             // constant bank, indexed (row 3: a pad step)
                        const double sum = Ar + Br;
                        // the file is compiled without FMA contraction (the reference's x86-64 build has none); these
                        // are explicit: this algebra differs from the reference's by rounding anyway (DESIGN.md
                        // section 4.4), and the edge loop is bound by instruction issue
                        const double L0 = __fma_rn(t1.y, f * sum, t1.x);                   // S = 0: the factor joins pr0
                        const double L1 = __fma_rn(t0.y, sum, t0.x);                       // S = 1: edge off
                        const double L2 = __fma_rn(t1.y, __fma_rn(f, Br, Ar), t1.x);       // S = 2: the factor joins pr2
                        // exp(log prior + log L) of Multinomial.cpp:66-75 as a product; the "0 +" and
                        // "0 * (1 - u)" terms of the reference's running sum / gsl_ran_flat add exact zeros
                        double cum0 = sp[0] * L0;
                        double cum1 = __fma_rn(sp[1], L1, cum0);
                        double cum2 = __fma_rn(sp[2], L2, cum1);
                        if (!(cum2 > 0.)) {                                                // Multinomial.cpp:78-85
                            // (L >= c0[n] > 0: only an all-zero or NaN prior row gets here.  Counted, so that the
                            // block stays a branch instead of six predicated instructions in the loop)
                            if (ch.dbg) atomicAdd(ch.dbg + 3, 1ull);
                            cum0 = sp[0]; cum1 = cum0 + sp[1]; cum2 = cum1 + sp[2];
                        }
                        const uint32_t word = u == 0 ? rnd.x : (u == 1 ? rnd.y : (u == 2 ? rnd.z : rnd.w));
                        double uu = unit_from_word(word);
                        if (INJ) { const uint32_t e = pEdge[32 * (j0 + u)]; uu = e != 0xffffffffu ? inj_flat[e] : 0.5; }
                        const double dice = cum2 * uu;
                        // first v with dice < cum[v] (Multinomial.cpp:92-99); dice < cum[2] always holds (u < 1)
                        // (an injected double can round up to cum[2], and an all-zero prior row leaves cum[2] = 0: no
                        // outcome is below the dice and the value is kept, Multinomial.cpp:92-99)
                        const bool q0 = dice < cum0, q1 = dice < cum1;
                        uint32_t ns = q0 ? 0u : (q1 ? 1u : 2u);
                        if (!(dice < cum2)) ns = s;
                        add_if(cw.x, 1u << (8 * u), ns == 0);                              // Multinomial.cpp:101-105 as counts
                        add_if(cw.y, 1u << (8 * u), ns == 2);
                        const uint32_t x = s ^ ns;
                        if (x) {
                            chg4 |= x << (2 * u);
                            A = ns == 0 ? Ar * f : Ar;
                            Bq = ns != 1 ? Br * f : Br;
                            n = nR + (ns != 1);
                            const uint32_t tp = u == 0 ? tp4.x : (u == 1 ? tp4.y : (u == 2 ? tp4.z : tp4.w));
                            GBN_CHECK(tp < net.Efp);
                            atomicXor(ch.Stfp + (stf_off + (tp >> 4)), x << (2 * (tp & 15)));
                        }
                    }
                }
                __stcs(ch.dC + (dc_off + 32 * bb), cw);              // streamed: the product cache stays in L2
                chg |= chg4 << (8 * (bb & 3u));
                sw >>= 8; mw >>= 8; aw >>= 4;
            }
            if (((bb & 3u) == 3 || bb + 1 == W4) && chg) ch.Sw[sw_off + 8 * (bb & ~3u)] = sw0 ^ chg;
        }
    }
    GBN_CHECK(n < D);
    if (valid) {
        // what the X phase reads (see k_fast_x): L = c0[n] + omz[n] (A + Bq) and the flip weights
        // w0 = omz (A + Bq) / L (an S = 0 edge's factor sits in both terms), w2 = omz Bq / L (an S = 2 edge's)
        const double2 t = cot[n];
        const double g2 = t.y * Bq, g02 = t.y * A + g2;
        const double L = t.x + g02, iL = 1. / L;
        const size_t o = ((size_t) (c / XQ) * (nH + 1) + dt) * XQ + (c % XQ);   // interleaved by chain bundle
        const double w0 = g02 * iL, w2 = g2 * iL;
        ch.tc2[o] = make_double2(w0, w2);
        ch.tL[o] = L;
        if (net.x1) {                                               // the single-chain X kernel's rows (k_fast_x1)
            const size_t of = (size_t) c * n_w32(net) + dt;
            ch.w0f[of] = (float) w0;
            ch.w2f[of] = (float) w2;
            ch.tq1[(size_t) c * n_q8(net) + dt] = l_bound_bits(L);
        } else {
            ch.tq[o] = l_bound_bits(L);
        }
    }
}

// a *= k under a predicate as ONE predicated multiply (see add_if): the state of a target is updated in place
__device__ __forceinline__ void mul_if(double &a, double k, bool p)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mul.rn.f64 %0, %0, %1;\n\t}" : "+d"(a) : "d"(k), "r"((int) p));
}
// {c0, omz} table pair at a 32-bit shared-memory address (+ 16: the next row).  The address is an ordinary register:
// behind a generic pointer the compiler rebuilds the shared window base (S2UR / UMOV / ULEA) in every step
__device__ __forceinline__ double2 lds_pair(uint32_t addr)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_pair_next(uint32_t addr)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_pair_volatile(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64_next(uint32_t addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1+16];" : "=d"(v) : "r"(addr));
    return v;
}
// a ^= k under a predicate as one predicated instruction
__device__ __forceinline__ void xor_if(uint32_t &a, uint32_t k, bool p)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q xor.b32 %0, %0, %1;\n\t}" : "+r"(a) : "r"(k), "r"((int) p));
}
// a copy of a positive double made by the fp64 pipe (one instruction; a register move of a double is two, and the
// S kernel is bound by instruction issue, not by the fp64 pipe)

// S phase, second form (the default; GBN_S_V=1 selects k_fast_s).  Same layout, same draws, same visiting order; what
// changes is the bookkeeping around the seventeen fp64 operations of a step:
//  * the state is updated IN PLACE by predicated multiplies: the edge's factor leaves A / Bq (if its code put it
//    there), the three candidates are evaluated, the drawn code's factor goes back in.  No select pairs on
//    doubles, no second copy of the state, no state update inside a divergent branch.  An unchanged edge costs the
//    state two roundings (A f^-1 f); the cache the X phase reads is rebuilt from the codes every sweep (pass 1),
//    so the drift is bounded by the in-degree (< 2e-14 relative at 72 parents - the size of the rounding
//    differences the fast path already has against the reference's own order of operations);
//  * the new codes of a word of 16 steps are collected unconditionally (one shift-add per step); a changed word
//    is written back and its changed codes are patched into the by-TF copy once per word, with the by-TF position
//    loaded only then (k_fast_s: a branch taken by 47 % of the warp-steps for 2 % of the lanes, and a 16-byte
//    position load per block);
//  * table pairs through a 32-bit shared-memory address that moves with n (no address arithmetic per step);
//  * an outer loop over words and an inner one over blocks: no predicated word loads in every block;
//  * only the table rows the CTA's widest group can reach are staged.
template <bool INJ>
__global__ void __launch_bounds__(FS_NT, GBN_S_MINB) k_fast_s2(DevNet net, DevChains ch, SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t nX = net.nX, nH = net.nH, D = net.stab_D;
    double2 *sTab = reinterpret_cast<double2 *>(smem_raw);                      // [3][D] {c0[y][n], omz[class of y][n]}
    double *sZp = reinterpret_cast<double *>(sTab + 3 * D);                     // [2][D] z^m by noise class
    double *sPr = sZp + 2 * D;                                                  // [4][4] prior rows (row 3: a pad step), 32 bytes each
    double *sZr = sPr + 16;                                                     // [2][2] {z, 1 / z} by noise class

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t c = blockIdx.y + a.chain0;
    const uint2 cmap = net.chain_map[c];                             // {dataset, global chain id}
    const uint32_t d = cmap.x, gchain = cmap.y;
    const double2 *gFX = ch.FXp + (size_t) c * nX;
    const uint32_t g0 = blockIdx.x * (FS_NT / 32), g = g0 + (tid >> 5);
    const bool live = g < net.n_groups;
    // The warp's first requests go out before the tables are staged, so that their latency (the S words and the
    // counters come from DRAM) overlaps the staging and the barrier: the group's record, the lane's evidence and
    // Philox key, the first word of S codes / activity bits / prior rows, the counters' lines (into L2)
    const uint4 gr = live ? net.grec[g] : make_uint4(0u, 0u, 0u, 0u);          // {gbase, width, sbase, dbase}
    const uint32_t gb = gr.x, W = gr.y;                              // group width = largest in-degree
    const uint32_t dt = 32 * g + lane;
    const bool valid = live && dt < nH;
    const uint32_t y = valid ? net.y[(size_t) d * nH + dt] : 1u;
    const uint32_t ref_i = valid ? net.ref_of_dt[dt] : 0u;
    const uint32_t sb = gr.z + lane, db = gr.w + lane;
    // 32-bit element offsets off the kernel-parameter base pointers (one IMAD.WIDE per address)
    const uint32_t sw_off = c * net.nSw + sb, dc_off = c * net.nD + db;
    const uint32_t W4 = (W + 3) >> 2;                                // 4-step blocks (the last may hold pads)
    uint32_t swA = 0x55555555u, awA = 0, mwA = 0xffffffffu;
    if (W) { swA = ch.Sw[sw_off]; awA = ch.Aw[sw_off]; mwA = net.mor2[sb]; }
#if GBN_S2_DC_PREFETCH
    for (uint32_t b = 0; b < W4; b++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ch.dC + (dc_off + 32 * b)));
#endif
    {
        // n never exceeds the in-degree: with groups in descending order of width the CTA's first group bounds it
        const uint32_t rows = net.groups_desc ? min(D, net.grec[g0].y + 2) : D;
        const double *tab = net.stab + (size_t) d * STAB_ROWS * D;
        for (uint32_t nn = tid; nn < rows; nn += FS_NT) {
            const double o0 = tab[3 * D + nn], o1 = tab[4 * D + nn];            // omz of the two noise classes
            sTab[nn] = make_double2(tab[nn], o0);
            sTab[D + nn] = make_double2(tab[D + nn], o1);
            sTab[2 * D + nn] = make_double2(tab[2 * D + nn], o0);
            sZp[nn] = tab[13 * D + nn];
            sZp[D + nn] = tab[14 * D + nn];
        }
        if (tid >= FS_NT - 16) { const uint32_t q = tid - (FS_NT - 16); sPr[q] = (q & 3) < 3 ? net.sprob[q >> 2][q & 3] : 0.; }
        if (tid == FS_NT - 17) { sZr[0] = net.z_y; sZr[1] = net.rz_y; sZr[2] = net.z_n; sZr[3] = net.rz_n; }
    }
    __syncthreads();
    if (!live) return;

    const uint32_t cls = (y != 1) ? 0 : 1;                          // ZY for DE targets, ZN otherwise
    const double z = cls ? net.z_n : net.z_y;
    const double ca = net.ynoise[2][y] - net.ynoise[0][y];          // a
    const double cb = net.ynoise[1][y] - net.ynoise[2][y];          // b
    const uint32_t cot_s = smem_addr_opaque(sTab + y * D);          // {c0, omz} over n for this target's y
    const uint32_t pr_s = smem_addr_opaque(sPr);
    const uint16_t *par16 = reinterpret_cast<const uint16_t *>(net.par16);   // step j of this lane: [4 (db + 32 (j / 4)) + j % 4]
    const uint32_t *tfpos = reinterpret_cast<const uint32_t *>(net.tfpos4);  // likewise

    // pass 1 (as in k_fast_s): A = a pr0, Bq = b pr0 pr2, n from two population counts per word and the factors of
    // the active parents
    double A = ca, Bq = cb;
    uint32_t n = 0;
    {
        uint32_t n0 = 0;
        for (uint32_t w = 0; 16 * w < W; w++) {
            const uint32_t sw = w == 0 ? swA : ch.Sw[sw_off + 32 * w];
            GBN_CHECK(sb + 32 * w < net.nSw);
            const uint32_t x = sw ^ 0x55555555u;
            n += __popc((x | (x >> 1)) & 0x55555555u);               // codes != 1
            n0 += __popc(~(sw | (sw >> 1)) & 0x55555555u);           // codes == 0
            for (uint32_t act = (w == 0 ? awA : ch.Aw[sw_off + 32 * w]) & 0xffffu; act; act &= act - 1) {
                const uint32_t u = __ffs(act) - 1, s = (sw >> (2 * u)) & 3u, j = 16 * w + u;
                const uint32_t kk = par16[4 * (db + 32 * (j >> 2)) + (j & 3)];
                GBN_CHECK(kk < nX);
                const double fx = __ldg(&gFX[kk].x);
                if (s == 0) A *= fx;
                if (s != 1) Bq *= fx;
            }
        }
        GBN_CHECK(n < D && n0 <= n);
        A *= sZp[cls * D + n0];
        Bq *= sZp[cls * D + n];
    }
    const double rz = cls ? net.rz_n : net.rz_y;
    const double *inj_flat = INJ ? ch.inj_flat + ((size_t) c * ch.inj_sweeps + a.inj_row) * (nX + net.E) + nX : nullptr;
    const uint32_t *pEdge = net.par_edge + gb + lane;
    const uint32_t stf_off = c * n_stf_words(net);
    uint32_t na = cot_s + 16 * n;                                    // shared-memory address of cot[n]
    // f = (1 - t x) z and its reciprocal live in registers as z and 1 / z; an active parent (about 1 % of the edges)
    // changes them for its step and puts them back, both in branches that read z and 1 / z from shared memory
    // (volatile loads: hoisted into registers they would be rebuilt from the constant bank in every step)
    const uint32_t zz_s = smem_addr_opaque(sZr + 2 * cls);
    double f = z, finv = rz;
    uint32_t sw0 = swA, mw = mwA, aw = awA;
    // one block of four steps; GUARD: the group's last block, which may end before its fourth step
    auto block = [&](auto guard_tag, const uint32_t w, const uint32_t b, uint32_t &sw, uint32_t &nsw) {
        constexpr bool GUARD = decltype(guard_tag)::value;
        const uint32_t bb = 4 * w + b, j0 = 4 * bb;
        uint2 cw = __ldcs(ch.dC + (dc_off + 32 * bb));
        u32x4 rnd = { 0, 0, 0, 0 };
        if (!INJ) rnd = philox_rk(ref_i, a.sweep, gchain, KIND_S | (bb << 8), a);
        uint32_t ns4 = 0x55u;                                        // four codes "off"; pads stay so
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (!GUARD || j0 + u < W) {                              // warp-uniform
                const bool s_is0 = (sw & (3u << (2 * u))) == 0, s_ne1 = (sw & (1u << (2 * u))) == 0;
                const uint32_t r32 = (u <= 2 ? mw << (5 - 2 * u) : mw >> (2 * u - 5)) & 0x60u;   // 32 x the prior row
                GBN_CHECK(((sw >> (2 * u)) & 3u) <= 2 && na >= cot_s && (na - cot_s) / 16 + 1 < D);
                const bool active = (aw >> u) & 1u;
                if (active) {                                        // (the diagnostic count keeps the block a branch)
                    if (ch.dbg) atomicAdd(ch.dbg + 7, 1ull);
                    const uint32_t kk = par16[4 * (db + 32 * bb) + u];
                    GBN_CHECK(kk < nX);
                    const double2 fx = __ldg(gFX + kk);
                    const double2 zr = lds_pair_volatile(zz_s);
                    f = fx.x * zr.x; finv = fx.y * zr.y;
                }
                // the state without this edge
                mul_if(A, finv, s_is0);
                mul_if(Bq, finv, s_ne1);
                add_if(na, 0xfffffff0u, s_ne1);
                const double2 t0 = lds_pair(na), t1 = lds_pair_next(na);
#if GBN_S2_PRIOR_SMEM
                const double2 sp01 = lds_pair(pr_s + r32);
                const double sp2 = lds_f64_next(pr_s + r32);
#else
                const double *spc = net.sprob[r32 >> 5];             // constant bank, indexed
                const double2 sp01 = make_double2(spc[0], spc[1]);
                const double sp2 = spc[2];
#endif
                const double sum = A + Bq;
                const double L0 = __fma_rn(t1.y, f * sum, t1.x);                       // S = 0: the factor joins pr0
                const double L1 = __fma_rn(t0.y, sum, t0.x);                           // S = 1: edge off
                const double L2 = __fma_rn(t1.y, __fma_rn(f, Bq, A), t1.x);            // S = 2: the factor joins pr2
                // exp(log prior + log L) of Multinomial.cpp:66-75 as a product.  The fallback of Multinomial.cpp:78-85
                // (all three terms zero) cannot happen here: L >= c0[n] >= min_h YNOISE > 0, and the model layer
                // sends prior rows that could make a product vanish to k_fast_s (sprior_safe)
                const double cum0 = sp01.x * L0;
                const double cum1 = __fma_rn(sp01.y, L1, cum0);
                const double cum2 = __fma_rn(sp2, L2, cum1);
                const uint32_t word = u == 0 ? rnd.x : (u == 1 ? rnd.y : (u == 2 ? rnd.z : rnd.w));
                double uu = unit_from_word(word);
                if (INJ) { const uint32_t e = pEdge[32 * (j0 + u)]; uu = e != 0xffffffffu ? inj_flat[e] : 0.5; }
                const double dice = cum2 * uu;
                // first v with dice < cum[v] (Multinomial.cpp:92-99); cum0 <= cum1, so "2" is "not below cum1"
                bool ns_is0 = dice < cum0, ns_is2 = !(dice < cum1);
                if (INJ && !(dice < cum2)) {                         // an injected double can round up to cum[2]: the value is kept
                    ns_is0 = s_is0; ns_is2 = s_ne1 && !s_is0;
                }
                const bool ns_ne1 = ns_is0 || ns_is2;
                add_if(cw.x, 1u << (8 * u), ns_is0);                                   // Multinomial.cpp:101-105 as counts
                add_if(cw.y, 1u << (8 * u), ns_is2);
                mul_if(A, f, ns_is0);
                mul_if(Bq, f, ns_ne1);
                add_if(na, 16u, ns_ne1);
                xor_if(ns4, 1u << (2 * u), ns_ne1);                  // code 1 -> 0 ...
                xor_if(ns4, 2u << (2 * u), ns_is2);                  // ... -> 2
                if (active) {
                    if (ch.dbg) atomicAdd(ch.dbg + 6, 1ull);
                    const double2 zr = lds_pair_volatile(zz_s);
                    f = zr.x; finv = zr.y;
                }
            }
        }
        __stcs(ch.dC + (dc_off + 32 * bb), cw);                      // streamed: the product cache stays in L2
        nsw |= ns4 << (8 * b);
        sw >>= 8; mw >>= 8; aw >>= 4;
    };
    const uint32_t Wf = W >> 2;                                      // blocks with all four steps inside the group's width
    for (uint32_t w = 0;;) {
        uint32_t sw = sw0, nsw = 0;
        const uint32_t nb = min(4u, W4 - 4 * w), nbf = min(nb, Wf - min(Wf, 4 * w));
        uint32_t b = 0;
        for (; b < nbf; b++) block(std::false_type{}, w, b, sw, nsw);
        if (b < nb) block(std::true_type{}, w, b, sw, nsw);
        // blocks past the group's width keep their pads
        uint32_t chg = nb == 4 ? sw0 ^ nsw : (sw0 ^ nsw) & ((1u << (8 * nb)) - 1u);
        if (chg) {
            ch.Sw[sw_off + 32 * w] = sw0 ^ chg;
            const uint32_t tb = 4 * db + 512 * w;                    // by-TF positions of this lane's steps of the word
            // the by-TF copy follows: a trip patches up to N changed codes, their by-TF positions requested together (one
            // code per trip waits for an L2 round trip per code: S prior 0.4/0.4/0.2, where 60 % of the codes move, 0.73 ->
            // 0.51 ms per launch).  Two codes cover the usual word; whatever is left goes four at a time.
            auto patch = [&](auto n_tag) {
                constexpr int N = decltype(n_tag)::value;
                uint32_t tp[N], xx[N];
#pragma unroll
                for (int i = 0; i < N; i++) {
                    xx[i] = 0; tp[i] = 0;
                    if (chg) {
                        const uint32_t u16 = (__ffs(chg) - 1) >> 1;
                        xx[i] = (chg >> (2 * u16)) & 3u;
                        chg &= ~(3u << (2 * u16));
                        tp[i] = tfpos[tb + 128 * (u16 >> 2) + (u16 & 3)];
                        GBN_CHECK(tp[i] < net.Efp);
                    }
                }
#pragma unroll
                for (int i = 0; i < N; i++)
                    if (xx[i]) atomicXor(ch.Stfp + (stf_off + (tp[i] >> 4)), xx[i] << (2 * (tp[i] & 15)));
            };
            patch(std::integral_constant<int, 2>{});
            while (chg) patch(std::integral_constant<int, 4>{});
        }
        w++;
        if (16 * w >= W) break;
        sw0 = ch.Sw[sw_off + 32 * w]; mw = net.mor2[sb + 32 * w]; aw = ch.Aw[sw_off + 32 * w];
    }
    n = (na - cot_s) >> 4;
    GBN_CHECK(n < D);
    if (valid) {
        // what the X phase reads (see k_fast_x): L = c0[n] + omz[n] (A + Bq) and the flip weights
        const double2 t = lds_pair(na);
        const double g2 = t.y * Bq, g02 = t.y * A + g2;
        const double L = t.x + g02, iL = 1. / L;
        const size_t o = ((size_t) (c / XQ) * (nH + 1) + dt) * XQ + (c % XQ);   // interleaved by chain bundle
        const double w0 = g02 * iL, w2 = g2 * iL;
        ch.tc2[o] = make_double2(w0, w2);
        ch.tL[o] = L;
        if (net.x1) {                                               // the single-chain X kernel's rows (k_fast_x1)
            const size_t of = (size_t) c * n_w32(net) + dt;
            ch.w0f[of] = (float) w0;
            ch.w2f[of] = (float) w2;
            ch.tq1[(size_t) c * n_q8(net) + dt] = l_bound_bits(L);
        } else {
            ch.tq[o] = l_bound_bits(L);
        }
    }
}

// ---- conversions between the byte view of S (ch.S: parity hooks, state get / set, checkpoints, strict kernel)
// and the packed views the fast sweep works on
// S bytes -> 2-bit words (after set_state, checkpoint load).  INITIAL (after creation, set_evidence): the S words of
// a freshly initialised chain are the network's initial words; only the activity bits depend on the chain (its X)
template <bool INITIAL>
__global__ void k_pack_s(DevNet net, DevChains ch)
{
    const uint32_t c = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= net.nSw) return;
    const uint32_t row = i >> 5, lane = i & 31, g = net.sw_grp[row];
    const uint32_t jw = row - (net.sbase[g] >> 5), gb = net.gbase[g], W = (net.gbase[g + 1] - gb) >> 5;
    const uint8_t *S = ch.S + (size_t) c * net.Ep + gb + lane;
    const uint8_t *X = ch.X + (size_t) c * net.nX;
    uint32_t w = INITIAL ? net.sw_init[i] : 0u, aw = 0;
    for (uint32_t u = 0; u < 16; u++) {
        const uint32_t j = 16 * jw + u;
        if (!INITIAL) {
            const uint32_t code = j < W ? S[32 * j] : 1u;
            GBN_CHECK(code <= 2);
            w |= code << (2 * u);
        }
        if (j < W && net.par_edge[gb + lane + 32 * j] != 0xffffffffu) aw |= (uint32_t) (X[net.par_tf[gb + lane + 32 * j]] & 1u) << u;
    }
    ch.Sw[(size_t) c * net.nSw + i] = w;
    ch.Aw[(size_t) c * net.nSw + i] = aw;
}

// 2-bit words -> S bytes (before anything reads ch.S after a fast sweep)
__global__ void k_unpack_s(DevNet net, DevChains ch)
{
    const uint32_t c = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= net.nSw) return;
    const uint32_t row = i >> 5, lane = i & 31, g = net.sw_grp[row];
    const uint32_t jw = row - (net.sbase[g] >> 5), gb = net.gbase[g], W = (net.gbase[g + 1] - gb) >> 5;
    uint8_t *S = ch.S + (size_t) c * net.Ep + gb + lane;
    const uint32_t w = ch.Sw[(size_t) c * net.nSw + i];
    for (uint32_t u = 0; u < 16; u++) {
        const uint32_t j = 16 * jw + u;
        if (j < W) S[32 * j] = (uint8_t) ((w >> (2 * u)) & 3u);
    }
}

// per-batch 8-bit counts -> 32-bit totals (Multinomial.cpp:101-105 as counts); dC is left at zero.
// OVERWRITE: the totals were never written (logically zero, capi.cu cs_zero): every slot of the block is set
template <bool OVERWRITE>
__global__ void k_fold_counts(DevNet net, DevChains ch)
{
    const uint32_t c = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= net.nD) return;
    uint2 *pd = ch.dC + (size_t) c * net.nD + i;
    const uint2 dv = *pd;
    if (!OVERWRITE && (dv.x | dv.y) == 0) return;
    const uint32_t row = i >> 5, lane = i & 31, g = net.dblk_grp[row];
    const uint32_t jb = row - (net.dbase[g] >> 5), gb = net.gbase[g], W = (net.gbase[g + 1] - gb) >> 5;
    uint2 *cS = ch.cS + (size_t) c * net.Ep + gb + lane;
#pragma unroll
    for (uint32_t u = 0; u < 4; u++) {
        const uint32_t j = 4 * jb + u;
        const uint32_t d0 = (dv.x >> (8 * u)) & 0xffu, d2 = (dv.y >> (8 * u)) & 0xffu;
        if (j < W && (OVERWRITE || (d0 | d2))) {
            uint2 t = OVERWRITE ? make_uint2(0u, 0u) : cS[32 * j];
            t.x += d0; t.y += d2;
            cS[32 * j] = t;
        }
    }
    if (dv.x | dv.y) *pd = make_uint2(0u, 0u);
}

// Stfp from S (refresh only): one thread per word of the by-TF list.  INITIAL: the network's initial words
template <bool INITIAL>
__global__ void k_fast_stf(DevNet net, DevChains ch)
{
    const uint32_t c = blockIdx.y;
    const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nStfW = n_stf_words(net);
    if (wi >= nStfW) return;
    if (INITIAL) { ch.Stfp[(size_t) c * nStfW + wi] = net.stfp_init[wi]; return; }
    uint32_t w = 0;
    for (uint32_t u = 0; u < 16; u++) {
        const uint32_t pos = 16 * wi + u;
        uint32_t code = 1u;
        if (pos < net.Efp && net.fp_slot[pos] != UINT32_MAX) { GBN_CHECK(net.fp_slot[pos] < net.Ep); code = ch.S[(size_t) c * net.Ep + net.fp_slot[pos]]; }
        w |= code << (2 * u);
    }
    ch.Stfp[(size_t) c * nStfW + wi] = w;
}

}  // namespace

static int check(cudaError_t e, const char **err)
{
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 0;
}

// shared memory of the S kernel: the per-dataset tables over n
static size_t s_smem_bytes(const DevNet &net)
{
    return sizeof(double2) * 3 * (size_t) net.stab_D + sizeof(double) * 2 * (size_t) net.stab_D + 128 + 32 + 16;   // + the prior rows and {z, 1 / z} of k_fast_s2
}

// shared memory of the X kernel (XQ chains per CTA): by-TF pointers, X and - when it fits - the per-target
// magnitude bounds of the bundle
static size_t x_smem_bytes(const DevNet &net, bool q_smem)
{
    const size_t nX = net.nX, nXr = (nX + 3) & ~(size_t) 3, Q = XQ;
    return 4 * ((nX + 4) & ~(size_t) 3) + Q * nXr + (q_smem ? Q * ((size_t) net.nH + 1) + 4 : 0) + 16;
}
static bool x_q_smem(const DevNet &net)
{
    if (env_u32("GBN_X_Q", 1) == 2) return false;        // 2 = read the bounds from global memory (tests of the large-network path)
    return x_smem_bytes(net, true) + 2048 <= 200 * 1024;
}

uint32_t fast_x_bundle() { return XQ; }

// shared memory of the single-chain X kernel: fp32 flip weights and bound bytes of every target, by-TF pointers, X
static size_t x1_smem_bytes(const DevNet &net)
{
    return 8 * (size_t) n_w32(net) + n_q8(net) + 4 * (((size_t) net.nX + 4) & ~(size_t) 3) + net.nX + 16;
}
// the X phase runs on k_fast_x1 when one chain's weights fit an SM's shared memory (GBN_X1=2: never)
bool fast_x1_supported(const DevNet &net)
{
    if (env_u32("GBN_X1", 1) == 2) return false;
    const uint64_t C = (uint64_t) net.n_datasets * net.n_graphs_local + XQ;
    if (C * net.nX >= (1ull << 32) || 4 * C * n_stf_words(net) >= (1ull << 32)) return false;
    return x1_smem_bytes(net) + 4096 <= 227 * 1024;
}

// the T draws of a sweep run beside its X phase (see launch_sweep_fast); the caller swaps ch.T / ch.T2 after
// an odd number of sweeps
bool fast_t_beside_x(const DevNet &net) { return !net.listen && !net.const_params && env_u32("GBN_T_BESIDE_X", 1) != 2; }

// a second stream per chain group pays off even for a single group: T work runs beside the X kernel
// (non-listening T nodes) or the two T kernels run side by side (listening ones)
bool fast_uses_x_stream(const DevNet &net) { return !net.const_params && env_u32("GBN_T_BESIDE_X", 1) != 2; }

bool fast_supported(const DevNet &net, const char **why)
{
    if (net.nX > 65535) { *why = "n_tf does not fit the 16-bit parent id"; return false; }
    if (s_smem_bytes(net) > 200 * 1024) { *why = "in-degree too large for the S-phase tables"; return false; }
    if (net.Eu != net.E) { *why = "duplicate (TF, target) edges"; return false; }
    {   // the X kernel addresses the per-chain caches with 32-bit element offsets
        const uint64_t C = (uint64_t) net.n_datasets * net.n_graphs_local + XQ;
        if (2 * C * (net.nH + 1) >= (1ull << 32) || C * n_stf_words(net) >= (1ull << 32) || C * net.nD >= (1ull << 32)) { *why = "chains x targets exceeds the 32-bit cache offsets"; return false; }
    }
    if (x_smem_bytes(net, false) + 2048 > 200 * 1024) { *why = "n_tf too large for the X window kernel's shared memory"; return false; }
    return true;
}

template <int MODE, bool INJ, bool PF>
static int launch_s_t(const DevNet &net, const DevChains &ch, const SweepArgs &a, uint32_t nchains, cudaStream_t st, const char **err)
{
    const size_t smem = s_smem_bytes(net);
    if (check(cudaFuncSetAttribute(k_fast_s<MODE, INJ, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), err)) return -1;
    const uint32_t gpb = FS_NT / 32;                                  // groups per block
    k_fast_s<MODE, INJ, PF><<<dim3((net.n_groups + gpb - 1) / gpb, nchains), FS_NT, smem, st>>>(net, ch, a);
    return 0;
}

template <bool INJ>
static int launch_s2_t(const DevNet &net, const DevChains &ch, const SweepArgs &a, uint32_t nchains, cudaStream_t st, const char **err)
{
    const size_t smem = s_smem_bytes(net);
    if (check(cudaFuncSetAttribute(k_fast_s2<INJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), err)) return -1;
    {
        static const uint32_t carve = env_u32("GBN_S_CARVE", 0);          // experiment knob, see GBN_X1_CARVE
        if (carve && check(cudaFuncSetAttribute(k_fast_s2<INJ>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) carve), err)) return -1;
    }
    const uint32_t gpb = FS_NT / 32;                                  // groups per block
    k_fast_s2<INJ><<<dim3((net.n_groups + gpb - 1) / gpb, nchains), FS_NT, smem, st>>>(net, ch, a);
    return 0;
}

static int launch_s(const DevNet &net, const DevChains &ch, const SweepArgs &a, int mode, uint32_t nc, cudaStream_t st, const char **err)
{
    static const bool pf = env_u32("GBN_S_PREFETCH", 0) == 1;         // 1 = on (measured: no gain at 64 registers)
    const bool v1 = env_u32("GBN_S_V", 2) == 1;                       // 1 = the first form of the S kernel (k_fast_s)
    if (mode == 1) return launch_s_t<1, false, false>(net, ch, a, nc, st, err);
    if (!v1 && net.sprior_safe) return a.use_inj ? launch_s2_t<true>(net, ch, a, nc, st, err) : launch_s2_t<false>(net, ch, a, nc, st, err);
    if (a.use_inj) return launch_s_t<0, true, false>(net, ch, a, nc, st, err);
    return pf ? launch_s_t<0, false, true>(net, ch, a, nc, st, err) : launch_s_t<0, false, false>(net, ch, a, nc, st, err);
}

int launch_fast_refresh(const DevNet &net, const DevChains &ch, cudaStream_t st, const char **err, bool initial)
{
    if (ch.n_chains == 0) return 0;
    SweepArgs a = sweep_args(net);
    const dim3 gp((net.nSw + 255) / 256, ch.n_chains), gs((n_stf_words(net) + 255) / 256, ch.n_chains);
    if (initial) { k_pack_s<true><<<gp, 256, 0, st>>>(net, ch); k_fast_stf<true><<<gs, 256, 0, st>>>(net, ch); }
    else { k_pack_s<false><<<gp, 256, 0, st>>>(net, ch); k_fast_stf<false><<<gs, 256, 0, st>>>(net, ch); }
    k_fast_t<<<dim3((net.nX + FT_NT - 1) / FT_NT, ch.n_chains), FT_NT, 0, st>>>(net, ch, a, 0, ch.T, 1);
    if (launch_s(net, ch, a, 1, ch.n_chains, st, err)) return -1;
    if (check(cudaGetLastError(), err)) return -1;
    return 4;
}

// ch.S (bytes) from the packed words the fast sweep keeps current
int launch_fast_unpack(const DevNet &net, const DevChains &ch, cudaStream_t st, const char **err)
{
    if (ch.n_chains == 0) return 0;
    k_unpack_s<<<dim3((net.nSw + 255) / 256, ch.n_chains), 256, 0, st>>>(net, ch);
    if (check(cudaGetLastError(), err)) return -1;
    return 1;
}

// per-batch 8-bit outcome counts of the S nodes -> the 32-bit totals in ch.cS
int launch_fast_fold(const DevNet &net, const DevChains &ch, cudaStream_t st, const char **err, bool overwrite)
{
    if (ch.n_chains == 0) return 0;
    const dim3 grid((net.nD + 255) / 256, ch.n_chains);
    if (overwrite) k_fold_counts<true><<<grid, 256, 0, st>>>(net, ch);
    else k_fold_counts<false><<<grid, 256, 0, st>>>(net, ch);
    if (check(cudaGetLastError(), err)) return -1;
    return 1;
}

int launch_sweep_fast(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0,
                      uint32_t stat_n0, cudaStream_t st, const char **err, cudaEvent_t *phase_ev,
                      uint32_t chain0, uint32_t nchains, cudaStream_t stx, cudaEvent_t evx, cudaEvent_t evs)
{
    if (stx == nullptr) stx = st;                      // X kernels on their own (high-priority) stream when given
    if (nchains == UINT32_MAX) nchains = ch.n_chains - chain0;
    if (n_sweeps == 0 || nchains == 0) return 0;
    const bool qs = x_q_smem(net);
    const size_t smem_x = x_smem_bytes(net, qs);
    if (chain0 % XQ) { *err = "chain groups must start at a multiple of the X kernel's bundle size"; return -1; }
    // Lanes per TF by the mean number of targets of a TF: half a warp (a 64-TF window), or a quarter in CTAs of 16
    // warps.  Fewer lanes per TF = less per-TF instruction overhead and, with the longer window, fewer
    // rounds: the fixed latency of a round grows when the S kernels of other chain groups load the memory
    // system.  A whole warp per TF is kept for GBN_X_LPT=32.
    uint32_t lpt = env_u32("GBN_X_LPT", 0);                  // 8 / 16 / 32 force a path (tests)
    const bool big = (uint64_t) net.E >= 48ull * net.nX;
    if (lpt != 8 && lpt != 16 && lpt != 32) lpt = (uint64_t) net.E >= 20ull * net.nX ? 16 : 8;
    void (*kx)(DevNet, DevChains, SweepArgs, uint32_t);
    uint32_t x_nt = FX_NT;
    if (lpt == 32) kx = qs ? k_fast_x<FX_NT, 32, XU, true> : k_fast_x<FX_NT, 32, XU, false>;
    else if (lpt == 16 && big) kx = qs ? k_fast_x<FX_NT, 16, XU, true> : k_fast_x<FX_NT, 16, XU, false>;
    else if (lpt == 16) kx = qs ? k_fast_x<FX_NT, 16, 1, true> : k_fast_x<FX_NT, 16, 1, false>;
    else { kx = qs ? k_fast_x<FX_NT / 2, 8, 1, true> : k_fast_x<FX_NT / 2, 8, 1, false>; x_nt = FX_NT / 2; }
    if (check(cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_x), err)) return -1;
    void (*kx1)(DevNet, DevChains, SweepArgs) = nullptr;
    uint32_t x1_nt = FX_NT;
    const size_t smem_x1 = x1_smem_bytes(net);
    if (net.x1) {
        // one chain per CTA: half as many lanes per TF as the pair kernel for the same children per gather
        uint32_t l1 = env_u32("GBN_X_LPT", 0);
        if (l1 != 4 && l1 != 8 && l1 != 16 && l1 != 32) l1 = big ? 8 : 4;
        // TFs with few targets and more chains than SMs: CTAs of 256 threads (a 64-TF window), four chains per SM -
        // a round is bound by its latency, not by the lanes (GBN_X1_NT=256 / 1024 force one)
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        uint32_t nt1 = env_u32("GBN_X1_NT", 0);
        // TFs with many targets: CTAs of 512 threads - the kernel is bound by the latency of its rounds, not by its
        // lanes (1,024 threads: 0.347 ms per launch at C3, 512: 0.358), and half the registers of the SM stay free
        // for S-kernel CTAs of another chain group (step 23.0 -> 22.5 ms)
        if (nt1 != 256 && nt1 != 512 && nt1 != 1024) nt1 = big ? 512 : (nchains > 2u * (uint32_t) sms ? 256 : 1024);
        if (l1 != 4 && nt1 == 256) nt1 = 1024;
        x1_nt = nt1;
        if (nt1 == 256) kx1 = k_fast_x1<256, 4, XU>;
        else if (nt1 == 512) kx1 = l1 >= 16 ? k_fast_x1<512, 16, XU> : l1 == 8 ? k_fast_x1<512, 8, XU> : k_fast_x1<512, 4, XU>;
        else kx1 = l1 == 32 ? k_fast_x1<FX_NT, 32, XU> : l1 == 16 ? k_fast_x1<FX_NT, 16, XU> : l1 == 8 ? k_fast_x1<FX_NT, 8, XU> : k_fast_x1<FX_NT, 4, XU>;
        if (check(cudaFuncSetAttribute(kx1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_x1), err)) return -1;
        {   // experiment knob: the shared-memory carve-out the kernel asks for, in percent of the maximum (0 = the driver's choice)
            static const uint32_t carve = env_u32("GBN_X1_CARVE", 0);
            if (carve && check(cudaFuncSetAttribute(kx1, cudaFuncAttributePreferredSharedMemoryCarveout, (int) carve), err)) return -1;
        }
    }
    const size_t smem_tl = 4 * (size_t) net.nX + 16;
    if (net.listen && !net.const_params &&
        check(cudaFuncSetAttribute(k_fast_tl, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int) (smem_tl > 48 * 1024 ? smem_tl : 48 * 1024)), err)) return -1;
    int launches = 0;
    // T nodes that do not listen to their children are drawn from the prior alone: their update needs
    // neither this sweep's X nor anything the X phase writes, so it runs beside the X kernel (on `st`,
    // the X kernel on `stx`) into the alternate T buffer - the X phase still reads the previous values -
    // and a small kernel publishes FXp once both are done.  The buffers swap roles every sweep.
    const bool t_beside_x = fast_t_beside_x(net);
    DevChains chl = ch;
    const dim3 grid_t((net.nX + FT_NT - 1) / FT_NT, nchains);
    for (uint32_t sw = 0; sw < n_sweeps; sw++) {
        SweepArgs a = sweep_args(net);
        a.sweep = sweep0 + sw;
        a.stat_n = stat_n0 + sw + 1;
        a.use_inj = (ch.inj_flat != nullptr && (ch.inj_pos + sw) < ch.inj_sweeps) ? 1 : 0;
        a.inj_row = ch.inj_pos + sw;
        a.chain0 = chain0;
        if (phase_ev) cudaEventRecord(phase_ev[3 * sw], st);
        if (stx != st && sw > 0) cudaStreamWaitEvent(stx, evs, 0);      // this chain group's previous S phase
        if (kx1) kx1<<<nchains, x1_nt, smem_x1, stx>>>(net, chl, a);
        else kx<<<(nchains + XQ - 1) / XQ, x_nt, smem_x, stx>>>(net, chl, a, nchains);
        if (stx != st) cudaEventRecord(evx, stx);
        if (t_beside_x) {
            if (phase_ev) cudaEventRecord(phase_ev[3 * sw + 1], st);
            k_fast_t<<<grid_t, FT_NT, 0, st>>>(net, chl, a, 1, chl.T2, 0);
            if (stx != st) cudaStreamWaitEvent(st, evx, 0);
            k_fast_fx<<<grid_t, FT_NT, 0, st>>>(net, chl, a, chl.T2);
            std::swap(chl.T, chl.T2);
            launches += 1;
        } else {
            if (stx != st) cudaStreamWaitEvent(st, evx, 0);
            if (phase_ev) cudaEventRecord(phase_ev[3 * sw + 1], st);
            const bool tl = net.listen && !net.const_params;
            // listening T nodes: the inactive TFs (prior draws, parallel) and the active ones (sequential
            // walk per chain) are disjoint, so the two kernels run side by side when there is an X stream
            cudaStream_t stt = tl ? stx : st;
            k_fast_t<<<grid_t, FT_NT, 0, stt>>>(net, chl, a, net.const_params ? 0 : (tl ? 2 : 1), chl.T, 1);
            if (tl) {
                k_fast_tl<<<nchains, FTL_NT, smem_tl, st>>>(net, chl, a);
                if (stt != st) { cudaEventRecord(evx, stt); cudaStreamWaitEvent(st, evx, 0); }
                launches += 1;
            }
        }
        if (phase_ev) cudaEventRecord(phase_ev[3 * sw + 2], st);
        if (launch_s(net, chl, a, 0, nchains, st, err)) return -1;
        if (stx != st) cudaEventRecord(evs, st);
        launches += 3;
    }
    if (phase_ev) cudaEventRecord(phase_ev[3 * n_sweeps], st);
    if (check(cudaGetLastError(), err)) return -1;
    return launches;
}

}  // namespace gbn
