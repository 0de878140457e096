// SMALL-NETWORK sweep path: one persistent CTA per chain, the chain's whole state in shared memory for all the
// sweeps of a batch.
//
// Replaces the loop of GraphBase::sample (GraphBase.cpp:22-27) - dN sweeps X -> T -> S of one chain - by ONE
// kernel launch per batch: no host launch, no trip through HBM between sweeps.  For networks up to a few
// thousand targets (BASELINE configs C1, C2 and the batched contrasts of C4) the multi-kernel path of
// sweep_fast.cu is bound by launch and round latency (31 us per sweep at C1, 75 us at C2, 4 launches per sweep);
// here the product cache, the 2-bit S codes (by target and by TF), X, T, the per-batch 8-bit S counters and
// the tables live in shared memory, a gather is a shared-memory load, and HBM is touched when the batch
// starts (state in) and when it ends (state, counters out).
//
// Same chain as the other kernels: same visiting order, same keyed Philox draws, same algebra as sweep_fast.cu
//   X phase  speculative windows over consecutive TFs, committed up to the first TF that changed
//            (the sequential chain of GraphBase.cpp:24-26, not a coloured variant)
//   T phase  Beta::metropolis_hastings_with_prior_proposal for non-listening T nodes (Beta.cpp:60-91)
//   S phase  one warp per group of 32 targets, lane = target, leave-one-out walk along the target's edges
// Not handled here (the model layer keeps these on sweep_fast.cu): T nodes that listen to their children.
#include <cstdio>
#include "kernels.cuh"
#include "device_math.cuh"

#include "topology.hpp"
#include "fast_common.cuh"

namespace gbn {

namespace {

constexpr int SM_NT = 1024;                  // threads per CTA (one chain): every phase is bound by the latency of a
                                             // thread's dependent chain, so the CTA takes all the warps an SM gives one block

struct SmallArgs {
    SweepArgs sa;                            // sweep = first sweep of the batch, stat_n = statistics count BEFORE the batch
    uint32_t n_sweeps;
    uint32_t smem_bytes;
};

// carve-up of the dynamic shared memory (offsets in bytes, every array 16-byte aligned)
struct SmallLayout {
    uint32_t tc, fx, tab, zp, T, dC, Sw, Stf, ptr, cx1, xu, ny, X, dec, fac, total;
};

__host__ __device__ inline uint32_t al16(uint32_t v) { return (v + 15u) & ~15u; }

__host__ __device__ inline SmallLayout small_layout(const DevNet &net, uint32_t W)
{
    SmallLayout L;
    const uint32_t nX = net.nX, nH = net.nH, D = net.stab_D;
    uint32_t o = 0;
    L.tc = o;  o = al16(o + 16 * (nH + 1));                  // double2 {beta, gamma} per target; [nH] = dummy {0, 0}
    L.fx = o;  o = al16(o + 16 * nX);                        // double2 {1 - t x, 1 / (1 - t x)}
    L.tab = o; o = al16(o + 16 * (3 * D + 1));               // double2 {c0[y][n], omz[class of y][n]}; [3 D] = {1, 0} (dummy)
    L.zp = o;  o = al16(o + 8 * 2 * D);                      // (unused rows kept for the layout of sweep_fast's tables)
    L.T = o;   o = al16(o + 8 * nX);
    L.dC = o;  o = al16(o + 8 * net.nD);                     // per-batch counts {S = 0, S = 2} of four steps, 4 x u8 each
    L.Sw = o;  o = al16(o + 4 * net.nSw);                    // S codes by target, 16 steps per word
    L.Stf = o; o = al16(o + 4 * n_stf_words(net));        // S codes in by-TF order
    L.ptr = o; o = al16(o + 4 * (nX + 1));
    L.cx1 = o; o = al16(o + 4 * nX);                         // sweeps of this batch with X = 1
    L.xu = o;  o = al16(o + 4 * ((nX + 3) & ~3u));           // Philox words of the sweep's X draws
    L.ny = o;  o = al16(o + 2 * (nH + 1));                   // table index y D + n per target; [nH] = 3 D
    L.X = o;   o = al16(o + nX);                             // value | prior class << 1
    L.dec = o; o = al16(o + 4 * W);
    L.fac = o; o = al16(o + 8 * W);
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------ S phase of one group
// The S kernel of sweep_fast.cu with every per-chain array in shared memory (see there for the algebra).
// MODE 1 rebuilds the product cache only (batch start).
template <int MODE>
__device__ __forceinline__ void small_s_group(const DevNet &net, const DevChains &ch, const SweepArgs &a, uint32_t sweep, bool use_inj,
                                              const double *inj_flat, uint32_t c, uint32_t d, uint32_t gchain, uint32_t g, uint32_t lane,
                                              const double2 *sTab, const double2 *sFX, double2 *sTc, uint16_t *sNy,
                                              uint32_t *sSw, uint32_t *sStf, uint2 *sDC)
{
    const uint32_t nH = net.nH, D = net.stab_D;
    const uint32_t dt = 32 * g + lane;
    const bool valid = dt < nH;
    const uint32_t gb = net.gbase[g];
    const uint32_t W = (net.gbase[g + 1] - gb) >> 5;
    const uint32_t y = valid ? net.y[(size_t) d * nH + dt] : 1u;
    const double z = (y != 1) ? net.z_y : net.z_n;
    const double ca = net.ynoise[2][y] - net.ynoise[0][y];
    const double cb = net.ynoise[1][y] - net.ynoise[2][y];
    const double2 *cot = sTab + y * D;
    uint32_t *pSw = sSw + net.sbase[g] + lane;
    const uint16_t *pPar = reinterpret_cast<const uint16_t *>(net.par16 + net.dbase[g] + lane);   // step j: [128 (j / 4) + j % 4]
    const uint32_t W4 = (W + 3) >> 2;

    double A = ca, Bq = cb;
    uint32_t n = 0;
    for (uint32_t b = 0; b < W4; b++) {
        const uint32_t sw = pSw[32 * (b >> 2)] >> (8 * (b & 3u));
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t s = (sw >> (2 * u)) & 3u;
            const uint32_t kk = pPar[128 * b + u];
            GBN_CHECK(kk < net.nX && s <= 2);
            const double f = sFX[kk].x * z;
            n += (s != 1);
            A *= (s == 0) ? f : 1.;
            Bq *= (s != 1) ? f : 1.;
        }
    }
    if (MODE == 0) {
        const double rz = 1. / z;
        const uint32_t ref_i = valid ? net.ref_of_dt[dt] : 0u;
        const uint32_t *pEdge = net.par_edge + gb + lane;
        const uint32_t *pMor = net.mor2 + net.sbase[g] + lane;
        const uint4 *pTp = net.tfpos4 + net.dbase[g] + lane;
        uint2 *pC = sDC + net.dbase[g] + lane;
        uint32_t sw0 = 0, sw = 0, mw = 0, chg = 0;
        for (uint32_t bb = 0; bb < W4; bb++) {
            if ((bb & 3u) == 0) { sw0 = pSw[8 * bb]; sw = sw0; mw = pMor[8 * bb]; chg = 0; }
            const uint32_t j0 = 4 * bb;
            uint2 cw = pC[32 * bb];
            u32x4 rnd = { 0, 0, 0, 0 };
            if (!use_inj) rnd = philox_rk(ref_i, sweep, gchain, KIND_S | (bb << 8), a);
            uint32_t chg4 = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (j0 + u < W) {                                    // warp-uniform
                    const uint32_t s = (sw >> (2 * u)) & 3u, r = (mw >> (2 * u)) & 3u;
                    const uint32_t kk = pPar[128 * bb + u];
                    GBN_CHECK(kk < net.nX && s <= 2 && n + 1 < D);
                    const double2 fx = sFX[kk];
                    const double f = fx.x * z, finv = fx.y * rz;
                    const double Af = A * finv, Bf = Bq * finv;      // the state without this edge
                    const double Ar = s == 0 ? Af : A, Br = s != 1 ? Bf : Bq;
                    const uint32_t nR = n - (s != 1);
                    const double2 t0 = cot[nR], t1 = cot[nR + 1];
                    const double *sp = net.sprob[r];                 // row 3: a pad step, never moves
                    const double sum = Ar + Br;
                    const double L0 = __fma_rn(t1.y, f * sum, t1.x);
                    const double L1 = __fma_rn(t0.y, sum, t0.x);
                    const double L2 = __fma_rn(t1.y, __fma_rn(f, Br, Ar), t1.x);
                    double cum0 = sp[0] * L0;
                    double cum1 = __fma_rn(sp[1], L1, cum0);
                    double cum2 = __fma_rn(sp[2], L2, cum1);
                    if (!(cum2 > 0.)) { cum0 = sp[0]; cum1 = cum0 + sp[1]; cum2 = cum1 + sp[2]; }     // Multinomial.cpp:78-85
                    const uint32_t word = u == 0 ? rnd.x : (u == 1 ? rnd.y : (u == 2 ? rnd.z : rnd.w));
                    double uu = unit_from_word(word);
                    if (use_inj) { const uint32_t e = pEdge[32 * (j0 + u)]; uu = e != 0xffffffffu ? inj_flat[e] : 0.5; }
                    const double dice = cum2 * uu;
                    uint32_t ns = dice < cum0 ? 0u : (dice < cum1 ? 1u : (dice < cum2 ? 2u : s));
                    cw.x += (uint32_t) (ns == 0) << (8 * u);         // Multinomial.cpp:101-105 as counts
                    cw.y += (uint32_t) (ns == 2) << (8 * u);
                    const uint32_t x = s ^ ns;
                    if (x) {
                        chg4 |= x << (2 * u);
                        A = ns == 0 ? Ar * f : Ar;
                        Bq = ns != 1 ? Br * f : Br;
                        n = nR + (ns != 1);
                        const uint32_t tp = u == 0 ? pTp[32 * bb].x : (u == 1 ? pTp[32 * bb].y : (u == 2 ? pTp[32 * bb].z : pTp[32 * bb].w));
                        GBN_CHECK(tp < net.Efp);
                        atomicXor(sStf + (tp >> 4), x << (2 * (tp & 15)));
                    }
                }
            }
            pC[32 * bb] = cw;
            chg |= chg4 << (8 * (bb & 3u));
            sw >>= 8; mw >>= 8;
            if (((bb & 3u) == 3 || bb + 1 == W4) && chg) pSw[8 * (bb & ~3u)] = sw0 ^ chg;
        }
    }
    GBN_CHECK(n + (y * D) < 3 * D);
    if (valid) {
        sTc[dt] = make_double2(cot[n].y * A, cot[n].y * Bq);
        sNy[dt] = (uint16_t) (y * D + n);
    }
}

// ------------------------------------------------------------------------------------------ the kernel
// LPT lanes of a warp work on one TF of the X window (window = SM_NT / LPT consecutive TFs)
template <int LPT>
__global__ void __launch_bounds__(SM_NT, 1) k_small_sweep(DevNet net, DevChains ch, SmallArgs sa)
{
    constexpr int NW = SM_NT / 32, NS = 32 / LPT, W = NW * NS;
    extern __shared__ __align__(16) unsigned char smem[];
    const SweepArgs &a = sa.sa;
    const uint32_t nX = net.nX, nH = net.nH, D = net.stab_D, E = net.E;
    const uint32_t nXr = (nX + 3) & ~3u, nStfW = n_stf_words(net);
    const SmallLayout L = small_layout(net, W);
    double2 *sTc = reinterpret_cast<double2 *>(smem + L.tc);
    double2 *sFX = reinterpret_cast<double2 *>(smem + L.fx);
    double2 *sTab = reinterpret_cast<double2 *>(smem + L.tab);
    double *sT = reinterpret_cast<double *>(smem + L.T);
    uint2 *sDC = reinterpret_cast<uint2 *>(smem + L.dC);
    uint32_t *sSw = reinterpret_cast<uint32_t *>(smem + L.Sw);
    uint32_t *sStf = reinterpret_cast<uint32_t *>(smem + L.Stf);
    uint32_t *sPtr = reinterpret_cast<uint32_t *>(smem + L.ptr);
    uint32_t *sCX1 = reinterpret_cast<uint32_t *>(smem + L.cx1);
    uint32_t *sU = reinterpret_cast<uint32_t *>(smem + L.xu);
    uint16_t *sNy = reinterpret_cast<uint16_t *>(smem + L.ny);
    uint8_t *sX = smem + L.X;
    uint32_t *s_dec = reinterpret_cast<uint32_t *>(smem + L.dec);
    double *s_fac = reinterpret_cast<double *>(smem + L.fac);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c = blockIdx.x + a.chain0;
    const uint32_t d = chain_dataset(net, c), gchain = chain_global_id(net, c);
    uint8_t *gX = ch.X + (size_t) c * nX;
    double *gT = ch.T + (size_t) c * nX;
    uint32_t *gSw = ch.Sw + (size_t) c * net.nSw;

    // ---- batch start: state in, caches rebuilt
    for (uint32_t k = tid; k < nX; k += SM_NT) {
        const uint32_t x = gX[k];
        const double t = gT[k];
        sX[k] = (uint8_t) (x | (net.x_prior_active[k] << 1));
        sT[k] = t;
        const double fx = 1. - t * (double) x;
        sFX[k] = make_double2(fx, 1. / fx);
        sCX1[k] = 0;
    }
    for (uint32_t k = tid; k <= nX; k += SM_NT) sPtr[k] = net.fp_ptr[k];
    for (uint32_t i = tid; i < net.nSw; i += SM_NT) sSw[i] = gSw[i];
    for (uint32_t i = tid; i < net.nD; i += SM_NT) sDC[i] = make_uint2(0u, 0u);
    {
        const double *tab = net.stab + (size_t) d * STAB_ROWS * D;
        for (uint32_t j = tid; j < 3 * D; j += SM_NT) {
            const uint32_t yy = j / D, nn = j - yy * D;
            sTab[j] = make_double2(tab[j], tab[(3 + (yy != 1 ? 0 : 1)) * D + nn]);
        }
        if (tid == 0) { sTab[3 * D] = make_double2(1., 0.); sTc[nH] = make_double2(0., 0.); sNy[nH] = (uint16_t) (3 * D); }
    }
    __syncthreads();
    for (uint32_t wi = tid; wi < nStfW; wi += SM_NT) {           // by-TF copy of the S codes
        uint32_t w = 0x55555555u;
        for (uint32_t u = 0; u < 16; u++) {
            const uint32_t pos = 16 * wi + u;
            const uint32_t ab = pos < net.Efp ? net.fp_abit[pos] : UINT32_MAX;     // pads of the by-TF lists keep code 1
            if (ab != UINT32_MAX) w = (w & ~(3u << (2 * u))) | (((sSw[ab >> 4] >> (2 * (ab & 15))) & 3u) << (2 * u));
        }
        sStf[wi] = w;
    }
    for (uint32_t g = warp; g < net.n_groups; g += NW)
        small_s_group<1>(net, ch, a, 0, false, nullptr, c, d, gchain, g, lane, sTab, sFX, sTc, sNy, sSw, sStf, sDC);
    __syncthreads();

    const uint32_t sg = lane / LPT, sl = lane % LPT;             // this lane's TF of the warp, lane within it
    const uint32_t pos = warp * NS + sg;                         // window position of this lane's TF
    const bool tupd = !net.const_params;

    for (uint32_t it = 0; it < sa.n_sweeps; it++) {
        const uint32_t sweep = a.sweep + it;
        const bool use_inj = ch.inj_flat != nullptr && (ch.inj_pos + it) < ch.inj_sweeps;
        const size_t inj_row = (size_t) c * ch.inj_sweeps + ch.inj_pos + it;
        const double *inj = use_inj ? ch.inj_flat + inj_row * (nX + E) : nullptr;

        long long tck = clock64();
        auto lap = [&](int ph) {        // diagnostics (gbn_model_debug_counters): cycles of thread 0 per phase
            if (tid == 0 && ch.dbg) { const long long t1 = clock64(); atomicAdd(ch.dbg + 4 + ph, (unsigned long long) (t1 - tck)); tck = t1; }
        };
        // ------------------------------------------------------------ X phase
        if (!use_inj)
            for (uint32_t blk = tid; blk < nXr / 4; blk += SM_NT) {
                const u32x4 r = philox_rk(blk, sweep, gchain, KIND_X, a);
                reinterpret_cast<uint4 *>(sU)[blk] = make_uint4(r.x, r.y, r.z, r.w);
            }
        __syncthreads();
        uint32_t k0 = 0;
        while (k0 < nX) {
            const uint32_t kw = k0 + pos;
            const bool live = kw < nX;
            int dec = 0;                              // bit 0: drawn value, bit 1: value changed
            double fac = 1.;
            if (NS > 1 || live) {
                const uint32_t kr = live ? kw : 0u;
                const uint32_t xcur = sX[kr] & 1u;
                const double g1 = 1. - sT[kr];        // (1 - t*x) at x = 1   (HNodeORNOR.cpp:59)
                fac = g1;
                if (xcur) fac = 1. / g1;              // factor that turns the product into the flipped one
                uint32_t cbeg = 0, cend = 0;
                if (live) { cbeg = sPtr[kw]; cend = sPtr[kw + 1]; }
                double mc = 1., mf = 1.;
                int ec = 0, ef = 0, cnt = 0;
                uint32_t trips = (cend - cbeg + LPT - 1) / LPT;
                if (NS > 1) trips = __reduce_max_sync(0xffffffffu, trips);
                uint32_t cc = cbeg + sl;
                for (; trips > 0; trips--, cc += LPT) {
                    uint32_t id = nH, s = 1u;                         // idle lane: the dummy target, edge off
                    if (cc < cend) { id = net.fp_child[cc]; s = code_at(sStf, cc); }
                    GBN_CHECK(id <= nH && s <= 2);
                    const double2 bg = sTc[id];
                    const double c0 = sTab[sNy[id]].x;
                    const double2 m = make_double2(s == 0 ? fac : 1., s != 1 ? fac : 1.);
                    const double Lc = c0 + (bg.x + bg.y);
                    const double Lf = c0 + __fma_rn(m.x, bg.x, m.y * bg.y);
                    mc *= Lc; mf *= Lf;
                    if (++cnt >= 16) { normalise(mc, ec); normalise(mf, ef); cnt = 0; }
                }
                normalise(mc, ec);
                normalise(mf, ef);
#pragma unroll
                for (int o = LPT / 2; o >= 1; o >>= 1) {
                    mc *= __shfl_xor_sync(0xffffffffu, mc, o);
                    mf *= __shfl_xor_sync(0xffffffffu, mf, o);
                    ec += __shfl_xor_sync(0xffffffffu, ec, o);
                    ef += __shfl_xor_sync(0xffffffffu, ef, o);
                }
                normalise(mc, ec);
                normalise(mf, ef);
                const double *prob = net.xprob[sX[kr] >> 1];
                double lik_cur, lik_flip;
                common_scale(prob[xcur] * mc, ec, prob[1 - xcur] * mf, ef, lik_cur, lik_flip);
                const double l0 = xcur ? lik_flip : lik_cur, l1 = xcur ? lik_cur : lik_flip;
                double cum0 = 0. + l0, cum1 = cum0 + l1;
                if (!(cum1 > 0.)) { cum0 = 0. + prob[0]; cum1 = cum0 + prob[1]; }     // Multinomial.cpp:78-85
                const double u = use_inj ? inj[kr] : unit_from_word(sU[kr]);
                const double dice = 0. * (1 - u) + cum1 * u;
                const uint32_t nx = dice < cum0 ? 0u : (dice < cum1 ? 1u : xcur);
                dec = live ? ((int) nx | ((nx != xcur) ? 2 : 0)) : 0;
            }
            if (sl == 0) { s_dec[pos] = (uint32_t) dec; s_fac[pos] = fac; }
            __syncthreads();
            uint32_t first = W;
#pragma unroll
            for (uint32_t w0 = 0; w0 < (uint32_t) W; w0 += 32) {
                const uint32_t f = __ballot_sync(0xffffffffu, w0 + lane < (uint32_t) W && (s_dec[(w0 + lane) % W] & 2u) != 0);
                if (f && first == (uint32_t) W) first = w0 + (uint32_t) __ffs(f) - 1;
            }
            const uint32_t ncommit = first < W ? first + 1 : W;
            if (first < (uint32_t) W) {
                // the window ends at the TF that flipped: apply the flip to the cached terms of its children
                const uint32_t kf = k0 + first, fb0 = sPtr[kf], fe0 = sPtr[kf + 1];
                const double fq = s_fac[first];
                for (uint32_t c2 = fb0 + tid; c2 < fe0; c2 += SM_NT) {
                    const uint32_t s = code_at(sStf, c2);
                    if (s == 1) continue;
                    double2 bg = sTc[net.fp_child[c2]];
                    if (s == 0) bg.x *= fq;
                    bg.y *= fq;
                    sTc[net.fp_child[c2]] = bg;
                }
            }
            __syncthreads();
            if (pos < ncommit && live && sl == 0) sX[kw] = (uint8_t) ((sX[kw] & 2u) | (dec & 1));
            if (tid == 0 && ch.dbg) {
                atomicAdd(ch.dbg + 0, 1ull);
                atomicAdd(ch.dbg + 1, (unsigned long long) (k0 + ncommit <= nX ? ncommit : nX - k0));
                atomicAdd(ch.dbg + 2, first < W ? 1ull : 0ull);
            }
            k0 += ncommit;
        }
        __syncthreads();
        lap(0);

        // ------------------------------------------------------------ T phase (Beta.cpp:60-104) and FXp
        for (uint32_t k = tid; k < nX; k += SM_NT) {
            const uint32_t x = sX[k] & 1u;
            sCX1[k] += x;                                             // Multinomial.cpp:101-105 as a count
            double t = sT[k];
            if (tupd) {
                const double al = net.t_a[k], be = net.t_b[k], mean = net.t_mean[k], lnB = net.t_lnB[k];
                const size_t row = inj_row * nX + k;
                const double draw = use_inj ? ch.inj_beta[row] : draw_beta(net.seed, gchain, sweep, k, al, be);
                t = mh_prior_step(t, draw, mean, al, be, lnB, use_inj, use_inj ? ch.inj_exp[row] : 0., net.seed, gchain, sweep, k);
                sT[k] = t;
                if (t == 1.) *ch.flags = 1u;
                const size_t o = (size_t) c * nX + k;
                double m = ch.tMean[o], m2 = ch.tM2[o];
                welford_add(m, m2, t, (double) (a.stat_n + it + 1));
                ch.tMean[o] = m; ch.tM2[o] = m2;
            }
            const double fx = 1. - t * (double) x;
            sFX[k] = make_double2(fx, 1. / fx);
        }
        __syncthreads();
        lap(1);

        // ------------------------------------------------------------ S phase
        for (uint32_t g = warp; g < net.n_groups; g += NW)
            small_s_group<0>(net, ch, a, sweep, use_inj, use_inj ? inj + nX : nullptr, c, d, gchain, g, lane,
                             sTab, sFX, sTc, sNy, sSw, sStf, sDC);
        __syncthreads();
        lap(2);
    }

    // ---- batch end: state and counters out (the per-batch counts are ADDED: bytes cannot carry, the model
    // layer folds them into the 32-bit totals before 255 sweeps have been counted)
    for (uint32_t k = tid; k < nX; k += SM_NT) {
        gX[k] = (uint8_t) (sX[k] & 1u);
        gT[k] = sT[k];
        ch.cX1[(size_t) c * nX + k] += sCX1[k];
    }
    for (uint32_t i = tid; i < net.nSw; i += SM_NT) gSw[i] = sSw[i];
    uint2 *gDC = ch.dC + (size_t) c * net.nD;
    for (uint32_t i = tid; i < net.nD; i += SM_NT) {
        const uint2 v = sDC[i];
        if (v.x | v.y) { uint2 t = gDC[i]; t.x += v.x; t.y += v.y; gDC[i] = t; }
    }
}

template <int LPT>
int launch_small_t(const DevNet &net, const DevChains &ch, const SmallArgs &sa, uint32_t nchains, cudaStream_t st, const char **err)
{
    cudaError_t e = cudaFuncSetAttribute(k_small_sweep<LPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sa.smem_bytes);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    k_small_sweep<LPT><<<nchains, SM_NT, sa.smem_bytes, st>>>(net, ch, sa);
    return 0;
}

uint32_t small_lpt(const DevNet &net)
{
    uint32_t lpt = env_u32("GBN_SMALL_LPT", 0);               // 4 / 8 / 16 / 32 force a path (tests)
    if (lpt != 4 && lpt != 8 && lpt != 16 && lpt != 32) {
        const double mean = (double) net.E / (double) net.nX;    // children per TF
        lpt = mean >= 48. ? 32 : (mean >= 20. ? 16 : (mean >= 8. ? 8 : 4));
    }
    return lpt;
}

}  // namespace

// the whole chain state fits one CTA's shared memory, and nothing needs the multi-kernel path
bool small_supported(const DevNet &net, const char **why)
{
    const char *w2 = nullptr;
    if (!fast_supported(net, &w2)) { *why = w2; return false; }
    if (net.listen && !net.const_params) { *why = "T nodes listen to their children"; return false; }
    if (env_u32("GBN_SMALL", 1) == 2) { *why = "disabled (GBN_SMALL=2)"; return false; }
    const uint32_t W = SM_NT / small_lpt(net);
    if (small_layout(net, W).total + 1024 > 200 * 1024) { *why = "state does not fit one SM's shared memory"; return false; }
    return true;
}

// n_sweeps sweeps X -> T -> S of local chains [chain0, chain0 + nchains) in ONE launch; returns the number of
// kernels launched (1), or -1 with *err set
int launch_sweep_small(const DevNet &net, const DevChains &ch, uint32_t n_sweeps, uint32_t sweep0, uint32_t stat_n0,
                       cudaStream_t st, const char **err, uint32_t chain0, uint32_t nchains)
{
    if (nchains == UINT32_MAX) nchains = ch.n_chains - chain0;
    if (n_sweeps == 0 || nchains == 0) return 0;
    const uint32_t lpt = small_lpt(net);
    SmallArgs sa;
    sa.sa = sweep_args(net);
    sa.sa.sweep = sweep0; sa.sa.stat_n = stat_n0; sa.sa.chain0 = chain0;
    sa.n_sweeps = n_sweeps;
    sa.smem_bytes = small_layout(net, SM_NT / lpt).total;
    int rc;
    if (lpt == 32) rc = launch_small_t<32>(net, ch, sa, nchains, st, err);
    else if (lpt == 16) rc = launch_small_t<16>(net, ch, sa, nchains, st, err);
    else if (lpt == 8) rc = launch_small_t<8>(net, ch, sa, nchains, st, err);
    else rc = launch_small_t<4>(net, ch, sa, nchains, st, err);
    if (rc) return -1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

}  // namespace gbn
