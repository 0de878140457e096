// Device-resident model description passed by value to every kernel.
//
// HBM layout (all arrays chain-major so that one chain's data is contiguous):
//   network (shared by all chains and datasets)      : CSR arrays, see topology.hpp
//   per dataset   y[d][nH] u8, yprob[d][3] f64
//   per chain     X[c][nX] u8 | T[c][nX] f64 | S[c][Ep] u8 in (padded, transposed) SLOT order
//   statistics    cX1[c][nX] u32  (#sweeps with X=1; X=0 is N - cX1)
//                 tMean[c][nX], tM2[c][nX] f64 (Welford, gsl_rstat order of operations)
//                 cS[c][Ep] u32x2 in slot order: #sweeps with S=0 and with S=2 (S=1 is N - both)
//   fast kernel   tc2[c][nH] {w0, w2} f64x2 flip weights and tL[c][nH] f64 likelihood: the per-target cache of the X phase
//                 Sw[c][nSw] u32 S as 2-bit codes in slot order (the fast sweep's master copy of S)
//                 dC[c][nD] u8x8 per-batch outcome counts of four steps, folded into cS on demand
//                 Stfp[c][Efp/16] u32 copy of S in by-TF order (padded lists, topology.hpp), 2-bit codes (coalesced reads in the X phase)
//                 FXp[c][nX] f64x2   1 - t*x and its reciprocal
// Integer counts replace the reference's gsl_rstat workspaces of 0/1 indicators
// (Multinomial.cpp:101-105): mean = c/N, variance = N/(N-1) p (1-p).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace gbn {

constexpr uint32_t STAB_ROWS = 15;

struct DevNet {
    uint32_t nX, nH, E, Eu;
    uint32_t Ep, n_groups;        // padded slot count, groups of 32 device targets (topology.hpp)
    uint32_t max_indeg, max_outdeg;
    const uint32_t *deg;          // [nH]   in-degree per device target
    const uint32_t *gbase;        // [n_groups+1] first slot of each group; slot(dt, j) = gbase[dt/32] + 32 j + dt%32
    const uint32_t *ref_of_dt;    // [nH]   device target -> reference target index (Philox key of S draws)
    const uint32_t *par_tf;       // [Ep]  slot -> TF (0 for pads)
    const uint32_t *par_edge;     // [Ep]  slot -> edge input index (UINT32_MAX for pads)
    const uint32_t *slot_of_edge; // [E]
    const uint32_t *slot_trg;     // [Ep]  slot -> device target
    const uint8_t *mor_idx;       // [Ep]  slot -> mor + 1 (1 for pads)
    const uint32_t *tf_ptr;       // [nX+1]
    const uint32_t *tf_child;     // [Eu]  device target ids, ascending reference target order
    const uint32_t *tf_slot;      // [Eu]
    const uint32_t *tfpos_of_slot;// [Ep]  slot -> position in the by-TF CSR (fast kernel; needs Eu == E)
    const uint32_t *par_pack;     // [Ep]  slot -> par_tf | (mor+1) << 30 (TF 0, mor+1 = 1 for pads)
    // packed views for the fast S kernel (topology.hpp): 16 steps per 2-bit word, 4 steps per block
    const uint32_t *sbase;        // [n_groups+1] first S word of each group
    const uint32_t *dbase;        // [n_groups+1] first 4-step block of each group
    const uint32_t *sw_grp;       // [nSw / 32]  group of every row of 32 S words
    const uint32_t *dblk_grp;     // [nD / 32]   group of every row of 32 blocks
    const uint32_t *mor2;         // [nSw] prior row per step (3 = pad)
    const uint2 *par16;           // [nD]  four 16-bit parent TFs per block
    const uint32_t *sw_init;      // [nSw] the S words of a freshly initialised chain (S = mor + 1 on every edge, GraphORNOR.cpp:237-263;
                                  // the same for every chain): set_evidence / create fill the packed views from these
    const uint32_t *stfp_init;    // [n_stf_words] likewise the by-TF copy
    const uint4 *grec;            // [n_groups] {gbase, width, sbase, dbase} of a group in one 16-byte record (k_fast_s2)
    const uint2 *chain_map;       // [n_chains] local chain -> {dataset, global chain id} (chain_dataset / chain_global_id tabulated)
    const uint4 *tfpos4;          // [nD]  fp position (padded by-TF lists) of the block's four steps (UINT32_MAX for pads)
    // by-TF lists of the fast path, every TF's list padded to a multiple of four with sentinel children (topology.hpp)
    const uint32_t *fp_ptr;       // [nX+1]
    const uint32_t *fp_child;     // [Efp + 4] device targets; pads and the four entries past the end: the dummy target nH
    const uint32_t *fp_slot;      // [Efp] slot of the edge (UINT32_MAX: pad)
    const uint32_t *fp_abit;      // [Efp] fp position -> 16 * (S word of the edge) + step inside the word (UINT32_MAX: pad)
    uint32_t Efp;
    uint32_t nSw, nD;
    const uint8_t *x_prior_active;// [nX]
    const uint8_t *tf_uf;         // [n_datasets][nX] 1: the product of this TF's children's likelihoods could underflow a double
                                  // (sum_children log min_h YNOISE[h][y] < -600), so the X kernel carries its magnitude
    const double *t_a, *t_b, *t_mean, *t_lnB;   // [nX]
    const uint8_t *y;             // [n_datasets][nH]
    const double *yprob;          // [n_datasets][3]
    const double *stab;           // [n_datasets][STAB_ROWS][stab_D] fast-path tables over n: c0[y=0..2], omz[cls], zc[cls],
                                  // A[y] = omz (N2 - N0), B[y] = omz (N1 - N2), z^n[cls]
    uint32_t stab_D;              // max_indeg + 2 rounded up to even
    uint32_t sprior_safe;         // 1: every S prior row has an entry that is not tiny and the likelihood tables are positive (YPROB >= 0),
                                  // so prior x likelihood cannot vanish for all outcomes: k_fast_s2 drops the fallback of
                                  // Multinomial.cpp:78-85 (set with the evidence, capi.cu; 0 = the S phase runs on k_fast_s)
    uint32_t groups_desc;         // 1: group widths never grow with the group index (targets sorted by in-degree): a CTA of the
                                  // fast S kernel stages only the table rows its first group can reach
    double xprob[2][2];           // [prior_active][v]   GraphORNOR.h:40-41
    double sprob[4][3];           // [mor+1][v]          GraphORNOR.cpp:25; row 3 = {0, 1, 0}: pad steps of the fast S kernel
    double ynoise[3][3];          // [h][y]              GraphORNOR.h:46-48
    double z_y, z_n;              // ZY / ZN constants   GraphORNOR.cpp:155-164
    double rz_y, rz_n;            // their reciprocals (the same correctly rounded quotient the device would compute)
    uint32_t n_datasets, n_graphs_local, chain_offset, n_graphs;
    int32_t listen, const_params;
    int32_t x1;                   // X phase by the single-chain kernel (k_fast_x1: fp32 flip weights of the whole chain in shared memory)
    int32_t sim;                  // simulation mode (empty evidence, GraphORNOR.cpp:28): X and S fixed, H and Y sampled
    uint64_t seed;
};

struct DevChains {
    uint32_t n_chains;            // local: n_datasets * n_graphs_local, dataset-major
    uint8_t *X;
    double *T;
    double *T2;                   // fast kernel: alternate T buffer (the T draws of a sweep run beside its X phase)
    uint8_t *S;
    uint32_t *cX1;
    double *tMean, *tM2;
    uint2 *cS;                    // {count S=0, count S=2}
    // fast kernel only
    double2 *tc2;                 // [c/XQ][nH + 1][XQ] (interleaved by chain bundle; entry nH is a dummy {0, 0} for idle lanes)
    double *tL;                   // [c/XQ][nH + 1][XQ] the target likelihood L itself (dummy: 1)
    uint8_t *tq;                  // [c/XQ][nH + 1][XQ] q with L >= 2^-q (from L's exponent field; dummy: 0): the X kernel's cheap
                                  // bound on the magnitude of a product of child likelihoods flip weights {w0, w2} = {pr0 omz (a + pr2 b), pr0 pr2 omz b} / L (sweep_fast.cu, k_fast_x)
    uint32_t *Sw;                 // [c][nSw] S as 2-bit codes, 16 steps per word: what the fast sweep reads and writes
                                  // (ch.S is refreshed from it on demand, capi.cu)
    uint32_t *Aw;                 // [c][nSw] bit u of a word: the parent TF of step u of the matching S word is active (X = 1);
                                  // kept current by the X kernel's flip patch, so the S kernel looks a TF's factor up only then
    uint2 *dC;                    // [c][nD] per-batch counts {S=0, S=2} of four steps as 4 x u8 each, folded into cS
                                  // before anything reads the statistics (and before a byte could overflow)
    uint32_t *Stfp;               // [c][n_stf_words] copy of S in by-TF order (padded lists), 2-bit codes (coalesced reads in the X phase)
    float *w0f, *w2f;             // [c][n_w32(net)] the flip weights rounded to fp32, one contiguous row per chain (net.x1): what the
                                  // single-chain X kernel stages in shared memory (entry nH .. : 0)
    uint8_t *tq1;                 // [c][n_q8(net)] the bound bytes of tq, one contiguous row per chain (net.x1; entry nH .. : 0)
    uint16_t *xqp;                // [c][nX] X kernel: the magnitude bound (bits) a TF's children gave at its last update - predicts
                                  // which TFs will need the exact product, so that it is gathered in the same pass
    uint32_t *xU;                 // [c][nX rounded up to 4] Philox words of the current sweep's X draws (X kernel scratch)
    double2 *FXp;                 // {1 - t*x, 1 / (1 - t*x)}
    // simulation mode only: latent / observed state per (chain, device target) and their outcome counts
    uint8_t *H, *Y;
    uint2 *cH, *cY;               // {count outcome 0, count outcome 2}
    unsigned long long *dbg;      // [8] diagnostics (gbn_model_debug_counters)
    volatile uint32_t *flags;     // host-mapped word: bit 0 = a T node reached exactly 1.0 on the fast path (factor 1 - t x = 0:
                                  // its reciprocal does not exist; the model layer reports it, capi.cu)
    // injected draws (NULL = keyed Philox streams)
    const double *inj_flat;       // [chain][inj_sweeps][nX + E]  X then S in edge input order
    const double *inj_beta;       // [chain][inj_sweeps][nX]
    const double *inj_exp;        // [chain][inj_sweeps][nX]
    uint32_t inj_sweeps;          // sweeps available
    uint32_t inj_pos;             // first unconsumed injected sweep
};

// local chain index -> dataset and global chain id (Philox counter word 2)
// first slot of device target dt; its j-th parent is 32 j further
__device__ __forceinline__ uint32_t slot0(const DevNet &n, uint32_t dt) { return n.gbase[dt >> 5] + (dt & 31); }

// words of the by-TF copy of the S codes (16 codes per word): fp positions 0 .. Efp-1 and the four sentinel positions
// Efp .. Efp+3 (code 1, child = the dummy target nH) that idle lanes of the X kernels' child loops read instead of branching
__host__ __device__ inline uint32_t n_stf_words(const DevNet &n) { return (n.Efp + 19) >> 4; }

// row lengths of the per-chain fp32 weight rows / bound-byte rows (targets 0 .. nH-1, the dummy nH; rounded so that
// every row starts on a 16-byte boundary: the rows are copied to shared memory with cp.async.bulk)
__host__ __device__ inline uint32_t n_w32(const DevNet &n) { return (n.nH + 4) & ~3u; }
__host__ __device__ inline uint32_t n_q8(const DevNet &n) { return (n.nH + 16) & ~15u; }

__host__ __device__ inline uint32_t chain_dataset(const DevNet &n, uint32_t c) { return c / n.n_graphs_local; }
__host__ __device__ inline uint32_t chain_global_id(const DevNet &n, uint32_t c)
{
    uint32_t d = c / n.n_graphs_local, l = c % n.n_graphs_local;
    return d * n.n_graphs + n.chain_offset + l;
}

}  // namespace gbn
