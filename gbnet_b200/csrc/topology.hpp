// Host-side topology compiler: integer edge list + evidence -> device arrays and priors.
//
// Replaces GraphORNOR::build_structure (reference libgbnet/src/GraphORNOR.cpp:13-265), which
// builds one heap object graph PER CHAIN (ModelORNOR.cpp:23-27).  Here the network is compiled
// once into flat arrays shared by every chain and every dataset.
//
// Orders.  The reference's orders are what the ABI speaks: TFs and targets in sorted-id order
// (std::set<string>, GraphORNOR.cpp:31-42), S nodes in edge input order.  On the device, targets are
// RE-LABELLED by descending in-degree ("device targets", dt) so that the 32 lanes of a warp own
// targets with (almost) equal numbers of parents, and the slots of each GROUP of 32 consecutive
// device targets are laid out transposed:
//
//     slot(dt, j) = gbase[dt / 32] + 32 * j + dt % 32,     j = 0 .. deg[dt]-1 in edge input order
//                                                          (= HNodeORNOR::append_parent order,
//                                                          GraphORNOR.cpp:237-263)
//
// so that when the lanes of a warp all visit their j-th parent, S bytes, packed parent ids and the
// statistics counters are read and written as consecutive addresses.  A group is padded to the
// in-degree of its first (largest) target; pad slots carry par_edge = UINT32_MAX, S = 1 (edge off)
// and are never referenced by the by-TF lists.  The host permutes at the boundary (state get/set
// through slot_of_edge, evidence through dt_of_ref, per-target outputs through ref_of_dt).
//
//   by-TF CSR      tf_ptr / tf_child / tf_slot : the UNIQUE targets of each TF in ascending
//                  REFERENCE target order (HParentNode::children is a std::set, HParentNode.h:18),
//                  stored as device target ids, with the slot of the (first) connecting edge.
//   per TF         Beta prior (a, b, mean, ln B) from the log out-degree (GraphORNOR.cpp:213-223)
//   per dataset    observed state y per target, YPROB (GraphORNOR.cpp:113-134)
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace gbn {

inline uint32_t env_u32(const char *name, uint32_t dflt)
{
    const char *v = std::getenv(name);
    if (!v || !*v) return dflt;
    long x = std::strtol(v, nullptr, 10);
    return x > 0 ? (uint32_t) x : dflt;
}

struct Topology {
    uint32_t nX = 0, nH = 0, E = 0, Eu = 0;
    uint32_t Ep = 0;                                // padded slot count
    uint32_t n_groups = 0;                          // ceil(nH / 32)
    uint32_t max_indeg = 0, max_outdeg = 0;
    std::vector<uint32_t> tf_uid, trg_uid;          // sorted unique ids (reference order)
    std::vector<uint32_t> trg_lut;                  // id -> rank in trg_uid when the ids are dense enough (else empty)
    std::vector<uint32_t> dt_of_ref, ref_of_dt;     // reference target index <-> device target index
    std::vector<uint32_t> deg;                      // [nH] in-degree per device target
    std::vector<uint32_t> gbase;                    // [n_groups+1] first slot of each group
    // per slot [Ep]
    std::vector<uint32_t> par_tf, par_edge, slot_trg, tfpos_of_slot, par_pack;
    std::vector<uint8_t> mor_idx;                   // mor + 1 (1 for pads)
    std::vector<uint32_t> slot_of_edge;             // [E]
    std::vector<uint32_t> tf_ptr, tf_child, tf_slot;
    std::vector<uint32_t> n_h_child;                // edges per TF (GraphORNOR.cpp:206-211)
    std::vector<uint8_t> x_prior_active;            // per TF
    std::vector<double> t_a, t_b, t_mean, t_lnB;    // per TF
    // packed views of the same slots for the fast S kernel (sweep_fast.cu): a lane's steps j = 0 .. W-1 are
    // grouped into words of 16 steps (2-bit S codes / prior rows) and blocks of 4 steps (16-bit parent ids,
    // 8-bit per-batch counter pairs):
    //     S word      sbase[g] + 32 * (j / 16) + lane,  code at bits 2 (j % 16)
    //     block       dbase[g] + 32 * (j / 4) + lane,   entry j % 4
    std::vector<uint32_t> sbase, dbase;             // [n_groups+1]
    std::vector<uint32_t> sw_grp, dblk_grp;         // group of every row of 32 words / 32 blocks
    std::vector<uint32_t> mor2;                     // [nSw] prior row per step, 3 = pad (row {0, 1, 0}: never moves)
    std::vector<uint16_t> par16;                    // [4 nD] parent TF per step (0 for pads)
    std::vector<uint32_t> tfpos4;                   // [4 nD] by-TF CSR position per step (UINT32_MAX for pads)
    // by-TF lists of the FAST path: every TF's list starts at a multiple of four and is padded to one with sentinel
    // children (the dummy target nH, slot UINT32_MAX; their S code in the by-TF copy stays 1 = edge off), so that a lane
    // of the X kernels reads four children with one 16-byte load and their four S codes with one byte load.  Positions
    // in these lists ("fp" positions) index the by-TF copy of S (Stfp), tfpos4 and fp_abit.
    std::vector<uint32_t> fp_ptr;                   // [nX+1] multiples of 4; fp_ptr[nX] = Efp
    std::vector<uint32_t> fp_child, fp_slot;        // [Efp + 4] (four sentinels at the end: what idle lanes read), [Efp]
    std::vector<uint32_t> fp_abit;                  // [Efp] fp position -> 16 * S word + step inside the word (UINT32_MAX: pad)
    uint32_t Efp = 0;
    uint32_t nSw = 0, nD = 0;

    uint32_t slot(uint32_t dt, uint32_t j) const { return gbase[dt >> 5] + 32 * j + (dt & 31); }
};

inline uint32_t rank_of(const std::vector<uint32_t> &sorted, uint32_t v)
{
    auto it = std::lower_bound(sorted.begin(), sorted.end(), v);
    if (it == sorted.end() || *it != v) return UINT32_MAX;
    return (uint32_t) (it - sorted.begin());
}

inline Topology compile_topology(uint32_t n_edge, const uint32_t *src, const uint32_t *trg,
                                 const int32_t *mor, uint32_t n_active, const uint32_t *active_tf,
                                 double t_focus, double t_lmargin, double t_hmargin)
{
    Topology tp;
    if (n_edge == 0) throw std::invalid_argument("network has no edges");
    for (uint32_t e = 0; e < n_edge; e++)
        if (mor[e] != -1 && mor[e] != 0 && mor[e] != 1)                 // GraphORNOR.cpp:39-41
            throw std::out_of_range("MOR values can only be either -1, 0 or 1");
    tp.E = n_edge;
    tp.tf_uid.assign(src, src + n_edge);
    tp.trg_uid.assign(trg, trg + n_edge);
    for (auto *v : { &tp.tf_uid, &tp.trg_uid }) {
        std::sort(v->begin(), v->end());
        v->erase(std::unique(v->begin(), v->end()), v->end());
    }
    tp.nX = (uint32_t) tp.tf_uid.size();
    tp.nH = (uint32_t) tp.trg_uid.size();
    if (!tp.trg_uid.empty() && (size_t) tp.trg_uid.back() < 16 * tp.trg_uid.size() + 4096) {
        tp.trg_lut.assign((size_t) tp.trg_uid.back() + 1, UINT32_MAX);     // evidence is re-ranked on every set_evidence
        for (uint32_t r = 0; r < tp.nH; r++) tp.trg_lut[tp.trg_uid[r]] = r;
    }
    const uint32_t nX = tp.nX, nH = tp.nH, E = tp.E;

    // reference ranks and in-degrees
    std::vector<uint32_t> e_tf(E), e_ref(E), indeg_ref(nH, 0);
    tp.n_h_child.assign(nX, 0);
    for (uint32_t e = 0; e < E; e++) {
        e_tf[e] = rank_of(tp.tf_uid, src[e]);
        e_ref[e] = rank_of(tp.trg_uid, trg[e]);
        indeg_ref[e_ref[e]]++;
        tp.n_h_child[e_tf[e]]++;                                        // GraphORNOR.cpp:206-211
    }
    // device target order: descending in-degree, ties in reference order
    tp.ref_of_dt.resize(nH);
    for (uint32_t i = 0; i < nH; i++) tp.ref_of_dt[i] = i;
    if (env_u32("GBN_SORT_TARGETS", 1) != 2)
        std::stable_sort(tp.ref_of_dt.begin(), tp.ref_of_dt.end(),
                         [&](uint32_t a, uint32_t b) { return indeg_ref[a] > indeg_ref[b]; });
    tp.dt_of_ref.resize(nH);
    tp.deg.resize(nH);
    for (uint32_t dt = 0; dt < nH; dt++) {
        tp.dt_of_ref[tp.ref_of_dt[dt]] = dt;
        tp.deg[dt] = indeg_ref[tp.ref_of_dt[dt]];
        tp.max_indeg = std::max(tp.max_indeg, tp.deg[dt]);
    }
    // groups of 32 device targets, padded to the widest member
    tp.n_groups = (nH + 31) / 32;
    tp.gbase.assign(tp.n_groups + 1, 0);
    for (uint32_t g = 0; g < tp.n_groups; g++) {
        uint32_t w = 0;
        for (uint32_t dt = 32 * g; dt < std::min(nH, 32 * g + 32); dt++) w = std::max(w, tp.deg[dt]);
        tp.gbase[g + 1] = tp.gbase[g] + 32 * w;
    }
    tp.Ep = tp.gbase[tp.n_groups];
    const uint32_t Ep = tp.Ep;
    tp.par_tf.assign(Ep, 0); tp.par_edge.assign(Ep, UINT32_MAX); tp.slot_trg.assign(Ep, 0);
    tp.mor_idx.assign(Ep, 1); tp.tfpos_of_slot.assign(Ep, UINT32_MAX); tp.par_pack.assign(Ep, 1u << 30);   // pad: TF 0, mor+1 = 1, never drawn
    tp.slot_of_edge.resize(E);
    for (uint32_t g = 0; g < tp.n_groups; g++)                          // pads also know their lane's target
        for (uint32_t q = tp.gbase[g]; q < tp.gbase[g + 1]; q++)
            tp.slot_trg[q] = std::min(nH - 1, 32 * g + ((q - tp.gbase[g]) & 31));
    {
        std::vector<uint32_t> fill(nH, 0);
        for (uint32_t e = 0; e < E; e++) {                              // edge input order = append_parent order
            const uint32_t dt = tp.dt_of_ref[e_ref[e]];
            const uint32_t q = tp.slot(dt, fill[dt]++);
            tp.par_tf[q] = e_tf[e];
            tp.par_edge[q] = e;
            tp.slot_of_edge[e] = q;
            tp.slot_trg[q] = dt;
            tp.mor_idx[q] = (uint8_t) (mor[e] + 1);
            tp.par_pack[q] = e_tf[e] | ((uint32_t) (mor[e] + 1) << 30);
        }
    }
    // unique children per TF, ascending REFERENCE target order
    tp.tf_ptr.assign(nX + 1, 0);
    {
        std::vector<int64_t> last(nX, -1);
        for (uint32_t r = 0; r < nH; r++) {
            const uint32_t dt = tp.dt_of_ref[r];
            for (uint32_t j = 0; j < tp.deg[dt]; j++) {
                uint32_t k = tp.par_tf[tp.slot(dt, j)];
                if (last[k] != (int64_t) r) { last[k] = r; tp.tf_ptr[k + 1]++; }
            }
        }
        for (uint32_t k = 0; k < nX; k++) {
            tp.max_outdeg = std::max(tp.max_outdeg, tp.tf_ptr[k + 1]);
            tp.tf_ptr[k + 1] += tp.tf_ptr[k];
        }
        tp.Eu = tp.tf_ptr[nX];
        tp.tf_child.assign((size_t) tp.Eu + 1, nH); tp.tf_slot.resize(tp.Eu);     // tf_child[Eu]: sentinel (the dummy target)
        std::vector<uint32_t> pos(nX, 0);
        std::fill(last.begin(), last.end(), -1);
        for (uint32_t r = 0; r < nH; r++) {
            const uint32_t dt = tp.dt_of_ref[r];
            for (uint32_t j = 0; j < tp.deg[dt]; j++) {
                const uint32_t q = tp.slot(dt, j);
                const uint32_t k = tp.par_tf[q];
                if (last[k] != (int64_t) r) {
                    last[k] = r;
                    const uint32_t c = tp.tf_ptr[k] + pos[k]++;
                    tp.tf_child[c] = dt;
                    tp.tf_slot[c] = q;
                    tp.tfpos_of_slot[q] = c;
                }
            }
        }
    }
    // packed views (fast S kernel)
    tp.sbase.assign(tp.n_groups + 1, 0); tp.dbase.assign(tp.n_groups + 1, 0);
    for (uint32_t g = 0; g < tp.n_groups; g++) {
        const uint32_t w = (tp.gbase[g + 1] - tp.gbase[g]) >> 5;
        tp.sbase[g + 1] = tp.sbase[g] + 32 * ((w + 15) / 16);
        tp.dbase[g + 1] = tp.dbase[g] + 32 * ((w + 3) / 4);
        for (uint32_t r = 0; r < (w + 15) / 16; r++) tp.sw_grp.push_back(g);
        for (uint32_t r = 0; r < (w + 3) / 4; r++) tp.dblk_grp.push_back(g);
    }
    tp.nSw = tp.sbase[tp.n_groups]; tp.nD = tp.dbase[tp.n_groups];
    tp.mor2.assign(tp.nSw, 0xffffffffu);
    tp.par16.assign(4 * (size_t) tp.nD, 0);
    tp.tfpos4.assign(4 * (size_t) tp.nD, UINT32_MAX);
    tp.fp_ptr.assign(nX + 1, 0);
    for (uint32_t k = 0; k < nX; k++) tp.fp_ptr[k + 1] = tp.fp_ptr[k] + ((tp.tf_ptr[k + 1] - tp.tf_ptr[k] + 3) & ~3u);
    tp.Efp = tp.fp_ptr[nX];
    tp.fp_child.assign((size_t) tp.Efp + 4, nH); tp.fp_slot.assign(tp.Efp, UINT32_MAX); tp.fp_abit.assign(tp.Efp, UINT32_MAX);
    std::vector<uint32_t> fp_of_slot(tp.Ep, UINT32_MAX);
    for (uint32_t k = 0; k < nX; k++)
        for (uint32_t c = tp.tf_ptr[k]; c < tp.tf_ptr[k + 1]; c++) {
            const uint32_t fpos = tp.fp_ptr[k] + (c - tp.tf_ptr[k]);
            tp.fp_child[fpos] = tp.tf_child[c];
            tp.fp_slot[fpos] = tp.tf_slot[c];
            fp_of_slot[tp.tf_slot[c]] = fpos;
        }
    if (nX <= 65536)
        for (uint32_t dt = 0; dt < nH; dt++) {
            const uint32_t g = dt >> 5, lane = dt & 31;
            for (uint32_t j = 0; j < tp.deg[dt]; j++) {
                const uint32_t q = tp.slot(dt, j);
                uint32_t &mw = tp.mor2[tp.sbase[g] + 32 * (j / 16) + lane];
                mw = (mw & ~(3u << (2 * (j % 16)))) | ((uint32_t) tp.mor_idx[q] << (2 * (j % 16)));
                tp.par16[4 * (size_t) (tp.dbase[g] + 32 * (j / 4) + lane) + (j % 4)] = (uint16_t) tp.par_tf[q];
                tp.tfpos4[4 * (size_t) (tp.dbase[g] + 32 * (j / 4) + lane) + (j % 4)] = fp_of_slot[q];
                if (fp_of_slot[q] != UINT32_MAX) tp.fp_abit[fp_of_slot[q]] = 16 * (tp.sbase[g] + 32 * (j / 16) + lane) + (j % 16);
            }
        }
    tp.x_prior_active.assign(nX, 0);
    for (uint32_t i = 0; i < n_active; i++) {
        uint32_t k = rank_of(tp.tf_uid, active_tf[i]);
        if (k != UINT32_MAX) tp.x_prior_active[k] = 1;
    }
    // T priors (GraphORNOR.cpp:213-223); the ctor's t_alpha / t_beta are overwritten there
    tp.t_a.resize(nX); tp.t_b.resize(nX); tp.t_mean.resize(nX); tp.t_lnB.resize(nX);
    const unsigned int n_zt_groups = 10;
    for (uint32_t k = 0; k < nX; k++) {
        double lognchild = std::log((double) tp.n_h_child[k]);
        double a = t_lmargin + std::max(n_zt_groups - lognchild - 1, 0.) * t_focus;
        double b = t_hmargin + lognchild * t_focus;
        if (!(a > 0.) || !(b > 0.)) throw std::invalid_argument("Beta prior of a T node is not positive");
        tp.t_a[k] = a; tp.t_b[k] = b; tp.t_mean[k] = a / (a + b);
        // gab - ga - gb of gsl_ran_beta_pdf (GSL randist/beta.c), evaluated once on the host
        tp.t_lnB[k] = std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b);
    }
    return tp;
}

// Evidence of one dataset -> observed state per DEVICE target and YPROB (GraphORNOR.cpp:108-134)
inline void compile_evidence(const Topology &tp, uint32_t n_evid, const uint32_t *evid_trg,
                             const int32_t *evid_val, bool comp_yprob,
                             uint8_t *y_out /* [nH], device target order */, double yprob_out[3])
{
    yprob_out[0] = 0.05; yprob_out[1] = 0.90; yprob_out[2] = 0.05;      // GraphORNOR.h:45
    for (uint32_t i = 0; i < n_evid; i++)
        if (evid_val[i] < -1 || evid_val[i] > 1)
            throw std::out_of_range("evidence values can only be either -1, 0 or 1");
    // evidence_dict_t is a std::map filled with insert(): unique keys, first occurrence wins
    // (reference gbnet/ModelORNOR.pyx:64-66).  An open-addressing set of the keys gives the same
    // result without the tree: this runs on the host inside every set_evidence call.
    size_t cap = 16;
    while (cap < 2 * (size_t) n_evid) cap <<= 1;
    std::vector<uint32_t> first(cap, UINT32_MAX);                      // slot -> index of the key's first occurrence
    unsigned int deg_counts[3] = { 0, 0, 0 };
    for (uint32_t i = 0; i < tp.nH; i++) y_out[i] = 1;                 // 1 -> not DEG
    for (uint32_t i = 0; i < n_evid; i++) {
        const uint32_t key = evid_trg[i];
        size_t hh = ((size_t) key * 0x9e3779b1u) & (cap - 1);
        while (first[hh] != UINT32_MAX && evid_trg[first[hh]] != key) hh = (hh + 1) & (cap - 1);
        if (first[hh] != UINT32_MAX) continue;                         // a repeated key: ignored
        first[hh] = i;
        deg_counts[evid_val[i] + 1] += 1;                              // keys outside the network count too
        const uint32_t h = tp.trg_lut.empty() ? rank_of(tp.trg_uid, key) : (key < tp.trg_lut.size() ? tp.trg_lut[key] : UINT32_MAX);
        if (h != UINT32_MAX) y_out[tp.dt_of_ref[h]] = (uint8_t) (evid_val[i] + 1);
    }
    if (comp_yprob) {
        yprob_out[0] = (double) deg_counts[0] / tp.nH;
        yprob_out[2] = (double) deg_counts[2] / tp.nH;
        yprob_out[1] = (double) 1. - yprob_out[0] - yprob_out[2];
    }
}

}  // namespace gbn
