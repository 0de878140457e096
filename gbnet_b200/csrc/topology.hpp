// Host-side topology compiler: integer edge list + evidence -> CSR arrays and priors.
//
// Replaces GraphORNOR::build_structure (reference libgbnet/src/GraphORNOR.cpp:13-265), which
// builds one heap object graph PER CHAIN (ModelORNOR.cpp:23-27).  Here the network is compiled
// once into flat arrays shared by every chain and every dataset:
//
//   by-target CSR  trg_ptr / par_tf / par_edge / mor_idx : parents of each target in edge input
//                  order, i.e. HNodeORNOR::append_parent order (GraphORNOR.cpp:237-263).  A
//                  position in this CSR is called a SLOT; S-node state and counters are stored
//                  in slot order on the device.
//   by-TF CSR      tf_ptr / tf_child / tf_slot : the UNIQUE targets of each TF in ascending
//                  target order (HParentNode::children is a std::set, HParentNode.h:18) with the
//                  slot of the (first) connecting edge.
//   per TF         Beta prior (a, b, mean, ln B) from the log out-degree (GraphORNOR.cpp:213-223)
//   per dataset    observed state y per target, YPROB (GraphORNOR.cpp:113-134)
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace gbn {

constexpr uint32_t TILE_TARGETS = 512;     // fast S kernel: 16 groups of 32 targets per tile
constexpr uint32_t TILE_SLOTS = 8192;

struct Topology {
    uint32_t nX = 0, nH = 0, E = 0, Eu = 0;
    uint32_t max_indeg = 0, max_outdeg = 0;
    std::vector<uint32_t> tf_uid, trg_uid;          // sorted unique ids
    std::vector<uint32_t> trg_ptr, par_tf, par_edge, slot_of_edge, slot_trg;
    std::vector<uint8_t> mor_idx;                   // per slot, mor + 1
    std::vector<uint32_t> tf_ptr, tf_child, tf_slot;
    std::vector<uint32_t> tfpos_of_slot;            // slot -> position in the by-TF CSR (UINT32_MAX for a duplicate edge)
    // fast S phase: targets are cut into TILES of consecutive targets (<= TILE_TARGETS targets and
    // <= tile_slots slots, so one tile's slots are one contiguous, coalescible range); inside a tile
    // threads take targets in descending in-degree order (trg_order) so a warp's lanes finish together
    std::vector<uint32_t> tile_trg0;                // [n_tiles+1] first target of each tile
    std::vector<uint32_t> trg_order;                // [nH] tile-local order
    std::vector<uint32_t> par_pack;                 // [E] slot -> par_tf | (mor+1) << 30
    uint32_t tile_slots = 0;                        // slot capacity of a tile
    std::vector<uint32_t> n_h_child;                // edges per TF (GraphORNOR.cpp:206-211)
    std::vector<uint8_t> x_prior_active;            // per TF
    std::vector<double> t_a, t_b, t_mean, t_lnB;    // per TF
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
    const uint32_t nX = tp.nX, nH = tp.nH, E = tp.E;

    std::vector<uint32_t> e_tf(E), e_trg(E);
    tp.trg_ptr.assign(nH + 1, 0);
    tp.n_h_child.assign(nX, 0);
    for (uint32_t e = 0; e < E; e++) {
        e_tf[e] = rank_of(tp.tf_uid, src[e]);
        e_trg[e] = rank_of(tp.trg_uid, trg[e]);
        tp.trg_ptr[e_trg[e] + 1]++;
        tp.n_h_child[e_tf[e]]++;
    }
    for (uint32_t i = 0; i < nH; i++) {
        tp.max_indeg = std::max(tp.max_indeg, tp.trg_ptr[i + 1]);
        tp.trg_ptr[i + 1] += tp.trg_ptr[i];
    }
    tp.par_tf.resize(E); tp.par_edge.resize(E); tp.slot_of_edge.resize(E);
    tp.slot_trg.resize(E); tp.mor_idx.resize(E);
    {
        std::vector<uint32_t> fill(nH, 0);
        for (uint32_t e = 0; e < E; e++) {
            uint32_t i = e_trg[e];
            uint32_t q = tp.trg_ptr[i] + fill[i]++;
            tp.par_tf[q] = e_tf[e];
            tp.par_edge[q] = e;
            tp.slot_of_edge[e] = q;
            tp.slot_trg[q] = i;
            tp.mor_idx[q] = (uint8_t) (mor[e] + 1);
        }
    }
    // unique children per TF, ascending target order
    tp.tf_ptr.assign(nX + 1, 0);
    {
        std::vector<int64_t> last(nX, -1);
        for (uint32_t i = 0; i < nH; i++)
            for (uint32_t q = tp.trg_ptr[i]; q < tp.trg_ptr[i + 1]; q++) {
                uint32_t k = tp.par_tf[q];
                if (last[k] != (int64_t) i) { last[k] = i; tp.tf_ptr[k + 1]++; }
            }
        for (uint32_t k = 0; k < nX; k++) {
            tp.max_outdeg = std::max(tp.max_outdeg, tp.tf_ptr[k + 1]);
            tp.tf_ptr[k + 1] += tp.tf_ptr[k];
        }
        tp.Eu = tp.tf_ptr[nX];
        tp.tf_child.resize(tp.Eu); tp.tf_slot.resize(tp.Eu);
        tp.tfpos_of_slot.assign(E, UINT32_MAX);
        std::vector<uint32_t> pos(nX, 0);
        std::fill(last.begin(), last.end(), -1);
        for (uint32_t i = 0; i < nH; i++)
            for (uint32_t q = tp.trg_ptr[i]; q < tp.trg_ptr[i + 1]; q++) {
                uint32_t k = tp.par_tf[q];
                if (last[k] != (int64_t) i) {
                    last[k] = i;
                    uint32_t c = tp.tf_ptr[k] + pos[k]++;
                    tp.tf_child[c] = i;
                    tp.tf_slot[c] = q;
                    tp.tfpos_of_slot[q] = c;
                }
            }
    }
    // tiles for the fast S phase
    tp.tile_slots = std::max<uint32_t>(TILE_SLOTS, tp.max_indeg);
    tp.tile_trg0.assign(1, 0);
    {
        uint32_t cnt = 0, slots = 0;
        for (uint32_t i = 0; i < nH; i++) {
            uint32_t d = tp.trg_ptr[i + 1] - tp.trg_ptr[i];
            if (cnt == TILE_TARGETS || slots + d > tp.tile_slots) { tp.tile_trg0.push_back(i); cnt = 0; slots = 0; }
            cnt++; slots += d;
        }
        tp.tile_trg0.push_back(nH);
    }
    tp.trg_order.resize(nH);
    for (uint32_t i = 0; i < nH; i++) tp.trg_order[i] = i;
    for (size_t t = 0; t + 1 < tp.tile_trg0.size(); t++)
        std::stable_sort(tp.trg_order.begin() + tp.tile_trg0[t], tp.trg_order.begin() + tp.tile_trg0[t + 1],
                         [&](uint32_t a, uint32_t b) { return tp.trg_ptr[a + 1] - tp.trg_ptr[a] > tp.trg_ptr[b + 1] - tp.trg_ptr[b]; });
    tp.par_pack.resize(E);
    for (uint32_t q = 0; q < E; q++) tp.par_pack[q] = tp.par_tf[q] | ((uint32_t) tp.mor_idx[q] << 30);
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

// Evidence of one dataset -> observed state per target and YPROB (GraphORNOR.cpp:108-134)
inline void compile_evidence(const Topology &tp, uint32_t n_evid, const uint32_t *evid_trg,
                             const int32_t *evid_val, bool comp_yprob,
                             uint8_t *y_out /* [nH] */, double yprob_out[3])
{
    yprob_out[0] = 0.05; yprob_out[1] = 0.90; yprob_out[2] = 0.05;      // GraphORNOR.h:45
    for (uint32_t i = 0; i < n_evid; i++)
        if (evid_val[i] < -1 || evid_val[i] > 1)
            throw std::out_of_range("evidence values can only be either -1, 0 or 1");
    // evidence_dict_t is a std::map filled with insert(): unique keys, first occurrence wins
    // (reference gbnet/ModelORNOR.pyx:64-66)
    std::map<uint32_t, int32_t> evidence;
    for (uint32_t i = 0; i < n_evid; i++) evidence.insert({ evid_trg[i], evid_val[i] });
    if (comp_yprob) {
        unsigned int deg_counts[3] = { 0, 0, 0 };
        for (auto &kv : evidence) deg_counts[kv.second + 1] += 1;
        yprob_out[0] = (double) deg_counts[0] / tp.nH;
        yprob_out[2] = (double) deg_counts[2] / tp.nH;
        yprob_out[1] = (double) 1. - yprob_out[0] - yprob_out[2];
    }
    for (uint32_t i = 0; i < tp.nH; i++) y_out[i] = 1;                 // 1 -> not DEG
    for (auto &kv : evidence) {
        uint32_t h = rank_of(tp.trg_uid, kv.first);
        if (h != UINT32_MAX) y_out[h] = (uint8_t) (kv.second + 1);
    }
}

}  // namespace gbn
