"""Synthetic TF->target networks and DE evidence following the reference generator's law.

Mirrors ``gbnet.tools.genData`` (reference gbnet/tools.py:13-87) as integer arrays, vectorised so
that STRING-scale shapes (1,500 TFs / 20,000 targets / 300,000 edges) are generated in seconds:

* per-target in-degree ``min(round(Gamma(shape=3, scale=AvgNTF/3)), NX)``        (tools.py:27-28)
* TFs drawn without replacement with weights ``gamma.pdf(i/NX, a=1, scale=0.2)``   (tools.py:31-35);
  successive weighted sampling is realised with the Efraimidis-Spirakis / Gumbel-top-k
  construction, which has exactly the same distribution
* edge sign -1 / +1 with probability 0.35 / 0.65                                    (tools.py:37)
* ``num_active_tfs`` planted TFs; target state = sign(sum sign*X) pushed through the noise rows
  ``[1-a-b,a,b],[a,1-2a,a],[b,a,1-a-b]``                                            (tools.py:39-61)
* targets that drew zero edges do not appear in the network                          (tools.py:64-68)

uids follow the reference: TFs are ``0..NX-1``, targets ``NX..NX+NY-1`` (tools.py:18-19,71-74).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SynthNetwork:
    src: np.ndarray        # [E] uint32 TF uid, edge input order (target-major, draw order inside)
    trg: np.ndarray        # [E] uint32 target uid
    mor: np.ndarray        # [E] int32 -1/+1
    evid_trg: np.ndarray   # [NY] uint32 every mRNA uid (the reference's evid frame lists all of them)
    evid_val: np.ndarray   # [NY] int32 -1/0/+1
    gt_active: np.ndarray  # planted TF uids
    NX: int
    NY: int

    @property
    def n_edges(self) -> int:
        return int(self.src.size)

    def realised(self):
        return dict(n_tf=int(np.unique(self.src).size), n_trg=int(np.unique(self.trg).size),
                    n_edge=int(self.src.size))


def gen_network(NX=3, NY=50, AvgNTF=0.5, num_active_tfs=2, active_tfs=None, a=0.05, b=0.005,
                seed=0, shuffle_edges=False) -> SynthNetwork:
    rng = np.random.default_rng(seed)
    indeg = np.minimum(np.rint(rng.gamma(3.0, AvgNTF / 3.0, size=NY)).astype(np.int64), NX)
    x = np.arange(NX) / NX
    logw = -x / 0.2                       # log gamma.pdf(x, a=1, scale=0.2) up to a constant
    src_parts, trg_parts = [], []
    chunk = max(1, int(4_000_000 // max(NX, 1)))
    for lo in range(0, NY, chunk):
        hi = min(NY, lo + chunk)
        keys = logw[None, :] + rng.gumbel(size=(hi - lo, NX))
        order = np.argsort(-keys, axis=1)                       # successive-draw order
        k = indeg[lo:hi]
        mask = np.arange(NX)[None, :] < k[:, None]
        rows = np.nonzero(mask)
        src_parts.append(order[rows[0], rows[1]])
        trg_parts.append(rows[0] + lo)
    src = np.concatenate(src_parts).astype(np.uint32)
    trg_local = np.concatenate(trg_parts)
    E = src.size
    mor = np.where(rng.random(E) < 0.35, -1, 1).astype(np.int32)

    if active_tfs is None or len(active_tfs) == 0:
        active = rng.choice(NX, size=min(num_active_tfs, NX), replace=False)
    else:
        active = np.asarray(active_tfs)
    xgt = np.zeros(NX, np.int64)
    xgt[active] = 1
    acc = np.zeros(NY, np.int64)
    np.add.at(acc, trg_local, mor * xgt[src])
    sgn = np.sign(acc)
    p = np.array([[1. - a - b, a, b], [a, 1. - a - a, a], [b, a, 1. - a - b]])
    cum = np.cumsum(p[sgn + 1], axis=1)
    u = rng.random(NY)
    yval = (u[:, None] >= cum).sum(axis=1).clip(0, 2) - 1

    if shuffle_edges:
        perm = rng.permutation(E)
        src, trg_local, mor = src[perm], trg_local[perm], mor[perm]

    return SynthNetwork(src=src, trg=(trg_local + NX).astype(np.uint32), mor=mor,
                        evid_trg=(np.arange(NY) + NX).astype(np.uint32),
                        evid_val=yval.astype(np.int32),
                        gt_active=np.sort(active).astype(np.uint32), NX=NX, NY=NY)


def redraw_evidence(net: SynthNetwork, num_active_tfs=2, a=0.05, b=0.005, seed=0):
    """A new contrast on the same network: re-plant active TFs and re-draw the noisy DE calls
    (the C4 'batched contrasts' workload of SURVEY.md §8(d))."""
    rng = np.random.default_rng(seed)
    NX, NY = net.NX, net.NY
    active = rng.choice(NX, size=min(num_active_tfs, NX), replace=False)
    xgt = np.zeros(NX, np.int64)
    xgt[active] = 1
    acc = np.zeros(NY, np.int64)
    np.add.at(acc, net.trg.astype(np.int64) - NX, net.mor * xgt[net.src])
    sgn = np.sign(acc)
    p = np.array([[1. - a - b, a, b], [a, 1. - a - a, a], [b, a, 1. - a - b]])
    cum = np.cumsum(p[sgn + 1], axis=1)
    u = rng.random(NY)
    yval = (u[:, None] >= cum).sum(axis=1).clip(0, 2) - 1
    return yval.astype(np.int32), np.sort(active).astype(np.uint32)


# the BASELINE.json shapes; hyper-parameters are the CLI defaults
# (reference gbnet/commands/ornor_inference.py:15-28 through gbnet/utils.py:70-87)
SHAPES = {
    "C1": dict(NX=50, NY=2000, AvgNTF=2.5, num_active_tfs=3, seed=1, chains=3),
    "C2": dict(NX=800, NY=2500, AvgNTF=3.6, num_active_tfs=8, seed=2, chains=64),
    "C3": dict(NX=1500, NY=20000, AvgNTF=15.0, num_active_tfs=12, seed=3, chains=256),
}


def cli_default_hyper(s_leniency=0.01):
    l = s_leniency
    return dict(
        sprior=(1.0 - l, 0.9 * l, 0.1 * l, 0.5 * l, 1.0 - l, 0.5 * l, 0.1 * l, 0.9 * l, 1.0 - l),
        z_alpha=2000., z_beta=1., z0_alpha=200., z0_beta=1., t_alpha=2., t_beta=2.,
        t_focus=1., t_lmargin=2., t_hmargin=2., zn_focus=2., zn_lmargin=8., zn_hmargin=2.,
        noise_listen_children=False, comp_yprob=False, const_params=False)


def shape(name: str, **over) -> SynthNetwork:
    cfg = dict(SHAPES[name])
    cfg.pop("chains")
    cfg.update(over)
    return gen_network(**cfg)
