"""``ModelORNOR`` — the drop-in surface (reference gbnet/ModelORNOR.pyx:10-130, exported as
``gbnet.ModelORNOR`` by gbnet/__init__.py:1).

Same constructor keywords, same methods, same return shapes as the Cython class; behind it the
network is integer-coded once (symbols ranked in the reference's std::set<string> order, i.e.
byte-wise lexicographic: GraphORNOR.cpp:31-42) and handed to the CUDA library through the C ABI.
Extra keywords (``seed``, ``device``, ``kernel``) default to the reference behaviour, except that
chains are reproducible: the reference seeds each chain from time(NULL) (ModelORNOR.cpp:24).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .engine import KERNEL_AUTO, Sampler


def _rank(symbols):
    """symbol -> rank in byte-wise lexicographic order (std::set<std::string>)"""
    order = sorted(set(symbols), key=lambda s: str(s).encode("utf8"))   # the reference requires str symbols (pyx:58)
    return {s: i for i, s in enumerate(order)}, order


def marshal_frames(ents, rels, evid=None, active_tf_set_py=frozenset()):
    """pandas -> integer problem: what ModelORNOR.pyx:18-66 does before calling the C++ ctor.

    Returns a dict with the integer edge list (edge order = insertion order of the reference's
    ``network_py`` dict, duplicates collapsed with the last sign winning), the evidence restricted to
    targets of the network, the active-TF ids, and the symbol tables in std::set<string> order."""
    assertion_msg = ("Please make sure that provided ents file is a "
                     "Pandas Dataframe with column 'uid' available")
    ents = ents[["uid", "name"]].dropna()                         # pyx:25-28
    if 'uid' in ents.columns:
        ents = ents.set_index('uid')
    assert ents.index.name == 'uid', assertion_msg

    # pyx:32-36 (the reference writes the symbol columns into the caller's frame as well)
    rels["srcsymbol"] = ents.loc[rels.srcuid, "name"].values
    rels["trgsymbol"] = ents.loc[rels.trguid, "name"].values
    trg_set = set(rels["trgsymbol"].values)
    network_py = {(src, trg): mor for src, trg, mor in rels[["srcsymbol", "trgsymbol", "type"]].values}

    assertion_msg = ("Please make sure that provided DEG file is either a "
                     "dictionary or a Pandas Series with uids as keys/index")
    if evid is not None:                                          # pyx:41-54
        if isinstance(evid, pd.DataFrame):
            assert 'uid' in evid.columns, assertion_msg
            evid = evid[["uid", "val"]].dropna()
        elif isinstance(evid, dict):
            evid = pd.DataFrame([dict(uid=k, val=v) for k, v in evid.items()]).dropna()
        else:
            raise TypeError(assertion_msg)
        evid = evid.copy()
        evid["symbol"] = ents.loc[evid.uid, "name"].values
        evid = evid[evid["symbol"].isin(trg_set)]
        evidence_py = dict(zip(evid["symbol"].values, evid["val"].values))
    else:
        evidence_py = {}

    for mor in network_py.values():                               # GraphORNOR.cpp:39-41
        if mor not in (-1, 0, 1):
            raise IndexError("MOR values can only be either -1, 0 or 1")

    src_rank, tf_symbols = _rank(s for s, _ in network_py)
    trg_rank, trg_symbols = _rank(t for _, t in network_py)
    edges = list(network_py.keys())
    ev_items = [(trg_rank[s], int(v)) for s, v in evidence_py.items() if s in trg_rank]
    return dict(
        src=np.fromiter((src_rank[s] for s, _ in edges), np.uint32, len(edges)),
        trg=np.fromiter((trg_rank[t] for _, t in edges), np.uint32, len(edges)),
        mor=np.fromiter((int(m) for m in network_py.values()), np.int32, len(edges)),
        evid_trg=np.array([k for k, _ in ev_items], np.uint32),
        evid_val=np.array([v for _, v in ev_items], np.int32),
        active=sorted(src_rank[s] for s in active_tf_set_py if s in src_rank),
        tf_symbols=tf_symbols, trg_symbols=trg_symbols, edges=edges)


def marshal_contrasts(ents, rels, evids, active_tf_set_py=frozenset()):
    """``marshal_frames`` once for a network and MANY evidence frames (batched contrasts, BASELINE config C4):
    the symbol tables and the integer edge list are built a single time; every contrast contributes one row of
    the evidence matrix over the union of its targets (0 = no evidence: the reference's default state of a
    target without evidence is "not DE", GraphORNOR.cpp:131)."""
    if len(evids) == 0:
        raise ValueError("from_contrasts needs at least one evidence frame")
    base = marshal_frames(ents, rels, evids[0], active_tf_set_py)
    trg_rank = {s: i for i, s in enumerate(base["trg_symbols"])}
    names = ents[["uid", "name"]].dropna().set_index("uid")["name"]
    rows = [dict(zip(base["evid_trg"].tolist(), base["evid_val"].tolist()))]
    for evid in evids[1:]:
        if isinstance(evid, dict):
            evid = pd.DataFrame([dict(uid=k, val=v) for k, v in evid.items()])
        elif not isinstance(evid, pd.DataFrame):
            raise TypeError("Please make sure that provided DEG file is either a dictionary or a Pandas Series with uids as keys/index")
        evid = evid[["uid", "val"]].dropna()
        sym = names.loc[evid.uid].values
        row = {}
        for s, v in zip(sym, evid["val"].values):                # dict semantics of pyx:54: the last value of a symbol wins
            if s in trg_rank:
                row[trg_rank[s]] = int(v)
        rows.append(row)
    keys = sorted(set().union(*[r.keys() for r in rows]))
    val = np.zeros((len(rows), len(keys)), np.int32)
    col = {k: i for i, k in enumerate(keys)}
    for d, r in enumerate(rows):
        for k, v in r.items():
            val[d, col[k]] = v
    base["evid_trg"] = np.array(keys, np.uint32)
    base["evid_val"] = val
    return base


def check_sprior(sprior):
    """pyx:70-82"""
    if sprior is None:
        return [0.40, 0.40, 0.20,
                0.30, 0.40, 0.30,
                0.20, 0.40, 0.40]
    sprior = np.array(sprior)
    assert sprior.shape == (3, 3), 'wrong shape for sprior'
    assert (sprior.sum(axis=1) == 1).all(), 'bad sprior, rowsums != 1'
    return sprior.flatten().tolist()


class ModelORNOR:
    def __init__(self, ents, rels, evid=None, active_tf_set_py=set(),
                 sprior=None, z_alpha=25., z_beta=25., z0_alpha=None, z0_beta=None, t_alpha=2., t_beta=2.,
                 n_graphs=3, noise_listen_children=True, comp_yprob=False, const_params=False, x_mode=1,
                 t_focus=2., t_lmargin=2., t_hmargin=2., zn_focus=2., zn_lmargin=8., zn_hmargin=2.,
                 seed=0, device=0, kernel=KERNEL_AUTO):
        z0_alpha = z_alpha if z0_alpha is None else z0_alpha          # pyx:18-19
        z0_beta = z_beta if z0_beta is None else z0_beta
        # evid may be a LIST of evidence frames (see from_contrasts): one network, many contrasts
        m = marshal_contrasts(ents, rels, evid, active_tf_set_py) if isinstance(evid, (list, tuple)) \
            else marshal_frames(ents, rels, evid, active_tf_set_py)
        sprior_py = check_sprior(sprior)
        self._tf_symbols, self._trg_symbols, self._edges = m["tf_symbols"], m["trg_symbols"], m["edges"]
        src, trg, mor, evid_trg, evid_val, active = m["src"], m["trg"], m["mor"], m["evid_trg"], m["evid_val"], m["active"]

        self._sampler = Sampler(src, trg, mor, evid_trg, evid_val, active_tf=active, sprior=sprior_py,
                                z_alpha=z_alpha, z_beta=z_beta, z0_alpha=z0_alpha, z0_beta=z0_beta,
                                t_alpha=t_alpha, t_beta=t_beta, n_graphs=n_graphs,
                                noise_listen_children=noise_listen_children, comp_yprob=comp_yprob,
                                const_params=const_params, t_focus=t_focus, t_lmargin=t_lmargin,
                                t_hmargin=t_hmargin, zn_focus=zn_focus, zn_lmargin=zn_lmargin,
                                zn_hmargin=zn_hmargin, seed=seed, device=device, kernel=kernel)
        self._const_params = bool(const_params)

    @classmethod
    def from_contrasts(cls, ents, rels, evids, **options):
        """One network, many DE contrasts (BASELINE config C4): the frames are marshalled ONCE (the reference
        rebuilds its string maps per model and per chain, ModelORNOR.pyx:13-83, ModelORNOR.cpp:23-27) and every
        contrast gets ``n_graphs`` chains of the same device-resident network.  The getters take ``dataset=``;
        ``get_max_gelman_rubin`` is the maximum over the contrasts."""
        return cls(ents, rels, list(evids), **options)

    @property
    def n_contrasts(self):
        return self._sampler.n_datasets()

    def run_protocol(self, burn_its=10, max_its=100, N=20, dN=30):
        """gbnet/utils.py:93-135 as one library call (gbn_model_run_protocol): (calls made, last max R-hat)"""
        return self._sampler.run_protocol(burn_its=burn_its, max_its=max_its, N=N, dN=dN)

    # ids exactly as RVNode builds them (RVNode.cpp:29-33; GraphORNOR.cpp:243): name + "_" + uid
    def _node_ids(self):
        if self._sampler.simulation():          # evid=None: H and Y per target, then T (GraphORNOR.cpp:89-106)
            ids = [[k, t] for t in self._trg_symbols for k in ("H", "Y")]
            if not self._const_params:
                ids += [["T", s] for s in self._tf_symbols]
            return ids
        ids = [["X", s] for s in self._tf_symbols]
        if not self._const_params:
            ids += [["T", s] for s in self._tf_symbols]
        ids += [["S", f"{s}-->{t}"] for s, t in self._edges]
        return ids

    def get_gelman_rubin(self, dataset=0):
        gr = self._sampler.gelman_rubin(dataset)
        # the reference splits "<name>_<uid>" on "_" and keeps two fields (pyx:85-87)
        return [f"{name}_{uid}".split('_')[:2] + [float(v)] for (name, uid), v in zip(self._node_ids(), gr)]

    def get_max_gelman_rubin(self):
        return self._sampler.max_gelman_rubin()

    def sample(self, N=50000, deltaN=30):
        self._sampler.sample(N, deltaN)
        return

    def sample_n(self, n):
        self._sampler.sample_n(n)

    def print_stats(self):
        for row in self.get_posterior_means("X") + self.get_posterior_means("T") + self.get_posterior_means("S"):
            print("_".join(row[:-1]), f"{row[-1]:.4f}")

    def burn_stats(self):
        self._sampler.burn_stats()

    def _posterior(self, values, var_name):
        per = {"X": 2, "T": 1, "S": 3, "H": 3, "Y": 3}.get(var_name)
        if per is None:
            return []
        ids = [i for i in self._node_ids() if i[0] == var_name]
        out = []
        for n, (name, uid) in enumerate(ids):
            for j in range(per):
                out.append(f"{name}_{uid}_{j}".split('_') + [float(values[n * per + j])])   # pyx:120-126
        return out

    def get_posterior_means(self, var_name, dataset=0):
        return self._posterior(self._sampler.posterior_means(var_name, dataset), var_name)

    def get_posterior_sdevs(self, var_name, dataset=0):
        return self._posterior(self._sampler.posterior_sdevs(var_name, dataset), var_name)

    @property
    def sampler(self) -> Sampler:
        return self._sampler
