"""Host-side logic that needs no GPU: the pandas -> integer marshaling of ModelORNOR
(reference gbnet/ModelORNOR.pyx:18-82), the synthetic generator, and the bench's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from gbnet_b200 import synth  # noqa: E402
from gbnet_b200.model import check_sprior, marshal_frames  # noqa: E402


def _frames():
    ents = pd.DataFrame({"uid": ["e1", "e2", "e3", "e4", "e5", "e6"],
                         "name": ["TFb", "TFa", "g2", "g1", "g3", np.nan]})
    rels = pd.DataFrame({"srcuid": ["e1", "e2", "e1", "e2", "e1"], "trguid": ["e3", "e3", "e4", "e4", "e3"],
                         "type": [1, -1, 1, 0, -1]})
    return ents, rels


def test_marshal_orders_and_duplicates():
    ents, rels = _frames()
    m = marshal_frames(ents, rels, {"e3": 1, "e4": -1, "e5": 0}, {"TFa", "nope"})
    # std::set<string> order: byte-wise lexicographic
    assert m["tf_symbols"] == ["TFa", "TFb"] and m["trg_symbols"] == ["g1", "g2"]
    # edge order = first insertion; the duplicated (TFb, g2) keeps its place and takes the LAST sign (dict semantics)
    assert m["edges"] == [("TFb", "g2"), ("TFa", "g2"), ("TFb", "g1"), ("TFa", "g1")]
    assert m["mor"].tolist() == [-1, -1, 1, 0]
    assert m["src"].tolist() == [1, 0, 1, 0] and m["trg"].tolist() == [1, 1, 0, 0]
    # evidence restricted to targets of the network (pyx:52)
    assert dict(zip(m["evid_trg"].tolist(), m["evid_val"].tolist())) == {1: 1, 0: -1}
    assert m["active"] == [0]
    # the reference writes the symbol columns into the caller's frame (pyx:32-33)
    assert "srcsymbol" in rels.columns and "trgsymbol" in rels.columns


def test_marshal_errors():
    ents, rels = _frames()
    with pytest.raises(TypeError):
        marshal_frames(ents, rels, [1, 2, 3])
    rels2 = rels.copy()
    rels2.loc[1, "type"] = 2
    with pytest.raises(IndexError, match="MOR values"):
        marshal_frames(ents, rels2, {"e3": 1})
    with pytest.raises(AssertionError):
        check_sprior(np.ones((2, 3)))
    with pytest.raises(AssertionError):
        check_sprior(np.full((3, 3), 0.5))
    assert len(check_sprior(None)) == 9
    ev = pd.DataFrame({"uid": ["e3", "e4"], "val": [1, np.nan]})
    m = marshal_frames(ents, rels, ev)
    assert m["evid_val"].tolist() == [1]


def test_synthetic_generator_follows_the_reference_law():
    net = synth.gen_network(NX=200, NY=3000, AvgNTF=6.0, num_active_tfs=5, seed=3)
    indeg = np.bincount(net.trg - net.NX, minlength=net.NY)
    assert abs(indeg.mean() - 6.0) < 0.3                       # Gamma(3, AvgNTF/3) rounded
    outdeg = np.bincount(net.src, minlength=net.NX)
    assert outdeg[:20].mean() > 3 * outdeg[-20:].mean()        # hub-skewed TF choice (tools.py:31-35)
    assert abs((net.mor == -1).mean() - 0.35) < 0.03           # tools.py:37
    assert set(np.unique(net.evid_val)) <= {-1, 0, 1} and net.evid_trg.size == net.NY
    pairs = set(zip(net.src.tolist(), net.trg.tolist()))
    assert len(pairs) == net.n_edges                           # TFs drawn without replacement per target
    a = synth.gen_network(NX=30, NY=100, AvgNTF=2.0, seed=9)
    b = synth.gen_network(NX=30, NY=100, AvgNTF=2.0, seed=9)
    assert np.array_equal(a.src, b.src) and np.array_equal(a.evid_val, b.evid_val)
    r = synth.shape("C1").realised()
    assert r["n_tf"] <= 50 and 1900 <= r["n_trg"] <= 2000 and 4000 <= r["n_edge"] <= 6000


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference: the CPU sampler on the host cores, same JSON contract."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--shape", "C1",
                          "--chains", "3", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "edge-updates/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_get_tests_matches_fisher_exact():
    """gbnet_b200.utils.get_tests against scipy.stats.fisher_exact, which the reference calls per TF
    (gbnet/utils.py:36-52)."""
    from scipy.stats import fisher_exact
    from gbnet_b200 import utils
    net = synth.gen_network(NX=25, NY=400, AvgNTF=3.0, num_active_tfs=3, seed=12)
    ents, rels, evid, tests = utils.frames_from_network(net)
    assert tests.gt_act.sum() == 3 and len(tests) == len(np.unique(net.src))
    ev = evid.set_index("uid").val
    all_ydeg, all_ndeg = int((ev != 0).sum()), int((ev == 0).sum())
    for uid, row in tests.iterrows():
        trg = rels.loc[rels.srcuid == uid]
        val = ev.loc[trg.trguid].values
        assert row.trg_ydeg == (val != 0).sum() and row.trg_ndeg == (val == 0).sum()
        p = fisher_exact([[row.trg_ydeg, all_ydeg - row.trg_ydeg], [row.trg_ndeg, all_ndeg - row.trg_ndeg]],
                         alternative="greater")[1]
        assert abs(p - row.enrichment) < 1e-12
        p = fisher_exact([[row.morp_degp, row.morn_degp], [row.morp_degn, row.morn_degn]], alternative="greater")[1]
        assert abs(p - row.mor_binary) < 1e-12
        assert row.trg_tot == len(trg)


def test_get_tests_empty_tables_give_p_one():
    """A TF whose 2x2 table is all zeros (no DE target behind a signed edge, or only edges of type 0) gets
    p = 1 from fisher_exact (reference gbnet/utils.py:36-52), not NaN."""
    import pandas as pd
    from scipy.stats import fisher_exact
    from gbnet_b200 import utils
    # TF 0: signed edges to DE targets; TF 1: all of its targets are non-DE; TF 2: edges of type 0 only
    rels = pd.DataFrame({"srcuid": [0, 0, 0, 1, 1, 2, 2], "trguid": [10, 11, 12, 12, 13, 10, 11],
                         "type": [1, -1, 1, 1, -1, 0, 0]})
    evid = pd.DataFrame({"uid": [10, 11, 12, 13], "val": [1, -1, 0, 0]})
    tests = utils.get_tests(evid, rels)
    assert not tests[["enrichment", "mor_binary", "combined"]].isna().any().any()
    for uid in (1, 2):
        row = tests.loc[uid]
        assert (row.morp_degp, row.morn_degp, row.morp_degn, row.morn_degn) == (0, 0, 0, 0)
        assert row.mor_binary == 1.0 == fisher_exact([[0, 0], [0, 0]], alternative="greater")[1]
    rho = tests.loc[1].enrichment * 1.0
    assert abs(tests.loc[1].combined - rho * (1 - np.log(rho))) < 1e-15


# ---- the C++ topology / evidence compiler behind gbn_model_create, through its host-only probe ----------

def _probe(src, trg, mor, evid_trg=None, evid_val=None, comp_yprob=False):
    import ctypes as C

    from gbnet_b200 import _capi
    lib = _capi.load()
    src = np.ascontiguousarray(src, np.uint32); trg = np.ascontiguousarray(trg, np.uint32)
    mor = np.ascontiguousarray(mor, np.int32)
    u32p, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    net = _capi.Network(len(src), src.ctypes.data_as(u32p), trg.ctypes.data_as(u32p), mor.ctypes.data_as(i32p), 0, None)
    ev = None
    if evid_trg is not None:
        et = np.ascontiguousarray(evid_trg, np.uint32); evv = np.ascontiguousarray(evid_val, np.int32)
        ev = _capi.Evidence(1, len(et), et.ctypes.data_as(u32p), evv.ctypes.data_as(i32p))
    d = _capi.TopologyDims()
    evp = C.byref(ev) if ev is not None else None
    _capi.check(lib.gbn_topology_probe(C.byref(net), evp, int(comp_yprob), C.byref(d), *([None] * 9)))
    out = {k: np.zeros(n, np.uint32) for k, n in (("slot_of_edge", d.n_edge), ("deg", d.n_trg), ("ref_of_dt", d.n_trg),
                                                  ("gbase", d.n_groups + 1), ("tf_ptr", d.n_tf + 1),
                                                  ("tf_child", d.n_edge_unique), ("tf_slot", d.n_edge_unique))}
    y = np.zeros(d.n_trg, np.uint8)
    yprob = np.zeros(3)
    _capi.check(lib.gbn_topology_probe(
        C.byref(net), evp, int(comp_yprob), C.byref(d),
        *[out[k].ctypes.data_as(u32p) for k in ("slot_of_edge", "deg", "ref_of_dt", "gbase", "tf_ptr", "tf_child", "tf_slot")],
        y.ctypes.data_as(C.POINTER(C.c_uint8)), yprob.ctypes.data_as(C.POINTER(C.c_double))))
    out.update(dims=d, y=y, yprob=yprob)
    return out


def test_topology_compiler_layout_invariants():
    """Device layout of the network (csrc/topology.hpp): targets sorted by in-degree, 32-target groups stored
    transposed and padded to the group's largest in-degree, the by-TF list in ascending reference order."""
    rng = np.random.default_rng(5)
    n_tf, n_trg = 37, 451
    ids_tf = rng.choice(10_000, n_tf, replace=False) + 1_000_000        # sparse, unordered ids
    ids_trg = rng.choice(10_000, n_trg, replace=False)
    pairs = {(int(rng.integers(n_tf)), int(rng.integers(n_trg))) for _ in range(2500)}
    pairs |= {(0, t) for t in range(n_trg)}                               # every target has a parent; TF 0 is a hub
    pairs = sorted(pairs, key=lambda _: rng.random())
    src = np.array([ids_tf[a] for a, _ in pairs]); trg = np.array([ids_trg[b] for _, b in pairs])
    mor = rng.integers(-1, 2, len(pairs))
    o = _probe(src, trg, mor)
    d = o["dims"]
    tf_sorted, trg_sorted = np.unique(src), np.unique(trg)
    assert (d.n_tf, d.n_trg, d.n_edge, d.n_edge_unique) == (len(tf_sorted), n_trg, len(pairs), len(pairs))
    indeg_ref = np.bincount(np.searchsorted(trg_sorted, trg), minlength=n_trg)
    # device targets: a permutation of the reference ranks, in-degree non-increasing
    assert sorted(o["ref_of_dt"].tolist()) == list(range(n_trg))
    assert np.array_equal(o["deg"], indeg_ref[o["ref_of_dt"]]) and np.all(np.diff(o["deg"].astype(int)) <= 0)
    assert d.max_indeg == indeg_ref.max() and d.n_groups == (n_trg + 31) // 32
    # groups: 32 targets x (largest in-degree of the group) slots
    widths = np.diff(o["gbase"].astype(np.int64)) // 32
    for g in range(d.n_groups):
        assert widths[g] == o["deg"][32 * g: 32 * g + 32].max()
    assert o["gbase"][0] == 0 and o["gbase"][-1] == d.n_slots
    # every edge has its own slot; slot -> (device target, parent position) is consistent with the edge
    slots = o["slot_of_edge"].astype(np.int64)
    assert len(set(slots.tolist())) == len(pairs) and slots.max() < d.n_slots
    dt_of_ref = np.argsort(o["ref_of_dt"])
    seen_j = {}
    for e, q in enumerate(slots):
        g = np.searchsorted(o["gbase"], q, side="right") - 1
        lane, j = (q - o["gbase"][g]) % 32, (q - o["gbase"][g]) // 32
        dt = 32 * g + lane
        assert dt == dt_of_ref[np.searchsorted(trg_sorted, trg[e])] and j < o["deg"][dt]
        assert seen_j.setdefault((dt, j), e) == e
        # parents of a target keep the input order of its edges (append_parent order, GraphORNOR.cpp:226-244)
    for dt in range(n_trg):
        es = [e for e in range(len(pairs)) if dt_of_ref[np.searchsorted(trg_sorted, trg[e])] == dt]
        js = [int((slots[e] - o["gbase"][dt // 32]) // 32) for e in es]
        assert js == list(range(len(es)))
    # by-TF list: the TF's targets (device ids) in ascending reference order, with the slots of those edges
    assert o["tf_ptr"][0] == 0 and o["tf_ptr"][-1] == len(pairs)
    for k, uid in enumerate(tf_sorted):
        es = sorted(np.nonzero(src == uid)[0], key=lambda e: np.searchsorted(trg_sorted, trg[e]))
        lo, hi = o["tf_ptr"][k], o["tf_ptr"][k + 1]
        assert o["tf_child"][lo:hi].tolist() == [int(dt_of_ref[np.searchsorted(trg_sorted, trg[e])]) for e in es]
        assert o["tf_slot"][lo:hi].tolist() == [int(slots[e]) for e in es]
    assert d.max_outdeg == np.diff(o["tf_ptr"].astype(int)).max()


@pytest.mark.parametrize("sparse", [False, True])
def test_evidence_compiler_first_key_wins_and_yprob_counts(sparse):
    """evidence_dict_t is a std::map filled with insert(): the first value of a repeated key stays
    (ModelORNOR.pyx:64-66); keys outside the network are ignored for y but counted by comp_yprob
    (GraphORNOR.cpp:113-125).  Dense ids use a lookup table, sparse ids a binary search."""
    scale = 3_000_001 if sparse else 1
    trg_ids = np.array([5, 9, 2, 7, 11]) * scale
    src = np.array([100, 100, 101, 101, 102]); mor = np.array([1, -1, 0, 1, -1])
    evid_trg = np.array([9, 2, 9, 400, 11, 2, 401]) * scale
    evid_val = np.array([1, -1, -1, 1, 0, 1, -1])
    o = _probe(src, trg_ids, mor, evid_trg, evid_val, comp_yprob=True)
    dt_of_ref = np.argsort(o["ref_of_dt"])
    want = {2: -1, 9: 1, 11: 0}                                           # first occurrences of keys in the network
    for r, uid in enumerate(np.unique(trg_ids)):
        assert o["y"][dt_of_ref[r]] == want.get(int(uid // scale), 0) + 1
    # unique keys: 9 -> +1, 2 -> -1, 400 -> +1, 11 -> 0, 401 -> -1 over 5 targets
    assert np.allclose(o["yprob"], [2 / 5, 1 - 2 / 5 - 2 / 5, 2 / 5])
    o2 = _probe(src, trg_ids, mor, evid_trg, evid_val, comp_yprob=False)
    assert np.allclose(o2["yprob"], [0.05, 0.90, 0.05]) and np.array_equal(o2["y"], o["y"])
    from gbnet_b200 import GbnError
    with pytest.raises(GbnError):
        _probe(src, trg_ids, mor, evid_trg, np.array([1, -1, -1, 2, 0, 1, -1]))
    with pytest.raises(GbnError):
        _probe(src, trg_ids, np.array([1, -1, 0, 5, -1]))
