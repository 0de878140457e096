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
