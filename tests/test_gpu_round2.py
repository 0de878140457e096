"""GPU tests of the round-2 paths: the persistent shared-memory kernel for networks that fit one SM, the
packed S views, the protocol / batched-contrast / pre-screen entry points of the C ABI, the compact moment
exchange over NCCL (2 ranks), the full BASELINE shapes C4 and C5, and the T = 1.0 guard."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from conftest import make_problem, rel_err, sampler_from_problem  # noqa: E402

pytestmark = pytest.mark.gpu


def _states(G, chains):
    return [G.get_state(c) for c in chains]


# ------------------------------------------------------------------------------------------- sweep paths
def test_path_selection(monkeypatch):
    """Networks that fit one SM with non-listening T nodes run on the persistent kernel (path 2); listening T
    nodes, GBN_SMALL=2 and large networks stay on the multi-kernel path (1); kernel=STRICT is path 0."""
    import gbnet_b200
    monkeypatch.delenv("GBN_SMALL", raising=False)
    pb, _ = make_problem(NX=20, NY=200, seed=7, n_graphs=2)
    assert sampler_from_problem(pb).path() == 2
    assert sampler_from_problem(pb, noise_listen_children=True).path() == 1
    assert sampler_from_problem(pb, noise_listen_children=True, const_params=True).path() == 2
    assert sampler_from_problem(pb, kernel=gbnet_b200.KERNEL_STRICT).path() == 0
    monkeypatch.setenv("GBN_SMALL", "2")
    assert sampler_from_problem(pb).path() == 1


@pytest.mark.parametrize("listen", [False, True])
@pytest.mark.parametrize("xq", ["1", "2"])
def test_pair_x_kernel_with_the_second_s_kernel(monkeypatch, listen, xq):
    """Networks whose flip weights do not fit an SM sweep X on the chain-pair kernel (k_fast_x, GBN_X1=2) and S on
    k_fast_s2, which then writes the bundle-interleaved bound bytes instead of the per-chain rows (the parity suites'
    "pair" mode runs k_fast_x with k_fast_s): the oracle's chain, with the bounds in shared memory (GBN_X_Q=1) and
    read from global memory (GBN_X_Q=2), listening and non-listening T nodes, keyed and injected streams."""
    from test_gpu_parity import _run_trajectory
    import gbnet_b200
    monkeypatch.setenv("GBN_SMALL", "2")
    monkeypatch.setenv("GBN_X1", "2")
    monkeypatch.setenv("GBN_X_Q", xq)
    monkeypatch.delenv("GBN_S_V", raising=False)
    pb, _ = make_problem(NX=25, NY=300, AvgNTF=4.0, seed=31, n_graphs=5, noise_listen_children=listen)
    for injected in (False, True):
        _run_trajectory(pb, gbnet_b200.KERNEL_FAST, 30, injected, chunks=(4, 9))


@pytest.mark.parametrize("lpt", [0, 4, 8, 16, 32])
def test_persistent_kernel_is_the_same_chain(monkeypatch, lpt):
    """The persistent shared-memory kernel, the multi-kernel fast path and the oracle draw the same chains:
    discrete states bit for bit, T to 1e-9, counters / moments / R-hat identical - in one batch or split into
    several, for every lanes-per-TF layout of its X phase."""
    monkeypatch.delenv("GBN_SMALL", raising=False)
    if lpt:
        monkeypatch.setenv("GBN_SMALL_LPT", str(lpt))
    pb, _ = make_problem(NX=40, NY=900, AvgNTF=3.0, seed=17, n_graphs=3, num_active_tfs=4)
    A = sampler_from_problem(pb)
    assert A.path() == 2
    monkeypatch.setenv("GBN_SMALL", "2")
    B = sampler_from_problem(pb)
    assert B.path() == 1
    P = oracle.PortModel(pb)
    A.sample_n(7); A.sample_n(30); A.sample_n(23)
    B.sample_n(60)
    P.sample_n(60)
    nX = A.n_tf()
    for c in range(3):
        a, b, p = A.get_state(c), B.get_state(c), P.get_state(c)
        assert np.array_equal(a[:nX], p[:nX]) and np.array_equal(a[2 * nX:], p[2 * nX:])
        assert np.array_equal(a[:nX], b[:nX]) and np.array_equal(a[2 * nX:], b[2 * nX:])
        assert np.abs(a[nX:2 * nX] - p[nX:2 * nX]).max() < 1e-9
        ma, va = A.node_moments(c)
        mb, vb = B.node_moments(c)
        disc = np.r_[0:2 * nX, 3 * nX:A.n_stats()]
        assert np.array_equal(ma[disc], mb[disc]) and np.array_equal(va[disc], vb[disc])
    assert rel_err(A.gelman_rubin(), P.gelman_rubin()).max() < 1e-9
    assert rel_err(A.posterior_means("X"), P.posterior_means("X")).max() < 1e-9
    assert rel_err(A.posterior_sdevs("S"), B.posterior_sdevs("S")).max() < 1e-9


def test_persistent_kernel_with_injected_draws():
    """Bit-exact discrete updates under an injected draw stream on the persistent kernel (the parity
    definition of BASELINE.json), including a run that consumes the injected rows across two batches."""
    pb, _ = make_problem(NX=25, NY=300, AvgNTF=2.5, seed=23, n_graphs=2)
    G = sampler_from_problem(pb)
    assert G.path() == (1 if os.environ.get("GBN_SMALL") == "2" else 2)
    P = oracle.PortModel(pb)
    n, nX, E = 12, G.n_tf(), G.n_edges()
    rng = np.random.default_rng(5)
    flat = rng.random((2, n, nX + E))
    beta = rng.beta(2., 5., (2, n, nX))
    expo = rng.exponential(1., (2, n, nX))
    G.inject(flat, beta, expo)
    for c in range(2):
        P.inject(c, flat[c], beta[c], expo[c])
    G.sample_n(5); G.sample_n(7)
    P.sample_n(12)
    for c in range(2):
        a, p = G.get_state(c), P.get_state(c)
        assert np.array_equal(a[:nX], p[:nX]) and np.array_equal(a[2 * nX:], p[2 * nX:])
        assert np.abs(a[nX:2 * nX] - p[nX:2 * nX]).max() < 1e-9


def test_long_batches_fold_the_byte_counters():
    """The S nodes count into 8-bit per-batch counters that are folded into the 32-bit totals before one could
    overflow: 700 sweeps in one call (three folds inside) give the counts of 700 single calls."""
    pb, _ = make_problem(NX=10, NY=80, seed=31, n_graphs=2)
    for env in ("1", "2"):
        os.environ["GBN_SMALL"] = env
        try:
            A, B = sampler_from_problem(pb), sampler_from_problem(pb)
        finally:
            del os.environ["GBN_SMALL"]
        A.sample_n(700)
        for n in (255, 1, 254, 190):
            B.sample_n(n)
        for c in range(2):
            assert np.array_equal(A.get_state(c)[:10], B.get_state(c)[:10])
            ma, va = A.node_moments(c)
            mb, vb = B.node_moments(c)
            assert np.array_equal(ma[30:], mb[30:]) and np.array_equal(va[30:], vb[30:])
        P = oracle.PortModel(pb)
        P.sample_n(700)
        assert rel_err(A.gelman_rubin(), P.gelman_rubin()).max() < 1e-9


# ------------------------------------------------------------------------------------------- protocol, contrasts
def test_run_protocol_is_the_python_protocol():
    """gbn_model_run_protocol (one call) = the loop of gbnet/utils.py:93-135 written out in Python."""
    from gbnet_b200 import utils
    pb, _ = make_problem(NX=20, NY=300, AvgNTF=2.5, seed=3, n_graphs=3)
    A, B = sampler_from_problem(pb), sampler_from_problem(pb)
    its_a, rhat_a = A.run_protocol(burn_its=4, max_its=40, N=20, dN=30)
    its_b = 0
    for _ in range(4):
        its_b += 1
        B.sample(20, 30)
    B.burn_stats()
    rhat_b = float("inf")
    while rhat_b > 1.1:
        its_b += 1
        B.sample(20, 30)
        rhat_b = B.max_gelman_rubin()
        if its_b >= 40:
            break
    assert (its_a, rhat_a) == (its_b, rhat_b)
    assert A.sweep_count() == B.sweep_count() == 30 * its_a
    assert np.array_equal(A.posterior_means("X"), B.posterior_means("X"))
    for c in range(3):
        assert np.array_equal(A.get_state(c), B.get_state(c))
    # utils.sample_model goes through the fused call and returns the reference's result table
    net = __import__("gbnet_b200").synth.gen_network(NX=15, NY=200, AvgNTF=2.5, num_active_tfs=2, seed=8)
    ents, rels, evid, tests = utils.frames_from_network(net)
    model = utils.get_model(ents, rels, evid, n_graphs=3, seed=4)
    res = utils.sample_model(model, ents, tests, max_its=30)
    assert {"X", "T", "enrichment", "mor_binary"} <= set(res.columns) and len(res) == len(tests)


def test_from_contrasts_marshals_once():
    """ModelORNOR.from_contrasts(ents, rels, [evid ...]): contrast 0 is the single-contrast model, every
    contrast equals the integer-level batched Sampler, and evidence frames may list different targets."""
    import gbnet_b200
    from gbnet_b200 import synth, utils
    net = synth.gen_network(NX=15, NY=150, AvgNTF=2.5, num_active_tfs=2, seed=19)
    ents, rels, evid, _ = utils.frames_from_network(net)
    evs = [evid]
    for i in range(2):
        v, _ = synth.redraw_evidence(net, num_active_tfs=2, seed=60 + i)
        f = evid.copy()
        f["val"] = v
        evs.append(f.iloc[i + 1:])                                  # frames of different length / target sets
    kw = dict(n_graphs=3, seed=12, noise_listen_children=False)
    M = gbnet_b200.ModelORNOR.from_contrasts(ents, rels.copy(), evs, **kw)
    assert M.n_contrasts == 3
    S = gbnet_b200.ModelORNOR(ents, rels.copy(), evid, **kw)
    M.sample(60, 30)
    S.sample(60, 30)
    assert M.get_posterior_means("X", dataset=0) == S.get_posterior_means("X")
    assert M.get_gelman_rubin(0) == S.get_gelman_rubin()
    for d in (1, 2):
        Sd = gbnet_b200.ModelORNOR(ents, rels.copy(), evs[d], **kw)
        Sd.sample(60, 30)
        # other global chain ids, hence other draws: same posterior within Monte-Carlo error is not a bit test;
        # what must agree exactly is the evidence the contrast sees
        a = M.sampler.all_conditionals(3 * d)
        Sd.sampler.set_state(0, M.sampler.get_state(3 * d))
        b = Sd.sampler.all_conditionals(0)
        m = ~np.isnan(b)
        assert rel_err(a[m], b[m]).max() < 1e-12


def test_fisher_counts_on_the_device_match_fisher_exact():
    """get_tests_batched (device segmented reduction of the per-TF contingency counts) = get_tests per
    contrast = scipy.stats.fisher_exact per TF (reference gbnet/utils.py:10-68), with targets missing from a
    contrast's evidence, edges of type 0 and duplicate edge rows."""
    import pandas as pd
    from scipy.stats import fisher_exact
    from gbnet_b200 import synth, utils
    net = synth.gen_network(NX=30, NY=500, AvgNTF=3.0, num_active_tfs=3, seed=77)
    ents, rels, evid, _ = utils.frames_from_network(net)
    rels = pd.concat([rels, rels.iloc[:7]], ignore_index=True)     # duplicate rows count twice in the group-by
    rels.loc[rels.index[::11], "type"] = 0
    evs = [evid]
    for i in range(3):
        v, _ = synth.redraw_evidence(net, num_active_tfs=3, seed=200 + i)
        f = evid.copy()
        f["val"] = v
        evs.append(f.iloc[(7 * i) % 50:].iloc[::2] if i else f)    # contrasts with missing targets
    got = utils.get_tests_batched(evs, rels)
    assert len(got) == len(evs)
    for d, ev in enumerate(evs):
        want = utils.get_tests(ev, rels)
        g = got[d].loc[want.index]
        for col in ("trg_ydeg", "trg_ndeg", "morp_degp", "morp_degn", "morn_degp", "morn_degn", "trg_tot"):
            assert (g[col] == want[col]).all(), (d, col)
        for col in ("enrichment", "mor_binary", "combined"):
            assert np.allclose(g[col], want[col], rtol=1e-12, atol=0)
    row = got[1].iloc[0]
    p = fisher_exact([[row.morp_degp, row.morn_degp], [row.morp_degn, row.morn_degn]], alternative="greater")[1]
    assert abs(p - row.mor_binary) < 1e-12


# ------------------------------------------------------------------------------------------- guards, printing
def test_t_exactly_one_is_refused_by_the_fast_kernels():
    """T = 1.0 with X = 1 makes the edge factor zero; the fast kernels carry its reciprocal, so the state is
    refused there (GBN_ERR_UNSUPPORTED) while the strict kernel evaluates it as the reference does."""
    import gbnet_b200
    pb, _ = make_problem(NX=8, NY=40, seed=2, n_graphs=1)
    P = oracle.PortModel(pb)
    st = P.get_state(0)
    nX = 8
    st[0] = 1.
    st[nX] = 1.0
    for env in ("1", "2"):
        os.environ["GBN_SMALL"] = env
        try:
            G = sampler_from_problem(pb)
        finally:
            del os.environ["GBN_SMALL"]
        with pytest.raises(gbnet_b200.GbnError, match="T = 1.0"):
            G.set_state(0, st)
    S = sampler_from_problem(pb, kernel=gbnet_b200.KERNEL_STRICT)
    S.set_state(0, st)
    P.set_state(0, st)
    a, b = S.all_conditionals(0), P.all_conditionals(0)
    m = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), np.isfinite(b)) and rel_err(a[m], b[m]).max() < 1e-12


def test_print_stats_output(capfd):
    """gbn_model_print_stats (ModelBase::print_stats, ModelBase.cpp:144-158): one line "<id>_<stat>\\tmean(sd)" per
    statistic over chains 1..M-1, X then T then S."""
    pb, _ = make_problem(NX=6, NY=30, seed=9, n_graphs=4)
    G = sampler_from_problem(pb)
    G.sample_n(40)
    G.print_stats()
    out = capfd.readouterr().out.strip().splitlines()
    nX, E = G.n_tf(), G.n_edges()
    assert len(out) == 2 * nX + nX + 3 * E
    mu, sd = G.posterior_means("X"), G.posterior_sdevs("X")
    tf = G.tf_ids()
    for i in (0, 1, 2 * nX - 1):
        assert out[i] == "X_%u_%u\t%.4f(%.4f)" % (tf[i // 2], i % 2, mu[i], sd[i])
    assert out[2 * nX].startswith("T_%u_0\t" % tf[0]) and out[3 * nX].startswith("S_")
    assert "-->" in out[-1] and out[-1].split("\t")[0].endswith("_2")


def test_zero_length_calls_are_no_ops():
    pb, _ = make_problem(NX=6, NY=30, seed=9, n_graphs=2)
    G = sampler_from_problem(pb)
    G.sample_n(0)
    G.sample(0, 30)
    assert G.sweep_count() == 0 and G.last_device_ms() == 0.
    G.set_stream_groups(7)                     # more groups than chain bundles: no empty group
    G.sample_n(3)
    assert G.sweep_count() == 3


# ------------------------------------------------------------------------------------------- full shapes
@pytest.mark.parametrize("small", ["1", "3"])
def test_c4_full_shape_batched_contrasts(monkeypatch, small):
    """BASELINE config C4 at full size on one device: the C2 network, 512 contrasts x 8 chains (4,096 chains), on
    the multi-kernel path (the default for more chains than SMs) and on the persistent kernel (GBN_SMALL=3).
    Contrast 0 draws the chains of the single-contrast model bit for bit; a late contrast's conditionals match
    the oracle holding that contrast's evidence; the maximum R-hat is the maximum over contrasts."""
    from gbnet_b200 import synth
    import gbnet_b200
    monkeypatch.setenv("GBN_SMALL", small)
    net = synth.shape("C2")
    hy = synth.cli_default_hyper()
    evs = np.stack([net.evid_val] + [synth.redraw_evidence(net, num_active_tfs=8, seed=100 + i)[0] for i in range(511)])
    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, evs, n_graphs=8, seed=13, **hy)
    assert G.n_chains() == 4096 and G.path() == (2 if small == "3" else 1)
    S = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=8, seed=13, **hy)
    G.sample_n(30)
    S.sample_n(30)
    for c in (0, 3, 7):
        assert np.array_equal(G.get_state(c), S.get_state(c))
    assert np.array_equal(G.gelman_rubin(0), S.gelman_rubin())
    d = 509
    pb = oracle.Problem(src=net.src, trg=net.trg, mor=net.mor, evid_trg=net.evid_trg, evid_val=evs[d], n_graphs=1, **hy)
    P = oracle.PortModel(pb)
    c = 8 * d + 5
    P.set_state(0, G.get_state(c))
    a, b = G.all_conditionals(c), P.all_conditionals(0)
    m = ~np.isnan(b)
    assert rel_err(a[m], b[m]).max() < 1e-12
    per = [np.nanmax(np.where(np.isfinite(G.gelman_rubin(q)), G.gelman_rubin(q), 0.)) for q in (0, 100, 511)]
    assert G.max_gelman_rubin() >= max(per)


def test_c5_full_shape_chain_scaling():
    """BASELINE config C5 at full size on one device: 4,096 chains of the 300k-edge network.  Chains are keyed
    by their global id: chains 0-3 and 4,092-4,095 equal the chains a 4-chain shard of the same model draws."""
    from gbnet_b200 import synth
    import gbnet_b200
    net = synth.shape("C3")
    hy = synth.cli_default_hyper()
    args = (net.src, net.trg, net.mor, net.evid_trg, net.evid_val)
    G = gbnet_b200.Sampler(*args, n_graphs=4096, seed=5, **hy)
    assert G.path() == 1
    lo = gbnet_b200.Sampler(*args, n_graphs=4096, n_graphs_local=4, chain_offset=0, seed=5, **hy)
    hi = gbnet_b200.Sampler(*args, n_graphs=4096, n_graphs_local=4, chain_offset=4092, seed=5, **hy)
    for M in (G, lo, hi):
        M.sample_n(3)
    for c in range(4):
        assert np.array_equal(G.get_state(c), lo.get_state(c))
        assert np.array_equal(G.get_state(4092 + c), hi.get_state(c))
    gr = G.max_gelman_rubin()
    assert np.isfinite(gr) and gr >= 1.0


# ------------------------------------------------------------------------------------------- 2 ranks over NCCL
def _nccl_worker(rank, world, q_uid, q_out):
    import gbnet_b200
    from gbnet_b200.distributed import shard_chains
    pb, _ = make_problem(NX=20, NY=300, AvgNTF=2.5, seed=3, n_graphs=6)
    off, loc = shard_chains(6, rank, world)
    G = sampler_from_problem(pb, device=rank, chain_offset=off, n_graphs_local=loc)
    if rank == 0:
        uid = gbnet_b200.Sampler.comm_unique_id()
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=120)
    G.attach_comm(rank, world, uid)
    G.sample_n(45)
    q_out.put((rank, G.max_gelman_rubin(), G.gelman_rubin(), G.posterior_means("X"), G.posterior_sdevs("S"),
               G.posterior_means("T")))
    G.close()


def test_two_rank_moment_exchange_over_nccl():
    """Chains sharded over two GPUs, the compact u64 + fp64 moment exchange through ncclAllReduce: both ranks hold
    the unsharded model's R-hat and posterior summaries (the exchange the driver's scaling run exercises)."""
    import multiprocessing as mp

    import gbnet_b200
    if gbnet_b200.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, q_uid, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q_out.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pb, _ = make_problem(NX=20, NY=300, AvgNTF=2.5, seed=3, n_graphs=6)
    F = sampler_from_problem(pb)
    F.sample_n(45)
    for r in res:
        assert r[1] == F.max_gelman_rubin()
        assert np.array_equal(r[2][:20], F.gelman_rubin()[:20]) and np.array_equal(r[2][40:], F.gelman_rubin()[40:])
        assert rel_err(r[2][20:40], F.gelman_rubin()[20:40]).max() < 1e-12          # T nodes: floating-point sums
        assert np.array_equal(r[3], F.posterior_means("X")) and np.array_equal(r[4], F.posterior_sdevs("S"))
        assert rel_err(r[5], F.posterior_means("T")).max() < 1e-12
