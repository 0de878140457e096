"""world_size-2 gloo tests of the multi-rank protocol (host-side): chains sharded over ranks, the exact
integer count sums (discrete nodes) and the two-pass moment all-reduce (T nodes); Gelman-Rubin and the
posterior summaries equal the unsharded oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gbnet_b200.distributed import (exact_count_moments, gelman_rubin_from_sums, shard_chains,  # noqa: E402
                                    two_pass_gelman_rubin)


def test_shard_chains_partitions():
    for n, w in [(4096, 8), (10, 4), (3, 2), (7, 7), (5, 8)]:
        blocks = [shard_chains(n, r, w) for r in range(w)]
        assert sum(b[1] for b in blocks) == n
        pos = 0
        for off, loc in blocks:
            assert off == pos
            pos += loc
    with pytest.raises(ValueError):
        shard_chains(4, 4, 4)


def _stat_table(P, M):
    """per-chain mean / variance of every stat in the reference's order, from the oracle"""
    n_nodes, nX = P.n_nodes(), P.n_tf()
    means, vars_, node_of = [], [], []
    for c in range(M):
        m, v = [], []
        for n in range(n_nodes):
            ns = 2 if n < nX else (1 if n < 2 * nX else 3)
            for s in range(ns):
                m.append(P.node_mean(c, n, s))
                v.append(P.node_variance(c, n, s))
                if c == 0:
                    node_of.append(n)
        means.append(m)
        vars_.append(v)
    return np.array(means), np.array(vars_), np.array(node_of)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from conftest import make_problem
        M = 5
        pb, _ = make_problem(NX=10, NY=50, seed=3, n_graphs=M)
        P = oracle.PortModel(pb)          # every rank runs the same deterministic chains ...
        P.sample_n(40)
        means, vars_, node_of = _stat_table(P, M)
        off, loc = shard_chains(M, rank, world)                  # ... and contributes only its shard

        def allreduce_sum(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.numpy()

        R, pmu, psd = two_pass_gelman_rubin(means[off:off + loc], vars_[off:off + loc], P.chain_length(), M, off,
                                            allreduce_sum)
        gr_nodes = np.zeros(P.n_nodes())
        np.maximum.at(gr_nodes, node_of, R)
        ref_gr = P.gelman_rubin()
        fin = np.isfinite(ref_gr)
        ok = np.array_equal(np.isfinite(gr_nodes), fin)
        ok &= bool(np.max(np.abs(gr_nodes[fin] - ref_gr[fin]) / ref_gr[fin]) < 1e-9)
        ref_pm = np.concatenate([P.posterior_means(k) for k in "XTS"])
        ref_ps = np.concatenate([P.posterior_sdevs(k) for k in "XTS"])
        ok &= bool(np.max(np.abs(pmu - ref_pm)) < 1e-12) and bool(np.max(np.abs(psd - ref_ps)) < 1e-12)

        # the integer route the library takes for the discrete nodes (csrc/moments.cu): per-chain counts are
        # mean * N exactly; X then S statistics
        nX, E, N = P.n_tf(), P.n_edges(), int(P.chain_length())
        counts = np.rint(means * N).astype(np.int64)
        is_s = node_of >= 2 * nX
        cx1 = counts[:, node_of < nX].reshape(M, nX, 2)[:, :, 1]
        cs = counts[:, is_s].reshape(M, E, 3)[:, :, [0, 2]]
        ex = exact_count_moments(cx1[off:off + loc], cs[off:off + loc], N, M, off, allreduce_sum)
        disc = np.concatenate([np.nonzero(node_of < nX)[0], np.nonzero(is_s)[0]])
        R_int = gelman_rubin_from_sums(ex["sums"], ex["dev"], N, M)
        both = np.isfinite(R[disc]) & np.isfinite(R_int)
        ok &= np.array_equal(np.isfinite(R[disc]), np.isfinite(R_int))
        ok &= bool(np.max(np.abs(R_int[both] - R[disc][both]) / R[disc][both]) < 1e-9)
        ok &= bool(np.max(np.abs(ex["psums"] / (M - 1.) - pmu[disc])) < 1e-12)
        ok &= bool(np.max(np.abs(np.sqrt(ex["pdev"] / (M - 2.)) - psd[disc])) < 1e-12)
        q.put((rank, bool(ok), float(np.max(gr_nodes[fin])), float(np.max(ref_gr[fin]))))
    finally:
        dist.destroy_process_group()


def test_two_pass_moments_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2]           # both ranks hold the same global result
