"""Chains sharded over ranks (one process per GPU): helpers around the C ABI's multi-GPU entry points.

The sampling path has no collective: chains are independent (reference ModelBase.cpp:133-135), so
rank r simply owns a contiguous block of every dataset's chains.  The only exchange step is the
cross-chain moment reduction behind Gelman-Rubin and the posterior getters (ModelBase.cpp:47-112,
160-209), done inside the library with ncclAllReduce on the model's stream
(csrc/capi.cu: compute_moments): ONE u64 all-reduce of exact sums of the stored outcome counts for the
discrete nodes (``exact_count_moments``: the layout and algebra of csrc/moments.cu k_count_sums /
k_expand_counts) and two small fp64 all-reduces for the T nodes (``two_pass_gelman_rubin``).  Both are
restated here on host arrays with any ``torch.distributed`` backend; the CPU test-suite runs them with gloo
at world_size 2 against the oracle, which pins the decomposition the CUDA kernels implement.
"""
from __future__ import annotations

import numpy as np


def shard_chains(n_graphs: int, rank: int, world: int):
    """Contiguous block partition of n_graphs chains: returns (chain_offset, n_graphs_local)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_graphs), int(world))
    local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, local


def attach_torch_comm(sampler, dist=None):
    """Create the library's NCCL communicator for the ranks of the default torch.distributed group
    (the unique id travels through torch.distributed; the all-reduces do not)."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    uid = [type(sampler).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sampler.attach_comm(rank, world, uid[0])


def two_pass_gelman_rubin(mean_local, var_local, N, M, chain_offset, allreduce_sum):
    """The multi-rank protocol of csrc/moments.cu on host arrays.

    mean_local / var_local: [local chains, n_stats] per-chain running mean and variance;
    N: chain length; M: global number of chains; allreduce_sum(array) -> array summed over ranks.
    Returns (R per stat, posterior mean per stat, posterior sd per stat) - identical on all ranks.
    """
    mean_local = np.asarray(mean_local, np.float64)
    var_local = np.asarray(var_local, np.float64)
    post = (chain_offset + np.arange(mean_local.shape[0])) >= 1          # chain 0 is skipped (ModelBase.cpp:172)
    s1 = allreduce_sum(np.stack([mean_local.sum(0), var_local.sum(0), mean_local[post].sum(0)]))
    mu, W = s1[0] / M, s1[1] / M
    pmu = s1[2] / (M - 1.) if M > 1 else np.zeros_like(mu)
    s2 = allreduce_sum(np.stack([((mean_local - mu) ** 2).sum(0), ((mean_local[post] - pmu) ** 2).sum(0)]))
    B = N * (s2[0] / (M - 1.) if M > 1 else 0. * s2[0])
    Vhat = W * (N - 1.) / N + B * (M + 1.) / (M * N)
    with np.errstate(divide="ignore", invalid="ignore"):
        R = np.where((W > 0) & (Vhat > 0), np.sqrt((N + 3.) / (N + 1.) * Vhat / np.where(W > 0, W, 1.)),
                     np.where(Vhat > 0, np.inf, 1.))
    n = M - 1.
    psd = np.sqrt(s2[1] / (n - 1.)) if n > 1 else np.zeros_like(mu)
    return R, pmu, psd


def exact_count_moments(cx1_local, cs_local, N, M, chain_offset, allreduce_sum):
    """The integer half of the protocol (csrc/moments.cu), on host arrays.

    cx1_local: [local chains, nX] sweeps with X = 1; cs_local: [local chains, E, 2] sweeps with S = 0 and S = 2
    (the counters the kernels keep); N: chain length; M: global number of chains.  What crosses the ranks:
    per X node {S1, S2, z}, per S node {A1, A2, B1, B2, AB, z} as unsigned 64-bit sums, z = the counts of
    GLOBAL chain 0 (a | b << 32 for S), zero on every rank that does not own it.  The complementary outcomes
    (X = 0, S = 1) follow from c = N - stored.  Returns dict(sums[stat, 2], psums, dev, pdev) in the stat order
    X (2 per TF) then S (3 per edge) - the quantities ModelBase.cpp:47-112, 160-209 derive R-hat and the posterior
    summaries from.  Python integers: exact, as the kernel's 128-bit arithmetic is.
    """
    cx1 = np.asarray(cx1_local, np.uint64)
    cs = np.asarray(cs_local, np.uint64)
    own0 = chain_offset == 0 and cx1.shape[0] > 0
    nX, E = cx1.shape[1], cs.shape[1]
    msg = np.zeros(3 * nX + 6 * E, np.uint64)
    mx, ms = msg[:3 * nX].reshape(nX, 3), msg[3 * nX:].reshape(E, 6)
    mx[:, 0] = cx1.sum(0); mx[:, 1] = (cx1 * cx1).sum(0)
    a, b = cs[:, :, 0], cs[:, :, 1]
    ms[:, 0] = a.sum(0); ms[:, 1] = (a * a).sum(0); ms[:, 2] = b.sum(0); ms[:, 3] = (b * b).sum(0); ms[:, 4] = (a * b).sum(0)
    if own0:
        mx[:, 2] = cx1[0]
        ms[:, 5] = a[0] | (b[0] << np.uint64(32))
    tot = allreduce_sum(msg.view(np.int64)).view(np.uint64)              # gloo has no u64: the bit patterns add the same
    tx, ts = tot[:3 * nX].reshape(nX, 3), tot[3 * nX:].reshape(E, 6)
    N, M = int(N), int(M)
    rows = []                                                            # (S1, S2, c0) per stat, exact
    for k in range(nX):
        S1, S2, z = (int(v) for v in tx[k])
        rows.append((M * N - S1, M * N * N - 2 * N * S1 + S2, N - z))   # outcome 0: c = N - c1
        rows.append((S1, S2, z))
    for e in range(E):
        A1, A2, B1, B2, AB, z = (int(v) for v in ts[e])
        za, zb = z & 0xffffffff, z >> 32
        rows.append((A1, A2, za))
        rows.append((M * N - A1 - B1, M * N * N - 2 * N * (A1 + B1) + A2 + B2 + 2 * AB, N - za - zb))
        rows.append((B1, B2, zb))
    n_stats = len(rows)
    sums, psums, dev, pdev = np.zeros((n_stats, 2)), np.zeros(n_stats), np.zeros(n_stats), np.zeros(n_stats)
    for j, (S1, S2, c0) in enumerate(rows):
        P1, P2 = S1 - c0, S2 - c0 * c0
        sums[j, 0] = S1 / N
        sums[j, 1] = (N * S1 - S2) / (N * (N - 1.)) if N > 1 else 0.
        psums[j] = P1 / N
        dev[j] = (M * S2 - S1 * S1) / (float(M) * N * N)
        pdev[j] = ((M - 1) * P2 - P1 * P1) / ((M - 1.) * N * N) if M > 1 else 0.
    return dict(sums=sums, psums=psums, dev=dev, pdev=pdev)


def gelman_rubin_from_sums(sums, dev, N, M):
    """ModelBase::get_gelman_rubin (ModelBase.cpp:72-92) per statistic from the globally reduced sums."""
    N, M = float(N), float(M)
    B = N * (dev / (M - 1.) if M > 1 else 0. * dev)
    W = sums[:, 1] / M
    Vhat = W * (N - 1.) / N + B * (M + 1.) / (M * N)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where((W > 0) & (Vhat > 0), np.sqrt((N + 3.) / (N + 1.) * Vhat / np.where(W > 0, W, 1.)),
                        np.where(Vhat > 0, np.inf, 1.))
