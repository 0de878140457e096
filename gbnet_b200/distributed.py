"""Chains sharded over ranks (one process per GPU): helpers around the C ABI's multi-GPU entry points.

The sampling path has no collective: chains are independent (reference ModelBase.cpp:133-135), so
rank r simply owns a contiguous block of every dataset's chains.  The only exchange step is the
cross-chain moment reduction behind Gelman-Rubin and the posterior getters (ModelBase.cpp:47-112,
160-209), done inside the library with ncclAllReduce on the model's stream
(csrc/capi.cu: compute_moments).  ``two_pass_gelman_rubin`` restates that protocol on host arrays
with any ``torch.distributed`` backend; the CPU test-suite runs it with gloo at world_size 2 against
the oracle, which pins the decomposition the CUDA kernels in csrc/moments.cu implement.
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
