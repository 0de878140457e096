"""Sharding self-check, stand-alone (the `parity_check` of bench.py's secondary block with the details printed):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/tools/shard_check.py

A small model's chains sharded over the N ranks (NCCL moment exchange) against the same model on one rank: R-hat of
every X / S node and the posterior X / S means must be identical (integer count sums cross the ranks), the T nodes'
to 1e-12 (fp64 sums reduced in another order)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import gbnet_b200  # noqa: E402
from gbnet_b200 import synth  # noqa: E402
from gbnet_b200.distributed import shard_chains  # noqa: E402

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
hy = synth.cli_default_hyper()
for per_rank, sweeps in ((4, 60), (3, 45)):
    nets = synth.gen_network(NX=40, NY=800, AvgNTF=3.0, num_active_tfs=4, seed=17)
    M = per_rank * world
    off, loc = shard_chains(M, rank, world)
    G = gbnet_b200.Sampler(nets.src, nets.trg, nets.mor, nets.evid_trg, nets.evid_val, n_graphs=M, chain_offset=off,
                           n_graphs_local=loc, seed=7, device=local_rank, **hy)
    bench._attach(G, dist, rank, world)
    G.sample_n(sweeps)
    gr, grn = G.max_gelman_rubin(), np.asarray(G.gelman_rubin())
    px, ps, pt = G.posterior_means("X"), G.posterior_means("S"), G.posterior_means("T")
    sx, st = G.posterior_sdevs("X"), G.posterior_sdevs("T")
    G.close()
    if rank == 0:
        F = gbnet_b200.Sampler(nets.src, nets.trg, nets.mor, nets.evid_trg, nets.evid_val, n_graphs=M, seed=7, device=local_rank, **hy)
        F.sample_n(sweeps)
        fgr, fgrn = F.max_gelman_rubin(), np.asarray(F.gelman_rubin())
        nX = F.n_tf()
        disc = np.r_[0:nX, 2 * nX:len(fgrn)]
        bad = np.flatnonzero(fgrn[disc] != grn[disc])
        print(json.dumps({
            "world": world, "chains": M, "sweeps": sweeps, "max_rhat_sharded": gr, "max_rhat_single": fgr,
            "argmax_single": int(np.argmax(fgrn)), "n_tf": nX,
            "rhat_discrete_nodes_identical": bool(len(bad) == 0), "n_discrete_nodes_differing": int(len(bad)),
            "max_rel_diff_discrete": float(np.max(np.abs(fgrn[disc] - grn[disc]) / fgrn[disc])),
            "rhat_t_nodes_max_rel_diff": float(np.max(np.abs(fgrn[nX:2 * nX] - grn[nX:2 * nX]) / fgrn[nX:2 * nX])),
            "x_means_identical": bool(np.array_equal(F.posterior_means("X"), px)),
            "s_means_identical": bool(np.array_equal(F.posterior_means("S"), ps)),
            "x_sdevs_max_abs_diff": float(np.max(np.abs(F.posterior_sdevs("X") - sx))),
            "t_means_max_abs_diff": float(np.max(np.abs(F.posterior_means("T") - pt))),
            "t_sdevs_max_abs_diff": float(np.max(np.abs(F.posterior_sdevs("T") - st)))}), flush=True)
        F.close()
    dist.barrier()
dist.destroy_process_group()
