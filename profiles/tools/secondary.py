"""Secondary measurements on the GPU box (not the bench line): the other BASELINE.json configs and
the "time to R-hat < 1.1" protocol of gbnet/utils.py:89-135 (10 x sample(20), burn_stats, then
sample(20) until max R-hat <= 1.1), CUDA path vs the compiled reference on the host cores."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import gbnet_b200  # noqa: E402
import oracle  # noqa: E402
from gbnet_b200 import synth  # noqa: E402

hy = synth.cli_default_hyper()
out = {}


def protocol(model, max_its=200):
    t0 = time.perf_counter()
    it = 0
    for _ in range(10):
        it += 1
        model.sample(20, 30)
    model.burn_stats()
    mgr = float("inf")
    while mgr > 1.1 and it < max_its:
        it += 1
        model.sample(20, 30)
        mgr = model.max_gelman_rubin()
    return time.perf_counter() - t0, it, mgr


def rate(G, E, n_sweeps=60):
    G.sample_n(10)
    G.timer_start()
    G.sample_n(n_sweeps)
    ms = G.timer_stop()
    return G.n_chains() * n_sweeps * E / (ms * 1e-3)


for shape, chains in (("C1", 3), ("C2", 64)):
    net = synth.shape(shape)
    E = net.n_edges
    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=chains, seed=11, **hy)
    r = rate(G, E)
    G.close()
    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=chains, seed=12, **hy)
    tg, itg, mg = protocol(G)
    G.close()
    pb = oracle.Problem(src=net.src, trg=net.trg, mor=net.mor, evid_trg=net.evid_trg, evid_val=net.evid_val,
                        n_graphs=chains, **hy)
    R = oracle.RefModel(pb) if oracle.have_ref() else oracle.PortModel(pb)
    R._f("omp_set_threads")(min(os.cpu_count() or 1, chains))
    tr, itr, mr = protocol(R)
    sweeps_ref = R.chain_length()
    R.close()
    out[shape] = dict(chains=chains, edges=E, gpu_edge_updates_per_s=r,
                      gpu_time_to_rhat_s=tg, gpu_iterations=itg, gpu_max_rhat=mg,
                      ref_time_to_rhat_s=tr, ref_iterations=itr, ref_max_rhat=mr,
                      ref_kind="reference" if oracle.have_ref() else "port", host_cores=os.cpu_count())
    print(shape, out[shape], flush=True)

# C4: batched contrasts on the C2 network, one GPU's share of 512 datasets x 8 chains over 8 GPUs
net = synth.shape("C2")
evs = np.stack([synth.redraw_evidence(net, num_active_tfs=8, seed=100 + i)[0] for i in range(64)])
G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, evs, n_graphs=8, seed=13, **hy)
out["C4_one_gpu_share"] = dict(datasets=64, chains_per_dataset=8, edges=net.n_edges,
                               gpu_edge_updates_per_s=rate(G, net.n_edges))
t0 = time.perf_counter()
G.sample(300, 30)
out["C4_one_gpu_share"]["sample_300_s"] = time.perf_counter() - t0
out["C4_one_gpu_share"]["max_rhat_over_datasets"] = G.max_gelman_rubin()
G.close()
print("C4", out["C4_one_gpu_share"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "secondary.json"), "w"), indent=1)
