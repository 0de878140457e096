"""Per-kernel time of one sweep on the small BASELINE configs (C1, C2, one GPU's share of C4)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gbnet_b200
from gbnet_b200 import synth

hy = synth.cli_default_hyper()
for shape, chains, nds in (("C1", 3, 1), ("C2", 64, 1), ("C2", 8, 64)):
    net = synth.shape(shape)
    ev = net.evid_val if nds == 1 else np.stack([synth.redraw_evidence(net, num_active_tfs=8, seed=100 + i)[0] for i in range(nds)])
    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, ev, n_graphs=chains, seed=11, **hy)
    G.sample_n(20)
    G.synchronize()
    t = time.perf_counter(); G.sample_n(200); G.synchronize(); wall = (time.perf_counter() - t) * 1e3
    G.timer_start(); G.sample_n(200); dev = G.timer_stop()
    G.set_stream_groups(1); G.set_phase_timing(True); G.sample_n(30); G.synchronize()
    ph, nl = G.phase_ms()
    print(f"{shape} chains={chains} datasets={nds} E={net.n_edges}: wall {wall/200*1e3:.1f} us/sweep, device {dev/200*1e3:.1f} us/sweep, "
          f"serialised per-kernel us/sweep X,T,S = {[round(p / max(nl, 1) * 1e3, 1) for p in ph]}, "
          f"rate {G.n_chains() * 200 * net.n_edges / (dev * 1e-3):.3e}", flush=True)
    G.close()
