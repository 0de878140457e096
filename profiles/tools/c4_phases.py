"""Per-kernel time of a sweep on BASELINE config C4 (the C2 network, 512 contrasts x 8 chains on one GPU) on the
multi-kernel path, kernels serialised on one stream, and the overlapped rate."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import gbnet_b200  # noqa: E402
from gbnet_b200 import synth  # noqa: E402

net = synth.shape("C2")
hy = synth.cli_default_hyper()
nd = int(os.environ.get("C4_CONTRASTS", "512"))
evs = np.stack([net.evid_val] + [synth.redraw_evidence(net, num_active_tfs=8, seed=100 + i)[0] for i in range(nd - 1)])
G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, evs, n_graphs=8, seed=13, **hy)
E = G.n_edges()
G.sample_n(30)
G.synchronize()
G.timer_start(); G.sample_n(60); ms = G.timer_stop()
print(f"C4 {nd} x 8 chains path {G.path()}: {1e3 * ms / 60:.1f} us per sweep, {G.n_chains() * E * 60 / (ms * 1e-3):.3e} edge-updates/s")
if G.path() == 1:
    G.set_stream_groups(1)
    G.set_phase_timing(True)
    G.sample_n(30)
    pm, n = G.phase_ms()
    G.set_phase_timing(False)
    print("serialised ms per sweep: X %.4f  T %.4f  S %.4f" % tuple(p / n for p in pm))
G.close()
