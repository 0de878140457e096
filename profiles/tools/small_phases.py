"""Cycles per phase of the persistent small-network kernel (sweep_small.cu), from its clock64 diagnostics:
mean cycles of thread 0 per sweep in the X, T and S phases, for BASELINE shapes C1 and C2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import gbnet_b200  # noqa: E402
from gbnet_b200 import synth  # noqa: E402

hy = synth.cli_default_hyper()
for shape, chains in (("C1", 3), ("C2", 64)):
    net = synth.shape(shape)
    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=chains, seed=11, **hy)
    assert G.path() == 2
    G.sample_n(30)
    G.debug_counters(True)
    G.timer_start()
    G.sample_n(60)
    ms = G.timer_stop()
    d = G.debug_counters(False)
    n = 60 * chains
    print(f"{shape} x {chains}: {1e3 * ms / 60:.1f} us per sweep; rounds/sweep {d[0] / n:.1f} (flips {d[2] / n:.1f}); "
          f"cycles per sweep X {d[4] / n:.0f} T {d[5] / n:.0f} S {d[6] / n:.0f}", flush=True)
    G.close()
