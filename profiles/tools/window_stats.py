"""Speculative X windows: rounds per sweep, TFs committed per round and rounds ended by a flip, per
chain bundle (the X kernel sweeps two chains per CTA).  python profiles/tools/window_stats.py [C2|C3] [listen]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import gbnet_b200
from gbnet_b200 import synth

shape = sys.argv[1] if len(sys.argv) > 1 else "C3"
net = synth.shape(shape)
hy = dict(synth.cli_default_hyper())
if len(sys.argv) > 2:
    hy["noise_listen_children"] = sys.argv[2] == "listen"
M = 64
G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=M, seed=1, **hy)
G.debug_counters(True)
bundles = M // 2
for blk in range(6):
    G.sample_n(10)
    r, t, f = G.debug_counters(True)[:3]
    print("%s sweeps %d-%d: rounds/sweep/bundle %.1f  TFs/round %.2f  flip rounds/sweep/bundle %.1f"
          % (shape, blk * 10, blk * 10 + 9, r / (10 * bundles), t / max(r, 1), f / (10 * bundles)))
st = G.get_state(0)
nX = G.n_tf()
print("X active", st[:nX].sum(), "S dist", np.bincount(st[2 * nX:].astype(int)))
