import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gbnet_b200, numpy as np
from gbnet_b200 import synth
net = synth.shape("C3"); hy = synth.cli_default_hyper()
G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=64, seed=1, **hy)
G.debug_counters(True)
for blk in range(4):
    G.sample_n(10)
    r, t, f, _ = G.debug_counters(True)
    print("sweeps %d-%d: rounds/sweep/chain %.1f  TFs/round %.2f  flip rounds/sweep/chain %.1f" % (blk*10, blk*10+9, r/640, t/max(r,1), f/640))
st = G.get_state(0); nX=G.n_tf(); print("X active", st[:nX].sum(), "S dist", np.bincount(st[2*nX:].astype(int)))
