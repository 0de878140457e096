"""Where an end-to-end step (bench.py's e2e leg) spends its wall time, per API call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gbnet_b200
from gbnet_b200 import synth

net = synth.shape("C3")
hy = synth.cli_default_hyper()
G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=256, **hy)
evs = [synth.redraw_evidence(net, num_active_tfs=12, seed=500 + i)[0] for i in range(4)]
G.set_evidence(net.evid_trg, evs[-1]); G.sample_n(30); G.max_gelman_rubin()
acc = {}
for i in range(3):
    for name, fn in (("set_evidence", lambda: G.set_evidence(net.evid_trg, evs[i])),
                     ("sample_n(30)", lambda: (G.sample_n(30), G.synchronize())),
                     ("max_gelman_rubin", lambda: G.max_gelman_rubin()),
                     ("posterior_means(X)", lambda: G.posterior_means("X"))):
        t = time.perf_counter(); fn(); acc.setdefault(name, []).append((time.perf_counter() - t) * 1e3)
for k, v in acc.items():
    print(f"{k:22s} {np.mean(v):8.3f} ms  {['%.3f' % x for x in v]}")
