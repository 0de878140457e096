"""Where the end-to-end step of bench.py spends its wall time (C3, 256 chains): set_evidence(host arrays), sample_n(dN),
max R-hat, posterior X means - each call timed on the host (the calls synchronise) - and the stages inside set_evidence
(GBN_TRACE=1 prints them on stderr)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gbnet_b200  # noqa: E402
from gbnet_b200 import synth  # noqa: E402

M, dN, K = 256, 30, 10
net = synth.shape("C3")
hy = synth.cli_default_hyper()
G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=M, seed=3, **hy)
evs = [synth.redraw_evidence(net, num_active_tfs=12, seed=500 + i)[0] for i in range(K + 1)]
G.set_evidence(net.evid_trg, evs[-1])
G.sample_n(dN)
G.max_gelman_rubin()
t = np.zeros(4)
for i in range(K):
    a = time.perf_counter(); G.set_evidence(net.evid_trg, evs[i])
    b = time.perf_counter(); G.sample_n(dN)
    c = time.perf_counter(); G.max_gelman_rubin()
    d = time.perf_counter(); G.posterior_means("X")
    e = time.perf_counter()
    t += [b - a, c - b, d - c, e - d]
print("ms per call: set_evidence %.3f  sample_n(%d) %.3f  max_gelman_rubin %.3f  posterior_means(X) %.3f  total %.3f"
      % (*(1e3 * t / K), dN, 1e3 * t.sum() / K) if False else
      "ms per call: set_evidence %.3f  sample_n(%d) %.3f  max_gelman_rubin %.3f  posterior_means(X) %.3f  total %.3f"
      % (1e3 * t[0] / K, dN, 1e3 * t[1] / K, 1e3 * t[2] / K, 1e3 * t[3] / K, 1e3 * t.sum() / K))
os.environ["GBN_TRACE"] = "1"
G.set_evidence(net.evid_trg, evs[0])
