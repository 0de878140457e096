"""How often the X kernel carries the magnitude bound and how often it needs the exact product (C3, 64 chains)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gbnet_b200
from gbnet_b200 import synth
net = synth.shape("C3")
hy = synth.cli_default_hyper()
G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=64, seed=1, **hy)
G.sample_n(30)
G.debug_counters(True)
G.sample_n(10)
d = G.debug_counters(False)
n = 10 * 64
if os.environ.get("GBN_X1") == "2":      # chain-pair kernel k_fast_x
    print("k_fast_x, per chain and sweep: rounds %.1f (x2 chains/bundle), TF evaluations carrying the bound %.1f, exact products %.2f "
          "(second passes %.2f; updates that gathered the likelihoods in the main pass %.2f), active-parent lookups in S %.0f of %d edges"
          % (2 * d[0] / n, d[9] / n, d[8] / n, d[10] / n, d[11] / n, d[7] / n, net.n_edges))
else:                                    # single-chain kernel k_fast_x1
    print("k_fast_x1, per chain and sweep: rounds %.1f, TF evaluations carrying the bound %.1f, TFs resolved in fp64 %.2f (of them with the "
          "exact product of the likelihoods %.2f), draws from the prior decided by the bound alone %.2f, active-parent lookups in S %.0f of %d edges"
          % (d[0] / n, d[9] / n, d[10] / n, d[8] / n, d[11] / n, d[7] / n, net.n_edges))
