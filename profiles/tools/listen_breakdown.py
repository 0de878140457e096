"""Sweep time with T nodes that listen to their children (the pyx default) against the CLI default."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gbnet_b200
from gbnet_b200 import synth

for shape, chains in (("C2", 64), ("C3", 256)):
    net = synth.shape(shape)
    for listen in (False, True):
        hy = dict(synth.cli_default_hyper()); hy["noise_listen_children"] = listen
        G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=chains, seed=11, **hy)
        G.sample_n(30)
        G.timer_start(); G.sample_n(60); dev = G.timer_stop()
        G.set_stream_groups(1); G.set_phase_timing(True); G.sample_n(30); G.synchronize()
        ph, nl = G.phase_ms()
        st = G.get_state(0)
        print(f"{shape} chains={chains} listen={listen}: {dev/60*1e3:.1f} us/sweep, serialised X,T,S us = "
              f"{[round(p / max(nl, 1) * 1e3, 1) for p in ph]}, active TFs in chain 0: {int(st[:G.n_tf()].sum())}, "
              f"rate {G.n_chains() * 60 * net.n_edges / (dev * 1e-3):.3e}", flush=True)
        G.close()
