"""How the speculative X windows and the throughput behave when X values change often.

C3 network, 256 chains, three regimes:
  (i)   CLI defaults (12 planted TFs, z = 2000/1 and 200/1, non-listening T): the bench line
  (ii)  the pyx constructor defaults (z = 25/25 for both noise classes, t_focus 2, T nodes listen to their children,
        default S prior 0.4/0.4/0.2): a diffuse posterior, many X values move every sweep
  (iii) CLI defaults with 150 planted TFs (10 % of the TFs active)
For each: rounds per sweep and X-kernel CTA (one chain on k_fast_x1, the default; a chain pair on k_fast_x, GBN_X1=2),
TFs committed per round, rounds ended by a flip, undecided TFs resolved in fp64 per chain and sweep, mean number of
active TFs per chain, and edge-updates/s with the per-kernel times of a serialised pass.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gbnet_b200  # noqa: E402
from gbnet_b200 import synth  # noqa: E402

M, dN = 256, 30
CPC = 2 if os.environ.get("GBN_X1") == "2" else 1          # chains per X-kernel CTA
out = {}
regimes = {
    "cli_defaults_12_planted": (dict(), synth.cli_default_hyper()),
    "pyx_defaults_listening": (dict(), dict(z_alpha=25., z_beta=25., t_focus=2., noise_listen_children=True)),
    "cli_defaults_150_planted": (dict(num_active_tfs=150), synth.cli_default_hyper()),
}
for name, (over, hy) in regimes.items():
    net = synth.shape("C3", **over)
    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=M, seed=3, **hy)
    G.sample_n(60)                                       # past the transient from the initial state
    G.debug_counters(True)
    G.sample_n(dN)
    dbg = G.debug_counters(False)
    r, t, f = dbg[:3]
    G.timer_start()
    for _ in range(3):
        G.sample_n(dN)
        G.max_gelman_rubin()
    ms = G.timer_stop()
    G.set_stream_groups(1)
    G.set_phase_timing(True)
    G.sample_n(dN)
    ph, nph = G.phase_ms()
    G.set_phase_timing(False)
    nX = G.n_tf()
    act = np.mean([G.get_state(c)[:nX].sum() for c in range(0, M, 16)])
    out[name] = {"chains_per_x_cta": CPC, "rounds_per_sweep_and_cta": r / (dN * M / CPC), "tfs_per_round": t / max(r, 1),
                 "flip_rounds_per_sweep_and_cta": f / (dN * M / CPC), "fp64_resolves_per_chain_and_sweep": dbg[10] / (dN * M),
                 "active_tfs_per_chain": float(act),
                 "edge_updates_per_s": M * dN * 3 * net.n_edges / (ms * 1e-3),
                 "ms_per_sweep_serialised": {k: ph[i] / max(nph, 1) for i, k in enumerate(("x", "t", "s"))}}
    print(name, json.dumps(out[name]), flush=True)
    G.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_flip_regimes.json"), "w"), indent=1)
