"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both sweep kernels, listening
and non-listening T, injected draws, batched datasets, moments, checkpoint."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gbnet_b200  # noqa: E402
from gbnet_b200 import synth  # noqa: E402

net = synth.gen_network(NX=37, NY=333, AvgNTF=3.0, num_active_tfs=3, seed=2)
hy = synth.cli_default_hyper()
evs = np.stack([net.evid_val, synth.redraw_evidence(net, seed=5)[0]])
for kernel in (gbnet_b200.KERNEL_FAST, gbnet_b200.KERNEL_STRICT):
    for listen in (False, True):
        h = dict(hy, noise_listen_children=listen)
        G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, evs, n_graphs=3, seed=1, kernel=kernel, **h)
        G.sample_n(4)
        f, b, e = G.export_draws(0, 0, 2)
        G.inject(np.stack([f] * G.n_chains()), np.stack([b] * G.n_chains()), np.stack([e] * G.n_chains()))
        G.sample_n(3)
        G.sample(40, 10)
        G.gelman_rubin(1), G.posterior_means("S", 1), G.posterior_sdevs("T"), G.node_moments(2)
        G.all_conditionals(1), G.target_likelihoods(4), G.joint_logprob(5)
        blob = G.checkpoint()
        G.restore(blob)
        G.set_state(0, G.get_state(3))
        G.set_evidence(net.evid_trg, evs)
        G.sample_n(2)
        print("kernel", kernel, "listen", listen, "max R-hat", G.max_gelman_rubin(), flush=True)
        G.close()
print("sanitize case done")
