"""Generate the golden fixtures from the REFERENCE ITSELF (its own C++, compiled unmodified into
oracle/_ref/libgbnet_ref.so by `make -C oracle ref`; needs /root/reference, so it only runs in the
build container).  The fixtures are what pins the oracle and the CUDA path where the reference
is absent:

    python tests/golden/make_golden.py        ->  tests/golden/ref_<case>.npz

Each case stores the integer problem, hyper-parameters, a state, and the reference's answers:
conditional log-likelihoods of every candidate of every node, target likelihoods, joint
log-probability, and - for a trajectory driven by an injected draw stream - the states after
every sweep block, per-chain running moments, Gelman-Rubin and posterior summaries.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from gbnet_b200 import synth  # noqa: E402

CASES = {
    # name: (network kwargs, hyper kwargs)
    "cli": (dict(NX=14, NY=70, AvgNTF=3.0, num_active_tfs=2, seed=101), dict(synth.cli_default_hyper())),
    "pyx": (dict(NX=10, NY=45, AvgNTF=2.5, num_active_tfs=2, seed=102), dict()),       # ctor defaults: T listens
    "const_yprob": (dict(NX=12, NY=60, AvgNTF=3.0, num_active_tfs=2, seed=103),
                    dict(synth.cli_default_hyper(), const_params=True, comp_yprob=True)),
}
N_SWEEPS = 12
BLOCK = 4
N_GRAPHS = 3


def main():
    if not oracle.have_ref():
        oracle.build()
    assert oracle.have_ref(), "oracle/_ref/libgbnet_ref.so is required (needs /root/reference)"
    for name, (nk, hy) in CASES.items():
        net = synth.gen_network(**nk)
        pb = oracle.Problem(src=net.src, trg=net.trg, mor=net.mor, evid_trg=net.evid_trg, evid_val=net.evid_val,
                            n_graphs=N_GRAPHS, seed=7, **hy)
        if name == "cli":
            pb.active_tf = np.array([2], np.uint32)
        R = oracle.RefModel(pb)
        P = oracle.PortModel(pb)          # only used as the source of a draw stream and initial state
        rng = np.random.default_rng(5)
        nX = P.n_tf()
        nT = 0 if pb.const_params else nX
        out = dict(src=pb.src, trg=pb.trg, mor=pb.mor, evid_trg=pb.evid_trg, evid_val=pb.evid_val,
                   active_tf=pb.active_tf, n_graphs=N_GRAPHS, hyper_keys=np.array(sorted(hy.keys())),
                   hyper_vals=np.array([repr(hy[k]) for k in sorted(hy.keys())]))
        # --- static quantities on a random state
        st = P.get_state(0)
        st[:nX] = rng.integers(0, 2, nX)
        st[nX:nX + nT] = rng.uniform(0.05, 0.95, nT)
        st[nX + nT:] = rng.integers(0, 3, st.size - nX - nT)
        R.set_state(0, st)
        out["state"] = st
        out["conditionals"] = R.all_conditionals(0)
        out["target_likelihoods"] = R.target_likelihoods(0)
        out["joint"] = np.float64(R.joint_logprob(0))
        # --- trajectory under an injected stream, all chains
        init = [P.get_state(c) for c in range(N_GRAPHS)]
        draws = [P.export_draws(c, 0, N_SWEEPS) for c in range(N_GRAPHS)]
        for c in range(N_GRAPHS):
            R.set_state(c, init[c])
            R.inject(c, *draws[c])
        out["init"] = np.stack(init)
        out["flat_u"] = np.stack([d[0] for d in draws])
        out["beta_v"] = np.stack([d[1] for d in draws])
        out["exp_v"] = np.stack([d[2] for d in draws])
        states = []
        for _ in range(N_SWEEPS // BLOCK):
            R.sample_n(BLOCK)
            states.append(np.stack([R.get_state(c) for c in range(N_GRAPHS)]))
        out["states"] = np.stack(states)                    # [blocks, chains, n_nodes]
        n_nodes = R.n_nodes()
        kinds = [R.node_info(n)[0] for n in range(n_nodes)]
        nst = [R.node_info(n)[2] for n in range(n_nodes)]
        means, vars_ = [], []
        for c in range(N_GRAPHS):
            m, v = [], []
            for n in range(n_nodes):
                for s in range(nst[n]):
                    m.append(R.node_mean(c, n, s))
                    v.append(R.node_variance(c, n, s))
            means.append(m)
            vars_.append(v)
        out["node_kinds"] = np.array(kinds)
        out["node_means"] = np.array(means)
        out["node_vars"] = np.array(vars_)
        out["gelman_rubin"] = R.gelman_rubin()
        out["max_gelman_rubin"] = np.float64(R.max_gelman_rubin())
        for k in ("X", "T", "S"):
            out["post_mean_" + k] = R.posterior_means(k)
            out["post_sd_" + k] = R.posterior_sdevs(k)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path), "bytes; nodes", n_nodes)


def main_sim():
    """Simulation mode (empty evidence, GraphORNOR.cpp:28): the reference's forward simulator of H / Y."""
    for name, listen in (("sim", False), ("sim_listen", True)):
        net = synth.gen_network(NX=12, NY=70, AvgNTF=3.0, num_active_tfs=2, seed=104)
        hy = dict(synth.cli_default_hyper(), noise_listen_children=listen)
        pb = oracle.Problem(src=net.src, trg=net.trg, mor=net.mor, evid_trg=np.zeros(0, np.uint32),
                            evid_val=np.zeros(0, np.int32), n_graphs=N_GRAPHS, seed=7, **hy)
        pb.active_tf = net.gt_active
        R = oracle.RefModel(pb)
        n_nodes = R.n_nodes()
        nX = int(np.unique(net.src).size)
        nH = (n_nodes - nX) // 2
        rng = np.random.default_rng(11)
        out = dict(src=pb.src, trg=pb.trg, mor=pb.mor, evid_trg=pb.evid_trg, evid_val=pb.evid_val,
                   active_tf=pb.active_tf, n_graphs=N_GRAPHS, hyper_keys=np.array(sorted(hy.keys())),
                   hyper_vals=np.array([repr(hy[k]) for k in sorted(hy.keys())]))
        st = R.get_state(0)
        st[:2 * nH] = rng.integers(0, 3, 2 * nH)
        st[2 * nH:] = rng.uniform(0.05, 0.95, nX)
        R.set_state(0, st)
        out["state"] = st
        out["conditionals"] = R.all_conditionals(0)
        init = np.stack([R.get_state(c) for c in (1, 1, 1)])            # fresh state (H = Y = 1, T at its prior mean)
        flat = rng.random((N_GRAPHS, N_SWEEPS, 2 * nH))
        beta = np.stack([[[rng.beta(*R.t_prior(2 * nH + k)) for k in range(nX)] for _ in range(N_SWEEPS)]
                         for _ in range(N_GRAPHS)])
        expo = rng.exponential(1.0, (N_GRAPHS, N_SWEEPS, nX))
        for c in range(N_GRAPHS):
            R.set_state(c, init[c])
            R.inject(c, flat[c], beta[c], expo[c])
        out["init"], out["flat_u"], out["beta_v"], out["exp_v"] = init, flat, beta, expo
        states = []
        for _ in range(N_SWEEPS // BLOCK):
            R.sample_n(BLOCK)
            states.append(np.stack([R.get_state(c) for c in range(N_GRAPHS)]))
        out["states"] = np.stack(states)
        nst = [R.node_info(n)[2] for n in range(n_nodes)]
        out["node_kinds"] = np.array([R.node_info(n)[0] for n in range(n_nodes)])
        out["node_means"] = np.array([[R.node_mean(c, n, s) for n in range(n_nodes) for s in range(nst[n])] for c in range(N_GRAPHS)])
        out["node_vars"] = np.array([[R.node_variance(c, n, s) for n in range(n_nodes) for s in range(nst[n])] for c in range(N_GRAPHS)])
        out["gelman_rubin"] = R.gelman_rubin()
        for k in ("H", "Y", "T", "X", "S"):
            out["post_mean_" + k] = R.posterior_means(k)
            out["post_sd_" + k] = R.posterior_sdevs(k)
        path = os.path.join(HERE, f"refsim_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path), "bytes; nodes", n_nodes)


if __name__ == "__main__":
    main()
    main_sim()
