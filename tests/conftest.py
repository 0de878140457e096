import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def make_problem(NX=12, NY=60, AvgNTF=2.5, seed=5, n_graphs=3, hyper="cli", num_active_tfs=2, **over):
    """A seeded synthetic problem in the oracle's Problem form (same integer arrays the CUDA
    Sampler takes)."""
    import oracle
    from gbnet_b200 import synth
    net = synth.gen_network(NX=NX, NY=NY, AvgNTF=AvgNTF, num_active_tfs=num_active_tfs, seed=seed)
    hy = synth.cli_default_hyper() if hyper == "cli" else {}
    hy.update(over)
    pb = oracle.Problem(src=net.src, trg=net.trg, mor=net.mor, evid_trg=net.evid_trg, evid_val=net.evid_val,
                        n_graphs=n_graphs, **hy)
    return pb, net


def sampler_from_problem(pb, **over):
    """The CUDA Sampler for an oracle Problem."""
    import gbnet_b200
    kw = dict(src=pb.src, trg=pb.trg, mor=pb.mor, evid_trg=pb.evid_trg, evid_val=pb.evid_val,
              active_tf=pb.active_tf, sprior=pb.sprior, z_alpha=pb.z_alpha, z_beta=pb.z_beta,
              z0_alpha=pb.z0_alpha, z0_beta=pb.z0_beta, t_alpha=pb.t_alpha, t_beta=pb.t_beta,
              n_graphs=pb.n_graphs, noise_listen_children=pb.noise_listen_children,
              comp_yprob=pb.comp_yprob, const_params=pb.const_params, t_focus=pb.t_focus,
              t_lmargin=pb.t_lmargin, t_hmargin=pb.t_hmargin, zn_focus=pb.zn_focus,
              zn_lmargin=pb.zn_lmargin, zn_hmargin=pb.zn_hmargin, seed=pb.seed)
    kw.update(over)
    return gbnet_b200.Sampler(**kw)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@pytest.fixture(scope="session")
def have_gpu():
    import gbnet_b200
    return gbnet_b200.device_count() > 0


# Networks that fit one SM run on the persistent shared-memory kernel (sweep_small.cu) unless GBN_SMALL=2; the
# parity suites run every test on BOTH fast paths, so that the multi-kernel path (sweep_fast.cu: the one the
# BASELINE shapes C3 / C5 use) is held to the same small-network oracle comparisons - with its X phase on the
# single-chain shared-memory kernel (k_fast_x1, the default: "multi"), on the chain-pair kernel together with the first
# form of the S kernel (k_fast_x + k_fast_s: "pair"; every other mode runs the S phase on k_fast_s2) and with every TF
# of the single-chain kernel resolved in fp64 ("resolve").
_BOTH_PATHS = ("test_gpu_parity", "test_gpu_edgecases", "test_gpu_golden")


def pytest_generate_tests(metafunc):
    if metafunc.module.__name__.split(".")[-1] in _BOTH_PATHS and "fast_path" in metafunc.fixturenames:
        metafunc.parametrize("fast_path", ["auto", "multi", "pair", "resolve", "resolve8"], indirect=True)


@pytest.fixture(autouse=True)
def fast_path(request, monkeypatch):
    mode = getattr(request, "param", None)
    if mode in ("multi", "pair", "resolve", "resolve8"):
        monkeypatch.setenv("GBN_SMALL", "2")
    if mode == "pair":                       # X phase on the chain-pair kernel that gathers from the fp64 cache (k_fast_x)
        monkeypatch.setenv("GBN_X1", "2")
        monkeypatch.setenv("GBN_S_V", "1")   # ... and the S phase on k_fast_s (state copies and selects) instead of k_fast_s2
    if mode == "resolve":                    # single-chain X kernel with every TF resolved in fp64 by the whole CTA
        monkeypatch.setenv("GBN_X1", "3")
    if mode == "resolve8":                   # ... every eighth TF (several warps per resolved TF)
        monkeypatch.setenv("GBN_X1", "4")
    return mode
