"""Edge cases of the hot path on the GPU, both kernels against the oracle (and the compiled reference
where oracle/_ref travelled with the snapshot): smallest networks, ragged degrees, a target with
hundreds of parents, a hub TF whose children push exp(sum log L) to underflow so the prior fallback
of Multinomial.cpp:78-85 decides the draw, evidence quirks, and the long-run C1 case against the
reference's own C++."""
import numpy as np
import pytest

import oracle
from conftest import make_problem, rel_err, sampler_from_problem
from test_gpu_parity import _assert_state_equal, _run_trajectory

pytestmark = pytest.mark.gpu

from gbnet_b200 import KERNEL_FAST, KERNEL_STRICT, synth  # noqa: E402


def _problem(src, trg, mor, evid, n_graphs=3, seed=77, **hy):
    h = dict(synth.cli_default_hyper())
    h.update(hy)
    evid_trg = np.array(sorted(evid), np.uint32)
    evid_val = np.array([evid[k] for k in sorted(evid)], np.int32)
    return oracle.Problem(src=np.asarray(src, np.uint32), trg=np.asarray(trg, np.uint32), mor=np.asarray(mor, np.int32),
                          evid_trg=evid_trg, evid_val=evid_val, n_graphs=n_graphs, seed=seed, **h)


@pytest.mark.parametrize("kernel", [KERNEL_STRICT, KERNEL_FAST])
def test_smallest_networks(kernel):
    # one edge; two TFs on one target; one TF on two targets; evidence that names a target not in the network
    for src, trg, mor, evid in [([0], [10], [1], {10: 1}),
                                ([0, 1], [10, 10], [1, -1], {10: -1, 99: 1}),
                                ([0, 0], [10, 11], [-1, 0], {10: 0, 11: 1}),
                                ([3, 1, 2, 1], [7, 7, 5, 6], [1, 1, -1, 0], {5: 1})]:
        pb = _problem(src, trg, mor, evid)
        P, G = _run_trajectory(pb, kernel, 30, injected=False, chunks=(10,))
        a, b = G.all_conditionals(0), P.all_conditionals(0)
        m = ~np.isnan(b)
        assert rel_err(a[m], b[m]).max() < 1e-12
        assert rel_err(G.joint_logprob(1), P.joint_logprob(1)) < 1e-12


@pytest.mark.parametrize("kernel", [KERNEL_STRICT, KERNEL_FAST])
def test_wide_target_and_ragged_degrees(kernel):
    """One target with 300 parents (group width 300, the other 31 lanes of its warp padded), targets
    with a single parent, TFs with a single child, edge input order scrambled."""
    rng = np.random.default_rng(5)
    nX = 300
    src = list(range(nX)) + list(rng.integers(0, nX, 120))
    trg = [1000] * nX + list(1001 + np.arange(120) // 2)
    mor = list(rng.choice([-1, 0, 1], nX + 120))
    perm = rng.permutation(len(src))
    src, trg, mor = np.array(src)[perm], np.array(trg)[perm], np.array(mor)[perm]
    keep = np.unique(np.stack([src, trg]), axis=1, return_index=True)[1]          # no duplicate (TF, target) pairs
    src, trg, mor = src[np.sort(keep)], trg[np.sort(keep)], mor[np.sort(keep)]
    evid = {int(t): int(rng.choice([-1, 0, 1])) for t in np.unique(trg)}
    pb = _problem(src, trg, mor, evid, n_graphs=2)
    _run_trajectory(pb, kernel, 12 if kernel == KERNEL_STRICT else 40, injected=False, chunks=(4,))


@pytest.mark.parametrize("kernel", [KERNEL_STRICT, KERNEL_FAST])
def test_hub_underflow_prior_fallback(kernel):
    """A hub TF with 4,000 DE children whose signs contradict each other: both outcome likelihoods
    underflow (sum log L ~ -10^4), so Multinomial::compute_outcome_likelihood falls back to the prior
    (Multinomial.cpp:78-85).  The fast kernel tracks the product as mantissa x 2^e and must take the
    same branch."""
    rng = np.random.default_rng(9)
    n_child = 4000
    src = [0] * n_child + list(rng.integers(1, 6, 200))
    trg = list(100 + np.arange(n_child)) + list(100 + rng.integers(0, n_child, 200))
    mor = [1] * n_child + list(rng.choice([-1, 1], 200))
    keep = np.sort(np.unique(np.stack([src, trg]), axis=1, return_index=True)[1])
    src, trg, mor = np.array(src)[keep], np.array(trg)[keep], np.array(mor)[keep]
    evid = {int(t): int(v) for t, v in zip(100 + np.arange(n_child), rng.choice([-1, 1], n_child))}
    pb = _problem(src, trg, mor, evid, n_graphs=2)
    P = oracle.PortModel(pb)
    c = P.all_conditionals(0)
    assert np.all(c[0, :2] < -746.)                    # exp() of both X_0 candidates underflows to zero
    P2, G = _run_trajectory(pb, kernel, 6 if kernel == KERNEL_STRICT else 25, injected=False, chunks=(3,))
    # with the prior deciding, X_0 is active in about 1 % of the sweeps
    assert G.posterior_means("X")[1] < 0.3


@pytest.mark.parametrize("n_child", [350, 360])
def test_hub_in_the_subnormal_band(n_child):
    """A hub TF whose better outcome likelihood hovers around the bottom of the double range (sum log L between
    -709 and -790 over the sweeps; exp() goes subnormal below -708.4 and to zero below -745.1): the draw switches
    between "the subnormal outcome wins", and "both are zero: the prior decides" (Multinomial.cpp:78-85) from sweep to
    sweep.  The X kernels must take the exact product of the children's likelihoods here (k_fast_x: gathered
    likelihoods; k_fast_x1: the fp64 resolve with the exact product - the bound bytes cannot decide inside the band)
    and reproduce the oracle's chain bit for bit."""
    rng = np.random.default_rng(9)
    src = [0] * n_child + list(rng.integers(1, 6, 60))
    trg = list(100 + np.arange(n_child)) + list(100 + rng.integers(0, n_child, 60))
    mor = [1] * n_child + list(rng.choice([-1, 1], 60))
    keep = np.sort(np.unique(np.stack([src, trg]), axis=1, return_index=True)[1])
    src, trg, mor = np.array(src)[keep], np.array(trg)[keep], np.array(mor)[keep]
    evid = {int(t): int(v) for t, v in zip(100 + np.arange(n_child), rng.choice([-1, 1], n_child))}
    pb = _problem(src, trg, mor, evid, n_graphs=2)
    P = oracle.PortModel(pb)
    seen = []
    for _ in range(12):
        seen.append(P.all_conditionals(0)[0, :2].max())
        P.sample_n(1)
    assert min(seen) < -745.2 and -745. < max(seen) < -708.4        # both sides of the underflow threshold, never normal
    _run_trajectory(pb, KERNEL_FAST, 40, injected=False, chunks=(5,))


def test_unsorted_targets_are_the_same_chain(monkeypatch):
    """GBN_SORT_TARGETS=2 keeps the device targets in reference order: group widths are no longer descending, so a
    CTA of the S kernel stages every table row (DevNet.groups_desc = 0) and wide and narrow groups share CTAs.  The
    chain does not depend on the device order (the Philox keys use reference indices)."""
    monkeypatch.setenv("GBN_SORT_TARGETS", "2")
    rng = np.random.default_rng(11)
    nX, nT = 40, 400
    deg = rng.integers(1, 30, nT)
    deg[::37] = 70                                                 # a few wide targets scattered through the order
    src = np.concatenate([rng.choice(nX, min(d, nX), replace=False) for d in deg])
    trg = np.concatenate([np.full(min(d, nX), 500 + i) for i, d in enumerate(deg)])
    mor = rng.choice([-1, 0, 1], len(src))
    evid = {int(t): int(rng.choice([-1, 0, 1])) for t in np.unique(trg)}
    pb = _problem(src, trg, mor, evid, n_graphs=2)
    _run_trajectory(pb, KERNEL_FAST, 25, injected=False, chunks=(5,))


@pytest.mark.parametrize("sprior", [
    (1., 0., 0., 0., 1., 0., 0., 0., 1.),                            # certain priors: zero entries, nothing tiny
    (1e-250, 1. - 1e-250, 0., 0.5, 0.25, 0.25, 0., 1. - 1e-250, 1e-250),   # tiny entries: not certified, runs on k_fast_s
    (0., 0., 0., 0.3, 0.4, 0.3, 0.2, 0.4, 0.4),                      # an all-zero row: the fallback of Multinomial.cpp:78-85
])
def test_degenerate_prior_rows(sprior):
    """Prior rows with zeros, with entries so small that prior x likelihood could vanish, and an all-zero row: the
    model layer certifies rows for k_fast_s2 (which has no all-zero fallback) or keeps the S phase on k_fast_s; either
    way the chain is the oracle's."""
    pb, _ = make_problem(NX=15, NY=120, AvgNTF=3.0, seed=8, n_graphs=3, sprior=sprior)
    for injected in (False, True):
        _run_trajectory(pb, KERNEL_FAST, 30, injected=injected, chunks=(7,))


def test_evidence_quirks():
    # all targets unchanged; all evidence for targets outside the network (every y defaults to "not DE")
    pb0 = _problem([0, 1, 1], [5, 5, 6], [1, -1, 1], {5: 0, 6: 0})
    pb1 = _problem([0, 1, 1], [5, 5, 6], [1, -1, 1], {50: 1, 60: -1})
    P0, G0 = _run_trajectory(pb0, KERNEL_FAST, 20, injected=False, chunks=(20,))
    P1, G1 = _run_trajectory(pb1, KERNEL_FAST, 20, injected=False, chunks=(20,))
    for c in range(3):
        assert np.array_equal(G0.get_state(c), G1.get_state(c))
    # comp_yprob counts the evidence entries, including those outside the network (GraphORNOR.cpp:113-125)
    pb2 = _problem([0, 1, 1], [5, 5, 6], [1, -1, 1], {5: 1, 6: -1, 70: 1}, comp_yprob=True)
    _run_trajectory(pb2, KERNEL_FAST, 20, injected=False, chunks=(20,))
    _run_trajectory(pb2, KERNEL_STRICT, 20, injected=False, chunks=(20,))
    # sparse target ids (no dense id table on the host), repeated evidence keys (the first one wins) and
    # replacing the evidence of a live model
    big = 4_000_000_000
    pb3 = _problem([0, 1, 1], [7, big, big - 9], [1, -1, 1], {7: 1, big: -1, 123456789: 1}, comp_yprob=True)
    P3, G3 = _run_trajectory(pb3, KERNEL_FAST, 20, injected=False, chunks=(20,))
    trg = np.array([big, 7, big, 99, big - 9], np.uint32)
    val = np.array([1, -1, -1, 1, 0], np.int32)
    G3.set_evidence(trg, val)
    pb4 = _problem([0, 1, 1], [7, big, big - 9], [1, -1, 1], {big: 1, 7: -1, 99: 1, big - 9: 0}, comp_yprob=True)
    P4 = oracle.PortModel(pb4)
    G3.sample_n(15)
    P4.sample_n(15)
    for c in range(pb4.n_graphs):
        _assert_state_equal(pb4, P4, G3.get_state(c), P4.get_state(c), f"replaced evidence, chain {c}")


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("listen", [False, True])
def test_c1_against_the_reference_cpp(listen):
    """BASELINE config C1 (50 TFs / 2,000 targets / ~5,000 edges, 3 chains): the CUDA fast path and the
    reference's own C++ (compiled unmodified, oracle/_ref) driven by the same injected draw stream for
    60 sweeps: identical discrete states, T within 1e-9, same Gelman-Rubin and posterior summaries."""
    from conftest import make_problem
    pb, _ = make_problem(NX=50, NY=2000, AvgNTF=2.5, seed=1, n_graphs=3, num_active_tfs=3, noise_listen_children=listen)
    P = oracle.PortModel(pb)                           # source of the draw stream and of the initial state
    R = oracle.RefModel(pb)
    G = sampler_from_problem(pb, kernel=KERNEL_FAST)
    n = 60
    draws = [P.export_draws(c, 0, n) for c in range(3)]
    for c in range(3):
        R.set_state(c, G.get_state(c))
        R.inject(c, *draws[c])
    G.inject(np.stack([d[0] for d in draws]), np.stack([d[1] for d in draws]), np.stack([d[2] for d in draws]))
    for _ in range(3):
        R.sample_n(n // 3)
        G.sample_n(n // 3)
        for c in range(3):
            _assert_state_equal(pb, P, G.get_state(c), R.get_state(c), f"chain {c}")
    gg, gr = G.gelman_rubin(), R.gelman_rubin()
    fin = np.isfinite(gr)
    assert np.array_equal(np.isfinite(gg), fin) and rel_err(gg[fin], gr[fin]).max() < 1e-9
    for k in "XTS":
        assert np.abs(G.posterior_means(k) - R.posterior_means(k)).max() < 1e-9
        assert np.abs(G.posterior_sdevs(k) - R.posterior_sdevs(k)).max() < 1e-9
    a, b = G.all_conditionals(0), R.all_conditionals(0)
    m = ~np.isnan(b)
    assert rel_err(a[m], b[m]).max() < 1e-12
