"""Callers of the hot path, mirroring reference gbnet/utils.py.

``get_model`` maps ``s_leniency`` to the 3x3 S prior and builds a ``ModelORNOR`` exactly as
utils.py:70-87 does; ``sample_model`` is the burn-in + Gelman-Rubin protocol of utils.py:89-144 that
defines "time to R-hat < 1.1" (10 x sample(20), burn_stats, then sample(20) until the max R-hat is
<= 1.1 or max_its is reached) and returns the same result table.  ``frames_from_network`` turns a
synthetic network (gbnet_b200.synth) into the ents / rels / evid frames gbnet.tools.genData
produces (tools.py:60-85).
"""
from __future__ import annotations

from datetime import datetime

import numpy as np
import pandas as pd

from .model import ModelORNOR


# columns of the screen: (name, sign of the edge or None, predicate on the evidence value)
_COUNT_COLUMNS = (("trg_ydeg", None, lambda v: v != 0), ("trg_ndeg", None, lambda v: v == 0),
                  ("morp_degp", 1, lambda v: v == 1), ("morp_degn", 1, lambda v: v == -1),
                  ("morn_degp", -1, lambda v: v == 1), ("morn_degn", -1, lambda v: v == -1))


def _fisher_greater(a, b, c, d):
    """One-sided ("greater") Fisher p-value of the table [[a, b], [c, d]] for whole columns at once:
    P(X >= a) with X ~ Hypergeom(a + b + c + d, a + b, a + c)."""
    from scipy.stats import hypergeom
    a, b, c, d = (np.asarray(v, np.int64) for v in (a, b, c, d))
    total = a + b + c + d
    # an all-zero table (a TF none of whose signed edges reaches a DE target) has p = 1 in fisher_exact;
    # the survival function of an empty population is NaN
    with np.errstate(invalid="ignore"):
        p = hypergeom.sf(a - 1, np.maximum(total, 1), a + b, a + c)
    return np.where(total == 0, 1.0, p)


def get_tests(evid, rels):
    """Per-TF enrichment / mode-of-regulation screen, the table reference gbnet/utils.py:10-68 builds:
    for every source of ``rels`` the number of its targets that are / are not differentially expressed
    and the 2x2 counts of edge sign against DE sign; Fisher's exact test ("greater") of DE enrichment and
    of sign agreement; their product rho combined as rho (1 - log rho); the target total and a signed
    score.  The reference calls scipy.stats.fisher_exact per TF and table; the p-value is the
    hypergeometric survival function, evaluated here for all TFs at once (checked against fisher_exact
    in tests/test_host_logic.py)."""
    values = evid.reset_index().set_index("uid")["val"]
    # left join on the target: NaN where the target has no evidence
    edges = rels.merge(values.to_frame(), how="left", left_on="trguid", right_index=True)
    n_de = int((values != 0).sum())
    n_not_de = int((values == 0).sum())

    counts = {}
    for name, sign, pred in _COUNT_COLUMNS:
        keep = pred(edges["val"]) & edges["val"].notna() if name == "trg_ndeg" else pred(edges["val"])
        if sign is not None:
            keep &= edges["type"] == sign
        counts[name] = edges.loc[keep].groupby("srcuid")["trguid"].count()
    # the reference's "val != 0" also keeps targets without evidence (NaN != 0): same here
    table = pd.DataFrame(counts).fillna(0).astype(int)
    table = table[[c[0] for c in _COUNT_COLUMNS]]
    table.index.name = "srcuid"

    table["enrichment"] = _fisher_greater(table.trg_ydeg, n_de - table.trg_ydeg, table.trg_ndeg, n_not_de - table.trg_ndeg)
    table["mor_binary"] = _fisher_greater(table.morp_degp, table.morn_degp, table.morp_degn, table.morn_degn)
    rho = (table.enrichment * table.mor_binary).to_numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        table["combined"] = np.where(rho > 0, rho * (1 - np.log(np.where(rho > 0, rho, 1.))), rho)
    table["trg_tot"] = table.trg_ndeg + table.trg_ydeg
    agree = table.morp_degp + table.morn_degn
    disagree = table.morp_degn + table.morn_degp
    table["score"] = (agree - disagree) / table.trg_ydeg
    table["gt_act"] = False
    return table


def _tests_from_counts(index, counts, n_de, n_not_de):
    """The table of ``get_tests`` from the six per-TF edge counts (columns of ``_COUNT_COLUMNS``)."""
    table = pd.DataFrame(np.asarray(counts, np.int64), index=pd.Index(index, name="srcuid"),
                         columns=[c[0] for c in _COUNT_COLUMNS])
    table["enrichment"] = _fisher_greater(table.trg_ydeg, n_de - table.trg_ydeg, table.trg_ndeg, n_not_de - table.trg_ndeg)
    table["mor_binary"] = _fisher_greater(table.morp_degp, table.morn_degp, table.morp_degn, table.morn_degn)
    rho = (table.enrichment * table.mor_binary).to_numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        table["combined"] = np.where(rho > 0, rho * (1 - np.log(np.where(rho > 0, rho, 1.))), rho)
    table["trg_tot"] = table.trg_ndeg + table.trg_ydeg
    agree = table.morp_degp + table.morn_degn
    disagree = table.morp_degn + table.morn_degp
    table["score"] = (agree - disagree) / table.trg_ydeg
    table["gt_act"] = False
    return table


def get_tests_batched(evids, rels, device=0):
    """``get_tests`` (reference gbnet/utils.py:10-68) for MANY contrasts that share ``rels``: one list of result
    tables, one per evidence frame.  The per-TF contingency counts - the reference's six pandas group-bys per
    contrast - are ONE segmented reduction over the by-TF edge list on the GPU (csrc/fisher.cu); the Fisher
    p-values are evaluated from them for all TFs at once as in ``get_tests``."""
    from .engine import EVID_ABSENT, fisher_counts
    frames = [e.reset_index().set_index("uid")["val"] for e in evids]
    uids = np.unique(np.concatenate([f.index.to_numpy() for f in frames])) if frames else np.zeros(0, np.int64)
    vals = np.full((len(frames), uids.size), EVID_ABSENT, np.int8)
    for d, f in enumerate(frames):
        f = f[~f.index.duplicated(keep="first")]
        vals[d, np.searchsorted(uids, f.index.to_numpy())] = f.to_numpy().astype(np.int8)
    ids, counts = fisher_counts(rels["srcuid"].to_numpy(), rels["trguid"].to_numpy(), rels["type"].to_numpy(),
                                uids, vals, device=device)
    out = []
    for d, f in enumerate(frames):
        out.append(_tests_from_counts(ids.astype(np.int64), counts[d], int((f != 0).sum()), int((f == 0).sum())))
    return out


def s_prior_from_leniency(s_leniency):
    """The 3x3 prior of the edge mode S given the annotated sign (rows: sign -1, 0, +1; columns: S = 0, 1, 2)
    that reference gbnet/utils.py:78-81 derives from one leniency number."""
    if not s_leniency <= 0.5:
        raise AssertionError("s_leniency must not exceed 0.5")
    keep, l = 1.0 - s_leniency, s_leniency
    return [[keep, 0.9 * l, 0.1 * l],
            [0.5 * l, keep, 0.5 * l],
            [0.1 * l, 0.9 * l, keep]]


_MODEL_DEFAULTS = dict(const_params=False, noise_listen_children=False, comp_yprob=False,
                       t_focus=2., t_lmargin=2., t_hmargin=2., zn_focus=2., zn_lmargin=8., zn_hmargin=2.,
                       z_alpha=198, z_beta=2, z0_alpha=198, z0_beta=2, t_alpha=2, t_beta=2, n_graphs=3)


def get_model(ents, rels, evid, s_leniency=0.1, **options):
    """A ``ModelORNOR`` with the defaults of reference gbnet/utils.py:70-87 (which differ from the class's own:
    z priors 198 / 2, non-listening T nodes); ``s_leniency`` becomes the S prior.  Further keywords
    (``seed``, ``device`` ...) go to the model unchanged."""
    settings = dict(_MODEL_DEFAULTS)
    settings.update(options)
    return ModelORNOR(ents, rels, evid, sprior=s_prior_from_leniency(s_leniency), **settings)


def run_protocol(model, max_its=100, burn_its=10, verbosity=0, batch=20):
    """The sampling schedule of the reference's sample_model (utils.py:93-135), which defines "time to
    R-hat < 1.1": ``burn_its`` batches, statistics dropped, then batches until the largest Gelman-Rubin
    statistic is at most 1.1 or ``max_its`` batches (burn-in included) have run.
    Returns (batches run, last max R-hat)."""
    def log(tag, value):
        if verbosity >= 1:
            print(f"[{datetime.now()}] {tag}mgr={value: 10.4f}", flush=True)

    fused = getattr(model, "run_protocol", None)
    if fused is not None and verbosity < 1:
        # one library call (gbn_model_run_protocol): no Python round trip, no stream sync during burn-in
        return fused(burn_its=burn_its, max_its=max_its, N=batch)
    done = 0
    for _ in range(burn_its):
        model.sample(batch)
        done += 1
        if verbosity >= 1:
            log("burn-in ", model.get_max_gelman_rubin())
    model.burn_stats()
    rhat = float("inf")
    while rhat > 1.1:
        model.sample(batch)
        done += 1
        rhat = model.get_max_gelman_rubin()
        log("", rhat)
        if done >= max_its:
            break
    return done, rhat


def _posterior_column(model, var_name, keep_idx, column):
    """Series ``column`` indexed by symbol from a posterior-mean getter (tuples id, symbol, outcome, value)."""
    rows = [(sym, val) for _, sym, idx, val in model.get_posterior_means(var_name) if idx == keep_idx]
    return pd.Series(dict(rows), name=column, dtype=float).rename_axis("symbol")


def sample_model(model, ents, tests, max_its=100, burn_its=10, verbosity=0):
    """Run the protocol and return the reference's result table (utils.py:137-144): the screen ``tests``
    joined with the posterior mean of X = 1 and of T per TF, most active TF first."""
    done, _ = run_protocol(model, max_its=max_its, burn_its=burn_its, verbosity=verbosity)
    print(f"Completed {done} iterations.")
    post = pd.concat([_posterior_column(model, "X", "1", "X"), _posterior_column(model, "T", "0", "T")],
                     axis=1, join="inner")
    proteins = ents.loc[ents.type == "Protein"].set_index("name")["uid"]
    post["uid"] = proteins.loc[post.index.values].to_numpy()
    by_uid = post.reset_index().set_index("uid")
    return tests.join(by_uid, how="inner").sort_values("X", ascending=False)


def frames_from_network(net):
    """ents / rels / evid / tests frames in the shape gbnet.tools.genData returns (tools.py:60-85)
    for a gbnet_b200.synth.SynthNetwork."""
    NX, NY = net.NX, net.NY
    ents = pd.DataFrame({"uid": np.arange(NX + NY), "name": [f"{i:04d}" for i in range(NX + NY)],
                         "type": ['Protein'] * NX + ['mRNA'] * NY})
    rels = pd.DataFrame({"srcuid": net.src.astype(np.int64), "trguid": net.trg.astype(np.int64),
                         "type": net.mor.astype(np.int64)})
    evid = pd.DataFrame({"uid": net.evid_trg.astype(np.int64), "val": net.evid_val.astype(np.int64)})
    tests = get_tests(evid, rels)
    tests["gt_act"] = tests.index.isin(net.gt_active.astype(np.int64))                 # tools.py:84-85
    return ents, rels, evid, tests
