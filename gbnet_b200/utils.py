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


def get_tests(evid, rels):
    """Per-TF enrichment / mode-of-regulation screen (reference gbnet/utils.py:10-68): Fisher's exact
    test (alternative "greater") of DE among the TF's targets and of sign agreement, their Brown-style
    combination rho (1 - log rho), target count and a signed score.  The reference calls
    scipy.stats.fisher_exact once per TF and table; the one-sided p-value is the hypergeometric
    survival function, evaluated here for all TFs at once."""
    from scipy.stats import hypergeom

    evid = evid.reset_index().set_index('uid')
    tmp = rels.merge(evid.loc[:, ['val']], how='left', left_on='trguid', right_index=True)
    all_degp = int((evid.val == 1).sum())
    all_degn = int((evid.val == -1).sum())
    all_ydeg = all_degp + all_degn
    all_ndeg = int((evid.val == 0).sum())

    def count(mask, name):
        return tmp.loc[mask, ['srcuid', 'trguid']].groupby('srcuid').count()['trguid'].rename(name)

    cols = [count(tmp.val != 0, 'trg_ydeg'), count(tmp.val == 0, 'trg_ndeg'),
            count((tmp.type == 1) & (tmp.val == 1), 'morp_degp'), count((tmp.type == 1) & (tmp.val == -1), 'morp_degn'),
            count((tmp.type == -1) & (tmp.val == 1), 'morn_degp'), count((tmp.type == -1) & (tmp.val == -1), 'morn_degn')]
    df = pd.DataFrame(cols).T.fillna(0).astype(int)
    df.index.name = 'srcuid'

    def fisher_greater(a, b, c, d):
        # table [[a, b], [c, d]]: P(X >= a), X ~ Hypergeom(total, a + b, a + c)
        a, b, c, d = (np.asarray(v, np.int64) for v in (a, b, c, d))
        return hypergeom.sf(a - 1, a + b + c + d, a + b, a + c)

    df['enrichment'] = fisher_greater(df.trg_ydeg, all_ydeg - df.trg_ydeg, df.trg_ndeg, all_ndeg - df.trg_ndeg)
    df['mor_binary'] = fisher_greater(df.morp_degp, df.morn_degp, df.morp_degn, df.morn_degn)
    rho = df.enrichment * df.mor_binary
    df['combined'] = rho.apply(lambda r: r * (1 - np.log(r)) if r > 0 else r)
    df['trg_tot'] = df.trg_ndeg + df.trg_ydeg
    df['score'] = (df.morp_degp + df.morn_degn - df.morp_degn - df.morn_degp) / df.trg_ydeg
    df['gt_act'] = False
    return df


def get_model(ents, rels, evid,
              const_params=False, noise_listen_children=False, comp_yprob=False,
              t_focus=2., t_lmargin=2., t_hmargin=2., zn_focus=2., zn_lmargin=8., zn_hmargin=2.,
              z_alpha=198, z_beta=2, z0_alpha=198, z0_beta=2, t_alpha=2, t_beta=2,
              s_leniency=0.1, n_graphs=3, **extra):
    assert s_leniency <= 0.5
    sprior = [[1.0 - s_leniency, 0.9 * s_leniency, 0.1 * s_leniency],
              [0.5 * s_leniency, 1.0 - s_leniency, 0.5 * s_leniency],
              [0.1 * s_leniency, 0.9 * s_leniency, 1.0 - s_leniency]]
    return ModelORNOR(ents, rels, evid,
                      n_graphs=n_graphs, sprior=sprior,
                      z_alpha=z_alpha, z_beta=z_beta, z0_alpha=z0_alpha, z0_beta=z0_beta, t_alpha=t_alpha, t_beta=t_beta,
                      t_focus=t_focus, t_lmargin=t_lmargin, t_hmargin=t_hmargin, zn_focus=zn_focus,
                      zn_lmargin=zn_lmargin, zn_hmargin=zn_hmargin,
                      noise_listen_children=noise_listen_children, comp_yprob=comp_yprob, const_params=const_params,
                      **extra)


def run_protocol(model, max_its=100, burn_its=10, verbosity=0):
    """The sampling schedule of sample_model (utils.py:93-135); returns (iterations, max R-hat)."""
    it = 0
    for _ in range(burn_its):
        it += 1
        model.sample(20)
        if verbosity >= 1:
            print(f"[{datetime.now()}] burn-in mgr={model.get_max_gelman_rubin(): 10.4f}", flush=True)
    model.burn_stats()
    mgr = float('inf')
    while mgr > 1.1:
        it += 1
        model.sample(20)
        mgr = model.get_max_gelman_rubin()
        if verbosity >= 1:
            print(f"[{datetime.now()}] mgr={mgr: 10.4f}", flush=True)
        if it >= max_its:
            break
    return it, mgr


def sample_model(model, ents, tests, max_its=100, burn_its=10, verbosity=0):
    it, _ = run_protocol(model, max_its=max_its, burn_its=burn_its, verbosity=verbosity)
    print(f"Completed {it} iterations.")
    X = pd.DataFrame([dict(symbol=symbol, X=val) for _, symbol, idx, val in model.get_posterior_means('X') if idx == '1']).set_index('symbol')
    T = pd.DataFrame([dict(symbol=symbol, T=val) for _, symbol, idx, val in model.get_posterior_means('T') if idx == '0']).set_index('symbol')
    df = X.merge(T, left_index=True, right_index=True)
    df['uid'] = ents.loc[ents.type == 'Protein'].set_index('name').loc[df.index.values, 'uid']
    return tests.merge(df.reset_index().set_index('uid'), left_index=True, right_index=True).sort_values('X', ascending=False)


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
