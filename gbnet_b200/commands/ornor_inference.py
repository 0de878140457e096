"""``gbn-ornor-inference``: three CSV files in (entities, relations, evidence), one result table out.

The command-line surface is the reference's (gbnet/commands/ornor_inference.py:7-45, console script
of setup.py:64-66): the same option names and defaults, so existing invocations keep working.
``--n_graphs``, ``--seed`` and ``--device`` are additions of this implementation.

    python -m gbnet_b200.commands.ornor_inference --ents ents.csv --rels rels.csv --evid evid.csv --out res.csv
"""
import argparse

import pandas as pd

from gbnet_b200 import utils

# option name -> default; grouped by what happens to the value
INPUT_FILES = ("ents", "rels", "evid")                      # read with pandas.read_csv
MODEL_SWITCHES = ("const_params", "noise_listen_children", "comp_yprob")
MODEL_FLOATS = {"t_focus": 1., "t_lmargin": 2., "t_hmargin": 2.,
                "z_alpha": 2000., "z_beta": 1., "z0_alpha": 200., "z0_beta": 1., "s_leniency": 0.01}
MODEL_INTS = {"n_graphs": 3, "seed": 0, "device": 0}        # not in the reference's CLI
BURN_ITERATIONS = 10                                         # fixed by the reference (ornor_inference.py:41)


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="gbn-ornor-inference",
                                 description="OR-NOR network inference on the GPU: TF activity from DE evidence.")
    for name in INPUT_FILES:
        ap.add_argument("--" + name, required=True, metavar="CSV")
    ap.add_argument("--out", required=True, metavar="CSV", help="where the result table goes")
    for name in MODEL_SWITCHES:
        ap.add_argument("--" + name, action="store_true")
    for name, default in MODEL_FLOATS.items():
        ap.add_argument("--" + name, type=float, default=default)
    for name, default in MODEL_INTS.items():
        ap.add_argument("--" + name, type=int, default=default)
    ap.add_argument("--max_its", type=int, default=1000, help="upper bound on sampling iterations")
    return ap


def run(opts: argparse.Namespace) -> pd.DataFrame:
    frames = {name: pd.read_csv(getattr(opts, name)) for name in INPUT_FILES}
    model_kwargs = {name: getattr(opts, name) for name in (*MODEL_SWITCHES, *MODEL_FLOATS, *MODEL_INTS)}
    screen = utils.get_tests(frames["evid"], frames["rels"])
    model = utils.get_model(frames["ents"], frames["rels"], frames["evid"], **model_kwargs)
    return utils.sample_model(model, frames["ents"], screen, burn_its=BURN_ITERATIONS, max_its=opts.max_its,
                              verbosity=1)


def main(argv=None):
    opts = build_parser().parse_args(argv)
    run(opts).to_csv(opts.out)


if __name__ == "__main__":
    main()
