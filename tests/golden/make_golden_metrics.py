"""Golden fixture for the evaluation metrics, produced by running the REFERENCE's own `shot_metrics`.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_metrics.py

`agedb-dir/train.py` cannot be imported (argparse / tensorboard_logger / folder creation run at import time,
SURVEY.md §8c), so the function's source is lifted out of the file with `ast` at generation time and executed
unmodified in a namespace holding the names it uses (numpy, torch, scipy's gmean, defaultdict).  Only its
outputs on seeded inputs are stored; nothing from the reference is copied into the repository.
"""
import ast
import os
from collections import defaultdict

import numpy as np
import pandas as pd
import torch
from scipy.stats import gmean

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/agedb-dir/train.py"


def reference_function(name):
    tree = ast.parse(open(SRC).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "torch": torch, "gmean": gmean, "defaultdict": defaultdict}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    return ns[name]


def main():
    shot_metrics = reference_function("shot_metrics")
    df = pd.read_csv("/root/reference/agedb-dir/data/agedb.csv")
    train = df[df["split"] == "train"]["age"].values
    test = df[df["split"] == "test"]["age"].values.astype(np.float32)
    out = {}
    for tag, seed, noise in (("a", 0, 7.0), ("b", 1, 2.0)):
        rng = np.random.RandomState(seed)
        preds = (test + rng.randn(test.size).astype(np.float32) * noise).astype(np.float32)
        if tag == "b":
            preds[::97] = test[::97]          # exact hits: |err| = 0 -> log 0 -> G-Mean 0 in the groups they fall in
        sd = shot_metrics(preds, test, train)
        out[f"{tag}_preds"] = preds
        out[f"{tag}_labels"] = test
        out[f"{tag}_ref"] = np.asarray([[sd[g][m] for m in ("mse", "l1", "gmean")] for g in ("many", "median", "low")],
                                       dtype=np.float64)
        err = np.abs(preds - test)
        with np.errstate(divide="ignore"):
            out[f"{tag}_overall"] = np.asarray([np.mean((preds - test) ** 2, dtype=np.float64),
                                                np.mean(err, dtype=np.float64), float(gmean(err, axis=None))])
    out["train_labels"] = train.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    for k, v in out.items():
        if k.endswith("_ref") or k.endswith("_overall"):
            print(k, v)


if __name__ == "__main__":
    main()
