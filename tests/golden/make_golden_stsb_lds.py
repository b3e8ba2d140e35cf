"""Golden fixture for the STS-B re-weighting / LDS block, produced by running the REFERENCE's own lines.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_stsb_lds.py

`sts-b-dir/tasks.py` cannot be imported (nltk is not installed) and the weighting code is inlined in `load_tsv`
(tasks.py:44-73), so the `if args is not None and args.reweight != 'none':` block is lifted out of the function's
AST at generation time and executed unmodified on the real training scores (column `score` of
glue_data/STS-B/train_new.tsv, parsed with the reference's `targ_fn = lambda x: np.float32(x)`), with
`get_lds_kernel_window` lifted the same way from sts-b-dir/util.py.  Only inputs and outputs are stored.
"""
import ast
import logging
import os
from types import SimpleNamespace

import numpy as np
from scipy.ndimage import convolve1d, gaussian_filter1d
from scipy.signal.windows import triang

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/sts-b-dir"


def lift_function(path, name, ns):
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def lift_weight_block(ns):
    """def reference_weights(targs, args, sent1s=None, sent2s=None): <the reference's block, verbatim AST>"""
    tree = ast.parse(open(os.path.join(REF, "tasks.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "load_tsv")
    block = next(n for n in fn.body if isinstance(n, ast.If) and "reweight" in ast.unparse(n.test))
    wrapper = ast.parse("def reference_weights(targs, args, sent1s=None, sent2s=None):\n    pass")
    wrapper.body[0].body = [block]
    ast.fix_missing_locations(wrapper)
    exec(compile(wrapper, os.path.join(REF, "tasks.py"), "exec"), ns)
    return ns["reference_weights"]


def main():
    ns = {"np": np, "convolve1d": convolve1d, "gaussian_filter1d": gaussian_filter1d, "triang": triang,
          "logging": logging}
    lift_function(os.path.join(REF, "util.py"), "get_lds_kernel_window", ns)
    ref = lift_weight_block(ns)
    scores = []
    with open(os.path.join(REF, "glue_data", "STS-B", "train_new.tsv"), encoding="utf-8") as fh:
        fh.readline()
        for row in fh:
            cols = row.rstrip("\n").split("\t")
            if len(cols) > 9 and cols[9]:
                scores.append(np.float32(cols[9]))
    out = {"scores": np.asarray(scores, dtype=np.float32)}
    for tag, kw in (("inv", dict(reweight="inverse", lds=False)),
                    ("sqrt", dict(reweight="sqrt_inv", lds=False)),
                    ("inv_lds_gau_5_2", dict(reweight="inverse", lds=True, lds_kernel="gaussian", lds_ks=5, lds_sigma=2)),
                    ("sqrt_lds_gau_5_2", dict(reweight="sqrt_inv", lds=True, lds_kernel="gaussian", lds_ks=5, lds_sigma=2)),
                    ("sqrt_lds_lap_9_1", dict(reweight="sqrt_inv", lds=True, lds_kernel="laplace", lds_ks=9, lds_sigma=1)),
                    ("inv_lds_tri_5", dict(reweight="inverse", lds=True, lds_kernel="triang", lds_ks=5, lds_sigma=2))):
        args = SimpleNamespace(bucket_num=50, lds_kernel="gaussian", lds_ks=5, lds_sigma=2, **kw) if "lds_kernel" not in kw \
            else SimpleNamespace(bucket_num=50, **kw)
        _, _, weights, _ = ref(list(scores), args)
        out[f"w_{tag}"] = np.asarray(weights, dtype=np.float64)
        print(tag, len(weights), float(np.min(weights)), float(np.max(weights)), float(np.mean(weights)))
    hist, edges = np.histogram(scores, bins=50, range=(0., 5.))
    out["hist"] = hist.astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "lds_stsb.npz"), **out)


if __name__ == "__main__":
    main()
