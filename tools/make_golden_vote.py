#!/usr/bin/env python
"""Generate tests/golden/vote_golden.json by running the REFERENCE's own vote code.

evaluate_ensemble.py cannot be imported (``from train import *`` pulls Keras/TF, absent
here), but the functions on the vote path only need numpy/pandas/ast/os/sklearn.  We
parse /root/reference/evaluate_ensemble.py with ``ast``, compile exactly the function
definitions named below - unmodified - into a namespace that provides those modules,
and run them on seeded inputs written to a temporary probabilities CSV in the
reference's own format (DataFrame(columns=['path','probabilities']).to_csv, :1061-1063).

Run only in the build container (needs /root/reference); the JSON it writes is the
committed fixture.  Usage: python tools/make_golden_vote.py
"""
import ast
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd
from sklearn.metrics import accuracy_score

REF = "/root/reference/evaluate_ensemble.py"
WANTED = ["convert_str2array", "convert_array2listofarrays", "evaluate_single_model",
          "ensemble_predictions", "evaluate_ensemble", "normalize"]


def load_reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np, "pd": pd, "ast": ast, "os": os, "accuracy_score": accuracy_score}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, REF, "exec"), ns)
    missing = [w for w in WANTED if w not in ns]
    if missing:
        raise RuntimeError("reference functions not found: %s" % missing)
    return ns


def softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def main():
    ns = load_reference_functions()
    rng = np.random.default_rng(20261018)
    cases = []
    specs = [  # (M, N, C, kind)
        (4, 7, 11, "random"), (4, 64, 11, "random"), (12, 33, 11, "random"),
        (1, 5, 11, "random"), (4, 16, 11, "ties"), (3, 9, 4, "ties"), (32, 10, 11, "random"),
        (4, 12, 11, "peaky"),
    ]
    with tempfile.TemporaryDirectory() as td:
        for ci, (m, n, c, kind) in enumerate(specs):
            if kind == "random":
                probs = softmax(rng.standard_normal((m, n, c)) * 2.0).astype(np.float32)
            elif kind == "peaky":
                probs = softmax(rng.standard_normal((m, n, c)) * 30.0).astype(np.float32)
            else:   # exact ties between classes and between members (dyadic values)
                probs = (rng.integers(0, 4, (m, n, c)) / 8.0).astype(np.float32)
            members = [os.path.join(td, "model_case%d_m%d_weights.hdf5" % (ci, j)) for j in range(m)]
            rows = [[os.path.splitext(p)[0], ns["convert_array2listofarrays"](probs[j])]
                    for j, p in enumerate(members)]
            csv = os.path.join(td, "probs_%d.csv" % ci)
            pd.DataFrame(rows, columns=["path", "probabilities"]).to_csv(csv)
            labels = rng.integers(0, c, n)
            cells = pd.read_csv(csv)["probabilities"].tolist()
            parsed = [ns["convert_str2array"](s) for s in cells]
            weights = {
                "SUM": np.ones(m),
                "WEIGHTED": ns["normalize"](rng.uniform(0.1, 1.0, m)) if m > 1 else np.ones(1),
            }
            out = {"M": m, "N": n, "C": c, "kind": kind,
                   "probs_f32_hex": probs.tobytes().hex(),
                   "cells": cells,
                   "parsed_f64_hex": np.asarray(parsed, np.float64).tobytes().hex(),
                   "labels": labels.tolist(), "votes": {}, "single": []}
            for wname, w in weights.items():
                pred = ns["ensemble_predictions"](members, w, labels, csv, c)
                acc, pred2 = ns["evaluate_ensemble"](members, w, csv, labels, c)
                assert (pred == pred2).all()
                out["votes"][wname] = {"weights_f64_hex": np.asarray(w, np.float64).tobytes().hex(),
                                       "pred": np.asarray(pred).tolist(), "acc": float(acc)}
            pred = ns["ensemble_predictions"](members, "MAXIMUM", labels, csv, c)
            out["votes"]["MAXIMUM"] = {"pred": np.asarray(pred).tolist()}
            for p in members:
                score, sp = ns["evaluate_single_model"](p, labels, csv, c)
                out["single"].append({"pred": np.asarray(sp).tolist(), "acc": float(score)})
            cases.append(out)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                       "vote_golden.json")
    with open(dst, "w") as f:
        json.dump({"generator": "tools/make_golden_vote.py", "reference": REF,
                   "numpy": np.__version__, "pandas": pd.__version__, "cases": cases}, f)
    print("wrote", os.path.normpath(dst), "cases:", len(cases))


if __name__ == "__main__":
    sys.exit(main())
