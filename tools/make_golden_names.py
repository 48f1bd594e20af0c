#!/usr/bin/env python
"""Generate tests/golden/names_golden.json by running the REFERENCE's own naming / lookup helpers
(extracted with ``ast`` from /root/reference/evaluate_ensemble.py - they only need os/re/np/itertools).
Build-container only; the JSON is the committed fixture.  Usage: python tools/make_golden_names.py"""
import ast
import itertools
import json
import os
import re
import tempfile

import numpy as np

REF = "/root/reference/evaluate_ensemble.py"
WANTED = ["getModelTypeAndTrainingCondition", "get_ModelsNameAndTrainedModelsSubfolder",
          "createModelsTrainingConditionsDictionary", "lookFor_probabilitiesFile",
          "lookFor_UniqueEnsemble_predictionsFile", "lookFor_GlobalEnsemble_predictionsFile", "compute_combinations",
          "normalize", "get_modeltraining_validation_loss"]


def load():
    tree = ast.parse(open(REF).read())
    ns = {"np": np, "os": os, "re": re, "itertools": itertools}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    assert all(w in ns for w in WANTED)
    return ns


def main():
    ns = load()
    out = {"names": [], "dict": [], "mt_tc": [], "lookups": [], "combos": [], "normalize": [], "vei": None}
    for fn, mt, tc, cs, ofs, aug, af in itertools.product(
            [3, 5], ["C3D", "TWOSTREAM_I3D", "R3D_34"], ["_SCRATCH", "_PRETRAINED"], ["unbalanced", "balanced"],
            ["TVL1_precomputed", "FarneBack_onTheFly"], ["non_augmented", "augmented_precomputed"], [0, 3]):
        name, sub = ns["get_ModelsNameAndTrainedModelsSubfolder"](fn, "Trained_models/", mt, tc, cs, ofs, aug, af)
        out["names"].append([[fn, "Trained_models/", mt, tc, cs, ofs, aug, af], name, sub])
        out["mt_tc"].append([sub, list(ns["getModelTypeAndTrainingCondition"](sub))])
    for lst in (["C3D_PRETRAINED", "C3D_SCRATCH", "I3D_SCRATCH"], ["SPECIALCASE_PRETRAINED", "R3D_34_SCRATCH",
                "TWOSTREAM_I3D_PRETRAINED", "TWOSTREAM_I3D_SCRATCH"], ["BOGUS", "R3D_152_SCRATCH"]):
        out["dict"].append([lst, ns["createModelsTrainingConditionsDictionary"](lst)])
        n, combos = ns["compute_combinations"](lst)
        out["combos"].append([lst, n, sorted([list(c) for c in combos])])
    with tempfile.TemporaryDirectory() as td:
        args = (5, td, "C3D", "_SCRATCH", "unbalanced", "TVL1_precomputed")
        for aug, af in (("non_augmented", 0), ("augmented_precomputed", 3)):
            for sets in ("test", "train_val"):
                miss = ns["lookFor_probabilitiesFile"](*args, aug, af, sets)
                stem = ns["get_ModelsNameAndTrainedModelsSubfolder"](5, "x", "C3D", "_SCRATCH", "unbalanced",
                                                                      "TVL1_precomputed", aug, af)[0]
                path = os.path.join(td, sets + "_predicted_probabilities_" + stem + ".csv")
                open(path, "w").close()
                hit = ns["lookFor_probabilitiesFile"](*args, aug, af, sets)
                out["lookups"].append(["prob", aug, af, sets, miss, os.path.relpath(hit, td)])
            stem = ns["get_ModelsNameAndTrainedModelsSubfolder"](5, "x", "C3D", "_SCRATCH", "unbalanced",
                                                                  "TVL1_precomputed", aug, af)[0]
            open(os.path.join(td, "weighted_prediction_results_" + stem + ".csv"), "w").close()
            hit = ns["lookFor_UniqueEnsemble_predictionsFile"](*args, aug, af)
            out["lookups"].append(["unique", aug, af, None, None, os.path.relpath(hit, td)])
        lst = ["C3D_SCRATCH", "I3D_SCRATCH"]
        miss = ns["lookFor_GlobalEnsemble_predictionsFile"](5, td, lst)
        open(os.path.join(td, "global_ensemble_summed_prediction_results_5_folds_C3D_SCRATCH_I3D_SCRATCH_.csv"), "w").close()
        hit = ns["lookFor_GlobalEnsemble_predictionsFile"](5, td, lst)
        out["lookups"].append(["global", lst, None, None, miss, os.path.relpath(hit, td)])
        # validation-error-inverse weights
        hist = os.path.join(td, "3folds_X")
        losses = {}
        for i in range(3):
            os.makedirs(os.path.join(hist, "TestSplit%d" % i))
            for j in range(3):
                if j != i:
                    v = np.random.default_rng(10 * i + j).uniform(0.3, 2.0, 7)
                    losses["%d_%d" % (i, j)] = v.tolist()
                    np.save(os.path.join(hist, "TestSplit%d" % i, "m_split_test%d_val%d_validation_losses.npy" % (i, j)), v)
        out["vei"] = {"losses": losses, "weights": [ns["get_modeltraining_validation_loss"](hist, i).tolist() for i in range(3)]}
    for v in ([0.2, 0.3, 0.5], [0.0, 0.0], [1.0, 3.0, 0.0, 4.0]):
        out["normalize"].append([v, np.asarray(ns["normalize"](np.array(v))).tolist()])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "names_golden.json")
    json.dump(out, open(dst, "w"))
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
