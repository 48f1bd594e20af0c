"""End-to-end drop-in test of the FarneBack_onTheFly TwoStream variant (train.py:294-332, 223-239; the SPECIALCASE of the global
list, evaluate_ensemble.py:1365-1386) on the B200: the reference's on-disk layout with TwoStream-I3D members and short
videos, `Evaluate_ensembles` driven like the CLI - clips assembled and the dense flow computed on the GPU (cse_farneback),
two members per fold with their stems fused, float32 flow into the flow tower - and the probabilities CSV checked against the
oracle fed with the REFERENCE's own loader outputs (cv2 resize + cv2.calcOpticalFlowFarneback)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from cse_b200 import clips as CL
from cse_b200 import ensemble as E
from cse_b200 import graph as G
from cse_b200 import hdf5 as H5
from cse_b200.weights import synthetic_weights
from oracle import models as OM
from oracle import vote as OV

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]

MT, TC, FOLDS, NCLS, OFS = "TWOSTREAM_I3D", "_SCRATCH", 3, 11, "FarneBack_onTheFly"


def _video(rng, k):
    """22 frames 90 x 120 of a smooth texture drifting by (1, 2) pixels per frame, BGR uint8."""
    import cv2
    base = cv2.GaussianBlur(rng.integers(0, 256, (90 + 30, 120 + 60, 3)).astype(np.float32), (9, 9), 2.5)
    return np.stack([np.clip(base[i:i + 90, 2 * i:2 * i + 120], 0, 255).astype(np.uint8) for i in range(22)])


def test_farneback_twostream_evaluate_ensembles(tmp_path, monkeypatch):
    root = str(tmp_path)
    shape = G.define_input_shape(MT)
    t, h, w = shape[:3]
    name, sub = E.get_ModelsNameAndTrainedModelsSubfolder(FOLDS, os.path.join(root, "Trained_models/"), MT, TC, "unbalanced", OFS,
                                                          "non_augmented", 0)
    rng = np.random.default_rng(5)
    os.makedirs(os.path.join(root, "clips"))
    folds = []
    for f in range(FOLDS):
        rows = []
        for k in range(NCLS):
            p = os.path.join(root, "clips", "fold%d_clip%d.npy" % (f, k))
            np.save(p, _video(rng, k))
            rows.append([p, "", "", k])
        folds.append(pd.DataFrame(rows, columns=["rgbclips_path", "x_axis_flowclips_path", "y_axis_flowclips_path", "class"]))
    g = G.build_model_graph(MT, shape, NCLS)
    rgb0, flow0 = CL.load_farneback_twostream_clip(folds[0]["rgbclips_path"].values[0], t, h, w)        # the reference's loader path
    assert rgb0.dtype == np.uint8 and flow0.dtype == np.float32 and float(np.abs(flow0).max()) > 1.0
    probe = synthetic_weights(g, seed=999)
    head_scale = np.float32(3.0 / float(OM.forward(MT, probe, [rgb0[None], flow0[None]], torch.float32)[0].abs().max()))
    weights = {}
    for i in range(FOLDS):
        d = os.path.join(sub, "TestSplit%d" % i)
        os.makedirs(d)
        folds[i].to_csv(os.path.join(d, "test.csv"))
        others = [k for k in range(FOLDS) if k != i]
        folds[others[0]].to_csv(os.path.join(d, "train.csv"))
        folds[others[1]].to_csv(os.path.join(d, "val.csv"))
        for j in others:
            wts = synthetic_weights(g, seed=2000 + 10 * i + j)
            wts["predictions"] = [wts["predictions"][0] * head_scale, wts["predictions"][1]]
            H5.save_member_weights(os.path.join(d, "%s_split_test%d_val%d_weights.hdf5" % (name, i, j)), g, wts)
            weights[(i, j)] = wts
    monkeypatch.chdir(root)
    monkeypatch.delenv("CSE_CPU_FLOW", raising=False)
    monkeypatch.delenv("CSE_CPU_RESIZE", raising=False)
    argv = ["-op", "Evaluate_ensembles", "-rf", os.path.join(root, "Results"), "-tmf", os.path.join(root, "Trained_models/"),
            "-fn", str(FOLDS), "-cs", "unbalanced", "-af", "3", "-is", "test", "-hf_vei", os.path.join(root, "Data/Weights/"),
            "-mt", MT, "-tc", TC, "-wt", "SUM", "-b", "4", "-w", "2", "-ofs", OFS, "-as", "non_augmented"]
    E.main(E.build_parser().parse_args(argv))
    df = pd.read_csv(os.path.join(root, "Results", "test_predicted_probabilities_%s.csv" % name))
    assert len(df) == FOLDS * (FOLDS - 1)
    table = {p: OV.parse_probabilities_cell(c) for p, c in zip(df["path"], df["probabilities"])}
    for got in table.values():
        assert got.shape == (NCLS, NCLS) and np.isfinite(got).all()
        np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-5)
    # fold 0, both members, three clips: the oracle on what the reference's loader (cv2 only) produces for the same videos
    check = [0, 4, 9]
    pairs = [CL.load_farneback_twostream_clip(folds[0]["rgbclips_path"].values[k], t, h, w) for k in check]
    x = [np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])]
    members = []
    for j in (1, 2):
        key = os.path.join(sub, "TestSplit0", "%s_split_test0_val%d_weights" % (name, j))
        _, exp = OM.forward(MT, weights[(0, j)], x, torch.float32)
        np.testing.assert_allclose(table[key][check], exp.numpy(), rtol=0, atol=2e-2)           # bf16 members vs fp32 oracle
        members.append(table[key])
    assert float(np.std(members[0], axis=0).max()) > 0                    # soft, clip-dependent probabilities
    preds = pd.read_csv(os.path.join(root, "Results", "weighted_prediction_results_%s.csv" % name))
    import ast
    pred_table = {p: ast.literal_eval(c) for p, c in zip(preds["path"], preds["predictions"])}
    exp_vote = OV.ensemble_predictions(np.array(members), np.ones(2))
    assert pred_table["Ensemble_%s_split_test0" % name] == exp_vote.tolist()
