"""End-to-end drop-in test on the B200: the reference's on-disk layout (Trained_models/<name>/
TestSplit<i>/{test,train,val}.csv + <name>_split_test<i>_val<j>_weights.hdf5) is
created with synthetic members and clips, then the reference-compatible entry points are driven
exactly like `python evaluate_ensemble.py -op ...` does, and every artefact is checked against the
oracle: probabilities CSV (bf16 tolerance on the values, exact on the wire format), predictions CSV
(bit-exact vote of the CSV's own float64 values, evaluate_ensemble.py:343-370), cache reuse
(:1161, :1406), GRID_SEARCH weights (:322-339) and the global ensemble (:1329-1474)."""
import ast
import os
from itertools import product

import numpy as np
import pandas as pd
import pytest
import torch

from cse_b200 import ensemble as E
from cse_b200 import graph as G
from cse_b200 import hdf5 as H5
from cse_b200.weights import synthetic_weights
from oracle import models as OM
from oracle import vote as OV

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]

# 11 classes like Crowd-11: the reference's own CSV parser (str.replace of ", dtype=float32)",
# evaluate_ensemble.py:65-73) only works when numpy's repr keeps the dtype suffix on the last value
# line, which it does for 11-wide rows - the test parses with that strict parser on purpose.
MT, TC, FOLDS, NCLS = "I3D", "_SCRATCH", 3, 11
CLIPS_PER_FOLD = 11


def _args(root, op, **kw):
    argv = ["-op", op, "-rf", os.path.join(root, "Results"), "-tmf", os.path.join(root, "Trained_models/"),
            "-fn", str(FOLDS), "-cs", "unbalanced", "-af", "3", "-is", "test", "-hf_vei", os.path.join(root, "Data/Weights/")]
    for k, v in kw.items():
        argv += [k] + (list(v) if isinstance(v, (list, tuple)) else [str(v)])
    return E.build_parser().parse_args(argv)


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("drop_in"))
    shape = G.define_input_shape(MT)
    name, sub = E.get_ModelsNameAndTrainedModelsSubfolder(FOLDS, os.path.join(root, "Trained_models/"), MT, TC,
                                                          "unbalanced", "TVL1_precomputed", "non_augmented", 0)
    rng = np.random.default_rng(77)
    clips_dir = os.path.join(root, "clips")
    os.makedirs(clips_dir)
    fold_rows = []
    for f in range(FOLDS):
        rows = []
        for k in range(CLIPS_PER_FOLD):
            p = os.path.join(clips_dir, "fold%d_clip%d.npy" % (f, k))
            # smooth-ish random frames so members disagree on some clips
            np.save(p, rng.integers(0, 256, shape, dtype=np.uint8))
            rows.append([p, "", "", k % NCLS])
        fold_rows.append(pd.DataFrame(rows, columns=["rgbclips_path", "x_axis_flowclips_path", "y_axis_flowclips_path",
                                                     "class"]))
    g = G.build_model_graph(MT, shape, NCLS)
    # random-weight members give logits in the thousands (softmax saturates to one-hot); rescale the
    # head so probabilities are soft, members disagree and the vote is not trivial
    probe = synthetic_weights(g, seed=999)
    x0 = np.load(fold_rows[0]["rgbclips_path"].values[0])[None]
    head_scale = np.float32(3.0 / float(OM.forward(MT, probe, x0, torch.float32)[0].abs().max()))
    weights = {}
    for i in range(FOLDS):
        d = os.path.join(sub, "TestSplit%d" % i)
        os.makedirs(d)
        fold_rows[i].to_csv(os.path.join(d, "test.csv"))
        others = [j for j in range(FOLDS) if j != i]
        fold_rows[others[0]].to_csv(os.path.join(d, "train.csv"))
        fold_rows[others[1]].to_csv(os.path.join(d, "val.csv"))
        for j in others:
            w = synthetic_weights(g, seed=1000 + 10 * i + j)
            w["predictions"] = [w["predictions"][0] * head_scale, w["predictions"][1]]
            path = os.path.join(d, "%s_split_test%d_val%d_weights.hdf5" % (name, i, j))
            H5.save_member_weights(path, g, w)
            weights[(i, j)] = w
    return dict(root=root, name=name, sub=sub, shape=shape, graph=g, weights=weights, folds=fold_rows)


def _oracle_probs(ds, i, j, frame):
    x = np.stack([np.load(p) for p in frame["rgbclips_path"].values])
    _, probs = OM.forward(MT, ds["weights"][(i, j)], x, torch.float32)
    return probs.numpy()


def test_evaluate_ensembles_sum_end_to_end(dataset, monkeypatch):
    ds = dataset
    monkeypatch.chdir(ds["root"])
    E.main(_args(ds["root"], "Evaluate_ensembles", **{"-mt": MT, "-tc": TC, "-wt": "SUM", "-b": 8, "-w": 1,
                                                         "-ofs": "TVL1_precomputed", "-as": "non_augmented"}))
    prob_csv = os.path.join(ds["root"], "Results", "test_predicted_probabilities_%s.csv" % ds["name"])
    pred_csv = os.path.join(ds["root"], "Results", "weighted_prediction_results_%s.csv" % ds["name"])
    assert os.path.isfile(prob_csv) and os.path.isfile(pred_csv)
    df = pd.read_csv(prob_csv)
    assert list(df.columns[1:]) == ["path", "probabilities"] and len(df) == FOLDS * (FOLDS - 1)
    table = {p: OV.parse_probabilities_cell(c) for p, c in zip(df["path"], df["probabilities"])}     # reference's parser
    preds = pd.read_csv(pred_csv)
    pred_table = {p: ast.literal_eval(c) for p, c in zip(preds["path"], preds["predictions"])}
    for i in range(FOLDS):
        members = []
        for j in [k for k in range(FOLDS) if k != i]:
            key = os.path.join(ds["sub"], "TestSplit%d" % i, "%s_split_test%d_val%d_weights" % (ds["name"], i, j))
            assert key in table, "row key must be the weight path without .hdf5 (evaluate_ensemble.py:1059)"
            got = table[key]
            exp = _oracle_probs(ds, i, j, ds["folds"][i])
            assert got.shape == (CLIPS_PER_FOLD, NCLS) and got.dtype == np.float64
            np.testing.assert_allclose(got, exp, rtol=0, atol=2e-2)            # bf16 members vs fp32 oracle
            np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-5)
            # single-member predictions = argmax of the CSV's own values (evaluate_single_model :86-100)
            assert pred_table[key] == OV.single_model_predictions(got, CLIPS_PER_FOLD, NCLS).tolist()
            members.append(got)
        ens_key = "Ensemble_%s_split_test%d" % (ds["name"], i)
        exp_vote = OV.ensemble_predictions(np.array(members), np.ones(FOLDS - 1))
        assert pred_table[ens_key] == exp_vote.tolist()                        # bit-exact fp64 soft vote


def test_cached_probabilities_are_reused(dataset, monkeypatch):
    """Evaluation resume = file-name cache (evaluate_ensemble.py:180-216, 1161): with the CSV present no
    member may be loaded or run again."""
    ds = dataset
    monkeypatch.chdir(ds["root"])

    def boom(*a, **k):
        raise AssertionError("store_probabilities must not run when the CSV exists")
    monkeypatch.setattr(E, "store_probabilities", boom)
    E.main(_args(ds["root"], "Evaluate_ensembles", **{"-mt": MT, "-tc": TC, "-wt": "MAXIMUM", "-b": 8, "-w": 1,
                                                         "-ofs": "TVL1_precomputed", "-as": "non_augmented"}))
    pred_csv = os.path.join(ds["root"], "Results", "weighted_prediction_results_%s.csv" % ds["name"])
    prob_csv = os.path.join(ds["root"], "Results", "test_predicted_probabilities_%s.csv" % ds["name"])
    df = pd.read_csv(prob_csv)
    table = {p: OV.parse_probabilities_cell(c) for p, c in zip(df["path"], df["probabilities"])}
    preds = pd.read_csv(pred_csv)
    pred_table = {p: ast.literal_eval(c) for p, c in zip(preds["path"], preds["predictions"])}
    for i in range(FOLDS):
        keys = [os.path.join(ds["sub"], "TestSplit%d" % i, "%s_split_test%d_val%d_weights" % (ds["name"], i, j))
                for j in range(FOLDS) if j != i]
        exp = OV.ensemble_predictions(np.array([table[k] for k in keys]), "MAXIMUM")
        assert pred_table["Ensemble_%s_split_test%d" % (ds["name"], i)] == exp.tolist()


def test_grid_search_weights_match_reference_loop(dataset, monkeypatch):
    """GRID_SEARCH (evaluate_ensemble.py:322-339): first best of itertools.product order over the
    normalised 0..1 grid, scored on the train+val probabilities; here all candidates are scored by one
    batched CUDA kernel and must pick the same vector as the sequential restatement."""
    ds = dataset
    monkeypatch.chdir(ds["root"])
    E.main(_args(ds["root"], "Evaluate_ensembles", **{"-mt": MT, "-tc": TC, "-wt": "GRID_SEARCH", "-b": 8, "-w": 1,
                                                         "-ofs": "TVL1_precomputed", "-as": "non_augmented"}))
    saved = np.load(os.path.join(ds["root"], "GRID_SEARCH_%s.npy" % ds["name"]))
    assert saved.shape == (FOLDS, FOLDS - 1)
    tv_csv = os.path.join(ds["root"], "Results", "train_val_predicted_probabilities_%s.csv" % ds["name"])
    df = pd.read_csv(tv_csv)
    table = {p: OV.parse_probabilities_cell(c) for p, c in zip(df["path"], df["probabilities"])}
    grid = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]
    for i in range(FOLDS):
        others = [j for j in range(FOLDS) if j != i]
        keys = [os.path.join(ds["sub"], "TestSplit%d" % i, "%s_split_test%d_val%d_weights" % (ds["name"], i, j))
                for j in others]
        yh = np.array([table[k] for k in keys])
        labels = np.concatenate([ds["folds"][others[0]]["class"].values, ds["folds"][others[1]]["class"].values])
        best_score, best_w = 0.0, None
        for w in product(grid, repeat=len(keys)):
            if len(set(w)) == 1:
                continue
            nw = np.array(w) / np.linalg.norm(w, 1)
            score = float(np.mean(OV.ensemble_predictions(yh, nw) == labels))
            if score > best_score:
                best_score, best_w = score, nw
        np.testing.assert_array_equal(saved[i], best_w)


def test_global_ensemble(dataset, monkeypatch):
    ds = dataset
    monkeypatch.chdir(ds["root"])
    mlist = [MT + TC]
    acc = E.global_evaluate_ensembles(FOLDS, os.path.join(ds["root"], "Trained_models/"), mlist,
                                      os.path.join(ds["root"], "Results"))
    out = os.path.join(ds["root"], "Results", "global_ensemble_summed_prediction_results_%d_folds_%s_.csv" % (FOLDS, mlist[0]))
    assert os.path.isfile(out) and 0.0 <= acc <= 1.0
    preds = pd.read_csv(out)
    assert list(preds["path"]) == ["Global_Ensemble_%s_split_test%d" % (mlist[0], i) for i in range(FOLDS)]
    for i in range(FOLDS):
        merged = os.path.join(ds["root"], "Results", "global_ensemble_probabilities_%s_TestFold%d_%dfolds.csv" % (mlist[0], i, FOLDS))
        assert os.path.isfile(merged)


# --------------------------------------------------------------------------- Combine_ensembles under two ranks
def _combine_worker(rank, world, port, root, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.chdir(root)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ordered = E.combine_ensembles(FOLDS, os.path.join(root, "Trained_models/"), ["C3D_SCRATCH", "R3D_18_SCRATCH", "I3D_PRETRAINED"],
                                      os.path.join(root, "Results"))
        with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
            f.write(repr([(tuple(k), round(float(v), 9)) for k, v in ordered.items()]))
    finally:
        dist.destroy_process_group()


def test_combine_ensembles_two_ranks(tmp_path):
    """Combine_ensembles (evaluate_ensemble.py:1298-1326) under torchrun-style ranks: every rank must walk the 2^n - 1
    model subsets in the same order (they meet in barriers; rank 0 alone writes the merged CSVs the others read).  The
    member probabilities are cached CSVs (the reference's resume-by-file-name, :180-216), so only the lookup / merge /
    vote path runs; both ranks share cuda:0 for the vote kernel."""
    import socket
    import torch.multiprocessing as mp
    root = str(tmp_path)
    rng = np.random.default_rng(3)
    n_clips = NCLS + 2          # nb_classes = len(set(test_data['class'])) (evaluate_ensemble.py:1028): every class occurs
    labels = [int(v) for v in rng.permutation(NCLS)] + [3, 7]
    for mt, tc in (("C3D", "_SCRATCH"), ("R3D_18", "_SCRATCH"), ("I3D", "_PRETRAINED")):
        name, sub = E.get_ModelsNameAndTrainedModelsSubfolder(FOLDS, os.path.join(root, "Trained_models/"), mt, tc, "unbalanced",
                                                              "TVL1_precomputed", "non_augmented", 0)
        rows = []
        for i in range(FOLDS):
            d = os.path.join(sub, "TestSplit%d" % i)
            os.makedirs(d)
            pd.DataFrame([["clip%d" % k, "", "", labels[k]] for k in range(n_clips)],
                         columns=["rgbclips_path", "x_axis_flowclips_path", "y_axis_flowclips_path", "class"]).to_csv(
                os.path.join(d, "test.csv"))
            for j in [k for k in range(FOLDS) if k != i]:
                logits = rng.standard_normal((n_clips, NCLS)).astype(np.float32)
                p = (np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(np.float32)
                rows.append([os.path.join(d, "%s_split_test%d_val%d_weights" % (name, i, j)), E.convert_array2listofarrays(p)])
        os.makedirs(os.path.join(root, "Results"), exist_ok=True)
        pd.DataFrame(rows, columns=["path", "probabilities"]).to_csv(
            os.path.join(root, "Results", "test_predicted_probabilities_%s.csv" % name))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_combine_worker, args=(2, port, root, root), nprocs=2, join=True)
    a, b = (open(os.path.join(root, "rank%d.txt" % r)).read() for r in range(2))
    assert a == b and a.count("(") >= 7          # 2^3 - 1 subsets, identical order and accuracies on both ranks
    merged = [f for f in os.listdir(os.path.join(root, "Results")) if f.startswith("global_ensemble_summed_prediction_results_")]
    assert len(merged) == 7
    for f in merged:
        for cell in pd.read_csv(os.path.join(root, "Results", f))["predictions"]:
            assert len(ast.literal_eval(cell)) == n_clips
