"""Host-side logic (no GPU): naming / lookup contracts against golden vectors generated from the
reference's own helpers (tools/make_golden_names.py), the probabilities-CSV wire format, the
Keras-HDF5 reader/writer, clip sharding, and the C-ABI library's exported symbols."""
import ctypes
import json
import os
import re

import numpy as np
import pandas as pd
import pytest

from cse_b200 import ensemble as E
from cse_b200 import graph as G
from cse_b200 import hdf5 as H5
from cse_b200 import runtime as rt
from cse_b200.weights import assign_positional, synthetic_weights
from oracle import vote as OV

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "names_golden.json")))


# --------------------------------------------------------------------------- naming contracts
def test_models_name_and_subfolder_match_reference():
    for args, name, sub in GOLD["names"]:
        got_name, got_sub = E.get_ModelsNameAndTrainedModelsSubfolder(*args)
        assert got_name == name
        assert got_sub == sub


def test_model_type_and_training_condition_match_reference():
    for sub, (mt, tc) in GOLD["mt_tc"]:
        assert list(E.getModelTypeAndTrainingCondition(sub)) == [mt, tc]


def test_models_dictionary_and_combinations_match_reference():
    for lst, expected in GOLD["dict"]:
        assert E.createModelsTrainingConditionsDictionary(lst) == expected
    for lst, n, combos in GOLD["combos"]:
        got_n, got = E.compute_combinations(lst)
        assert got_n == n
        assert sorted(list(c) for c in got) == combos


def test_combinations_order_is_identical_in_every_process():
    """Under torchrun every rank walks Combine_ensembles' combinations and meets the others in barriers /
    all-gathers per combination, so the order must not depend on per-process string hashing."""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); from cse_b200 import ensemble as E; "
            "print(E.compute_combinations(['C3D_SCRATCH', 'I3D_PRETRAINED', 'R3D_34_SCRATCH', 'SPECIALCASE_PRETRAINED'])[1])"
            % ROOT)
    outs = {subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONHASHSEED=str(seed)),
                           capture_output=True, text=True, check=True).stdout for seed in (1, 2, 3)}
    assert len(outs) == 1


def test_lookups_match_reference(tmp_path):
    td = str(tmp_path)
    args = (5, td, "C3D", "_SCRATCH", "unbalanced", "TVL1_precomputed")
    for kind, a, af, sets, miss, hit in GOLD["lookups"]:
        if kind == "prob":
            assert E.lookFor_probabilitiesFile(*args, a, af, sets) is miss      # None before the file exists
            open(os.path.join(td, hit), "w").close()
            assert os.path.relpath(E.lookFor_probabilitiesFile(*args, a, af, sets), td) == hit
        elif kind == "unique":
            open(os.path.join(td, hit), "w").close()
            assert os.path.relpath(E.lookFor_UniqueEnsemble_predictionsFile(*args, a, af), td) == hit
        else:
            assert E.lookFor_GlobalEnsemble_predictionsFile(5, td, a) is miss
            open(os.path.join(td, hit), "w").close()
            assert os.path.relpath(E.lookFor_GlobalEnsemble_predictionsFile(5, td, a), td) == hit


def test_normalize_and_validation_error_inverse_match_reference(tmp_path):
    for v, expected in GOLD["normalize"]:
        assert np.array_equal(np.asarray(E.normalize(np.array(v))), np.array(expected))
    hist = tmp_path / "3folds_X"
    for key, losses in GOLD["vei"]["losses"].items():
        i, j = key.split("_")
        d = hist / ("TestSplit" + i)
        d.mkdir(parents=True, exist_ok=True)
        np.save(str(d / ("m_split_test%s_val%s_validation_losses.npy" % (i, j))), np.array(losses))
    for i, w in enumerate(GOLD["vei"]["weights"]):
        assert np.array_equal(E.get_modeltraining_validation_loss(str(hist), i), np.array(w))


# --------------------------------------------------------------------------- CSV wire format
def test_probabilities_csv_roundtrip_through_pandas(tmp_path):
    """store_probabilities writes str(list of float32 arrays) cells (evaluate_ensemble.py:1058-1063);
    the parser must give the float64 values ast.literal_eval gives the reference (:65-73)."""
    rng = np.random.default_rng(3)
    logits = rng.normal(size=(3, 40, 11)).astype(np.float32) * 6
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
    rows = [["Trained_models/x/TestSplit0/m%d_weights" % j, E.convert_array2listofarrays(p[j])] for j in range(3)]
    path = str(tmp_path / "test_predicted_probabilities_x.csv")
    pd.DataFrame(rows, columns=["path", "probabilities"]).to_csv(path)
    df = pd.read_csv(path)
    for j in range(3):
        cell = df["probabilities"].values[j]
        got = E.convert_str2array(cell)
        assert got.dtype == np.float64 and got.shape == (40, 11)
        assert np.array_equal(got, OV.parse_probabilities_cell(cell))      # the reference's own parse
        assert np.array_equal(got.astype(np.float32), p[j])                # 8 digits round-trip float32 exactly
    tab = E._ProbabilityCache().table(path)
    assert sorted(tab) == sorted(r[0] for r in rows)
    assert np.array_equal(E._ProbabilityCache().member(path, rows[1][0] + ".hdf5"), E.convert_str2array(
        df["probabilities"].values[1]))


def test_predictions_csv_is_literal_evaluable(tmp_path):
    import ast
    preds = [np.array([3, 5, 0, 10], dtype=np.int64)]
    path = str(tmp_path / "weighted_prediction_results_x.csv")
    pd.DataFrame([["Ensemble_x_split_test0", E._plain_ints(preds[0])]],
                 columns=["path", "predictions"]).to_csv(path)
    cell = pd.read_csv(path)["predictions"].values[0]
    assert cell == "[3, 5, 0, 10]"                          # what numpy 1.x / the reference wrote
    assert ast.literal_eval(cell) == [3, 5, 0, 10]          # the reference's consumers (evaluate_ensemble.py:424)


# --------------------------------------------------------------------------- HDF5
@pytest.mark.parametrize("mt,shape", [("C3D", (16, 32, 32, 3)), ("R3D_18", (16, 32, 32, 3)), ("I3D", (12, 64, 64, 3)),
                                      ("TWOSTREAM_I3D", (12, 64, 64, 0))])
def test_keras_hdf5_write_read_positional(tmp_path, mt, shape):
    """Writer -> reader -> positional assignment (model.load_weights(by_name=False), App. C) is the
    identity on every tensor, with weight-less layers present in layer_names."""
    g = G.build_model_graph(mt, shape, 11)
    w = synthetic_weights(g, seed=7, nontrivial=True)
    path = str(tmp_path / "m_weights.hdf5")
    H5.save_member_weights(path, g, w)
    layer_names, weight_names, arrays = H5.read_keras_weights(path)
    assert layer_names == g.keras_layer_order()
    assert any(len(a) == 0 for a in arrays)                 # pools / activations are listed without weights
    back = assign_positional(g, arrays)
    assert set(back) == set(w)
    for name in w:
        assert len(back[name]) == len(w[name])
        for a, b in zip(back[name], w[name]):
            assert a.dtype == np.float32 and np.array_equal(a, b)
    for ln, wn in zip(layer_names, weight_names):
        for n in wn:
            assert n.startswith(ln + "/") and n.endswith(":0")


def test_hdf5_reader_on_third_party_file():
    """A real HDF5 file not produced by our writer (superblock v0 + symbol-table groups, the family
    h5py 2.x / Keras 2.2.4 wrote): scipy ships one as a MATLAB v7.4 test fixture."""
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    f = H5.H5File(path)
    root = f["/"]
    keys = sorted(root.keys())
    assert len(keys) > 0
    found = 0
    for k in keys:
        obj = root[k]
        if not obj.is_group:
            arr = obj.read()
            assert isinstance(arr, np.ndarray)
            found += 1
    assert found > 0


def test_hdf5_positional_mismatch_fails_loudly(tmp_path):
    g = G.build_model_graph("C3D", (16, 32, 32, 3), 11)
    w = synthetic_weights(g, seed=1)
    path = str(tmp_path / "m_weights.hdf5")
    H5.save_member_weights(path, g, w)
    other = G.build_model_graph("R3D_18", (16, 32, 32, 3), 11)
    _, _, arrays = H5.read_keras_weights(path)
    with pytest.raises(Exception):
        assign_positional(other, arrays)


# --------------------------------------------------------------------------- sharding
@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (8, 4), (10, 8), (257, 8)])
def test_shard_indices_partition(n, world):
    parts = [E.shard_indices(n, r, world) for r in range(world)]
    assert np.array_equal(np.concatenate(parts), np.arange(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1


# --------------------------------------------------------------------------- C ABI
def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cse.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cse_[a-z0-9_]+)\s*\(", src)))


def test_c_abi_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert "cse_plan_run" in names and "cse_vote" in names and "cse_preprocess" in names
    assert sorted(rt.EXPORTS) == names, "runtime.EXPORTS and include/cse.h disagree"
    lib = rt.load_library()                      # raises if the library is not built: no CPU fallback
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.cse_abi_version() == rt.ABI_VERSION


def test_cse_op_struct_matches_header_layout():
    """ctypes mirror of struct cse_op: field order and count follow include/cse.h."""
    src = open(os.path.join(ROOT, "include", "cse.h")).read()
    body = src[src.index("typedef struct cse_op {") + len("typedef struct cse_op {"):src.index("} cse_op;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        m = re.match(r"\s*(int32_t|int64_t|float)\s+(.*)", decl.strip(), flags=re.S)
        if m:
            for part in m.group(2).split(","):
                fields.append(re.sub(r"\[.*\]", "", part).strip())
    assert fields == [f[0] for f in rt.CseOp._fields_]
    assert ctypes.sizeof(rt.CseOp) % 8 == 0


def test_compute_entry_points_fail_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rt.CseError):
        rt.require_cuda()
    lib = rt.load_library()
    sm = ctypes.c_int()
    assert lib.cse_device_info(ctypes.byref(sm), None, None) != 0          # CUDA error surfaces as a status
    assert b"CUDA" in lib.cse_last_error()
