"""Host-side mirror of the reference's ensemble evaluation (evaluate_ensemble.py), hot path only.

Same function names, argument order, file/folder/CSV contracts and printed lines as the reference
for: Store_models_probabilities, Evaluate_ensembles, Global_evaluate_models, Combine_ensembles
(evaluate_ensemble.py:1002-1474, 1481-1796).  What changed underneath:

* members run through libcse_b200 (`zoo.evaluate_load_model` -> `Member`), all members of a test
  fold share one device workspace, every clip is decoded and uploaded ONCE for all members
  (the reference re-decodes it per member, evaluate_ensemble.py:1048-1056) and batched
  (the reference hard-wires batch 1, :1032-1040);
* the vote (`ensemble_predictions`, :343-370) runs in the CUDA vote kernel on the float64 values
  parsed from the probabilities CSV (bit-exact argmax), each CSV is parsed once per process
  instead of once per call (:345, called ~14.6k times by grid_search :335);
* `grid_search` / differential evolution (:302-339) evaluate all candidate weight vectors in one
  batched vote kernel;
* with torch.distributed initialised (one process per GPU), clips are sharded over the ranks and
  the per-clip probabilities are all-gathered; every rank then writes identical results.

Plotting operations (Confusion_matrices, Difference_matrices, StickDiagrams..., :384-999) are
post-hoc matplotlib reports outside the accelerated path and are not provided.
"""
from __future__ import annotations

import argparse
import ast
import itertools
import os
import re
import traceback
from itertools import product
from typing import Dict, List, Tuple

import numpy as np
import pandas as pd

from . import zoo
from .clips import ClipSequence

# --------------------------------------------------------------------------- #
# name / lookup helpers (evaluate_ensemble.py:17-31, 105-275)
# --------------------------------------------------------------------------- #
MODEL_TYPE_REGEX = "(TWOSTREAM_I3D|I3D|C3D|R3D_18|R3D_34|R3D_50|R3D_101|R3D_152)"


def getModelTypeAndTrainingCondition(model_name):
    training_condition = re.search("(_PRETRAINED|_SCRATCH)", model_name)[0]
    model_type = re.search(MODEL_TYPE_REGEX, model_name)[0]
    return model_type, training_condition


def _stem(folds_number, model_type, training_condition, classes_status, optical_flow_status, augmentation_status,
          augmentation_frequency):
    name = "%sfolds_%s%s_CS_%s_OF_%s_AS_%s" % (folds_number, model_type, training_condition, classes_status,
                                               optical_flow_status, augmentation_status)
    if augmentation_status == "augmented_precomputed":
        name += "_Freq" + str(augmentation_frequency)
    return name


def get_ModelsNameAndTrainedModelsSubfolder(folds_number, trained_models_folder, model_type, training_condition,
                                            classes_status, optical_flow_status, augmentation_status,
                                            augmentation_frequency):
    models_name = _stem(folds_number, model_type, training_condition, classes_status, optical_flow_status,
                        augmentation_status, augmentation_frequency)
    return models_name, os.path.join(trained_models_folder, models_name)


def createModelsTrainingConditionsDictionary(models_list):
    """{"C3D": ["_PRETRAINED", ...], ...}; 'SPECIALCASE' = the augmented Farneback TwoStream model."""
    out: Dict[str, List[str]] = {}
    model_types = ['SPECIALCASE', 'TWOSTREAM_I3D', 'C3D', 'I3D', 'R3D_18', 'R3D_34', 'R3D_50', 'R3D_101', 'R3D_152']
    for model in models_list:
        for model_type in model_types:
            for training_condition in ('_PRETRAINED', '_SCRATCH'):
                if model_type + training_condition == model:
                    out.setdefault(model_type, []).append(training_condition)
    return out


def _existing(path):
    return path if os.path.isfile(path) else None


def lookFor_probabilitiesFile(nb_folds, results_folder, model_type, training_condition, classes_status,
                              optical_flow_status, augmentation_status, augmentation_frequency, involved_sets):
    stem = _stem(nb_folds, model_type, training_condition, classes_status, optical_flow_status, augmentation_status,
                 augmentation_frequency)
    return _existing(os.path.join(results_folder, involved_sets + "_predicted_probabilities_" + stem + ".csv"))


def lookFor_UniqueEnsemble_predictionsFile(nb_folds, results_folder, model_type, training_condition, classes_status,
                                           optical_flow_status, augmentation_status, augmentation_frequency):
    stem = _stem(nb_folds, model_type, training_condition, classes_status, optical_flow_status, augmentation_status,
                 augmentation_frequency)
    return _existing(os.path.join(results_folder, "weighted_prediction_results_" + stem + ".csv"))


def lookFor_GlobalEnsemble_predictionsFile(nb_folds, results_folder, models_list):
    return _existing(os.path.join(results_folder, "global_ensemble_summed_prediction_results_" + str(nb_folds) +
                                  "_folds_" + "_".join(models_list) + "_.csv"))


def get_modeltraining_validation_loss(histories_folder, test_index):
    """VALIDATION_ERROR_INVERSE weights: 1/min(val_loss) per member, normalised to sum 1 (:33-62)."""
    nb_folds = int(os.path.basename(histories_folder)[0])
    val_folds_indices = [i for i in range(nb_folds) if i != test_index]
    history_subfolder = os.path.join(histories_folder, "TestSplit" + str(test_index))
    histories_list = os.listdir(history_subfolder)
    weights = []
    for val_index in val_folds_indices:
        spec = "split_test" + str(test_index) + "_val" + str(val_index)
        fname = [h for h in histories_list if re.search(spec, h)][0]
        weights.append(1 / np.min(np.load(os.path.join(history_subfolder, fname))))
    weights = np.array(weights)
    return np.array(weights / np.sum(weights))


# --------------------------------------------------------------------------- #
# probabilities CSV <-> arrays (evaluate_ensemble.py:65-83, 1058-1063)
# --------------------------------------------------------------------------- #
_DTYPE_TAIL = re.compile(r",\s*dtype=float32\)")


def convert_str2array(raw_probabilities_str):
    """'[array([...], dtype=float32), ...]' -> float64 array.  Unlike the reference's plain
    str.replace, the dtype suffix is also recognised when numpy wrapped it onto its own line."""
    s = raw_probabilities_str.replace("array(", "")
    s = _DTYPE_TAIL.sub("", s)
    return np.array(ast.literal_eval(s))


def convert_array2listofarrays(probabilities_array):
    return [p for p in probabilities_array]


def _plain_ints(predictions):
    """Predictions column of the results CSVs: plain Python ints.  The reference's consumers
    ast.literal_eval this cell (evaluate_ensemble.py:424, 513, 553, 661); numpy >= 2 would print a list of
    np.int64 scalars as "[np.int64(1), ...]", which literal_eval rejects."""
    return [int(p) for p in predictions]


class _ProbabilityCache:
    """One parse per (file, mtime) instead of one per call (evaluate_ensemble.py:345)."""

    def __init__(self):
        self._files: Dict[str, Tuple[float, Dict[str, np.ndarray]]] = {}

    def table(self, path: str) -> Dict[str, np.ndarray]:
        key = os.path.abspath(path)
        mtime = os.path.getmtime(key)
        hit = self._files.get(key)
        if hit is None or hit[0] != mtime:
            df = pd.read_csv(key)
            tab = {p: convert_str2array(s) for p, s in zip(df["path"].values, df["probabilities"].values)}
            self._files[key] = (mtime, tab)
            hit = self._files[key]
        return hit[1]

    def member(self, path: str, model: str) -> np.ndarray:
        tab = self.table(path)
        key = os.path.splitext(model)[0]
        if key not in tab:
            raise KeyError("no probabilities stored for %s in %s" % (key, path))
        return tab[key]


_CACHE = _ProbabilityCache()


def _accuracy(y_true, y_pred) -> float:
    return float(np.mean(np.asarray(y_true) == np.asarray(y_pred)))


# --------------------------------------------------------------------------- #
# vote (evaluate_ensemble.py:86-100, 343-378)
# --------------------------------------------------------------------------- #
def _device_vote(yhats: np.ndarray, weights) -> np.ndarray:
    """[M,N,C] float64 -> int64 [N] through the CUDA vote kernel (no CPU fallback)."""
    from . import runtime as rt
    torch = rt.require_cuda()
    d = torch.from_numpy(np.ascontiguousarray(yhats, dtype=np.float64)).cuda()
    if isinstance(weights, str):
        pred = rt.vote(d, None, "MAXIMUM")
    else:
        pred = rt.vote(d, torch.from_numpy(np.asarray(weights, np.float64)), "WEIGHTED")
    return pred.cpu().numpy().astype(np.int64)


def evaluate_single_model(trained_model_path, test_labels, probabilities_file, nb_classes):
    yhat = np.reshape(_CACHE.member(probabilities_file, trained_model_path), (len(test_labels), nb_classes))
    predictions = _device_vote(yhat[None], np.ones(1))
    return _accuracy(test_labels, predictions), predictions


def ensemble_predictions(members, weights, testy, probabilities_file, nb_classes):
    yhats = np.array([_CACHE.member(probabilities_file, m) for m in members])
    yhats = np.reshape(yhats, (len(members), len(testy), nb_classes))
    if isinstance(weights, str):
        if weights == "MAXIMUM":
            return _device_vote(yhats, "MAXIMUM")
        print("Weights is %s, Unknown weights variable type.", weights)
        return None
    if isinstance(weights, np.ndarray):
        return _device_vote(yhats, weights)
    print("Unknown weights variable type.")
    return None


def evaluate_ensemble(members, weights, probabilities_file, testy, nb_classes):
    yhat = ensemble_predictions(members, weights, testy, probabilities_file, nb_classes)
    return _accuracy(testy, yhat), yhat


# --------------------------------------------------------------------------- #
# ensemble-weight search (evaluate_ensemble.py:282-339)
# --------------------------------------------------------------------------- #
def normalize(weights):
    result = np.linalg.norm(weights, 1)
    if result == 0.0:
        return weights
    return weights / result


def _search_scores(members, probabilities_file, testy, weight_matrix) -> np.ndarray:
    """Accuracy of every candidate weight vector (rows of weight_matrix) in one kernel launch."""
    from . import runtime as rt
    torch = rt.require_cuda()
    nb_classes = len(np.unique(testy))
    yhats = np.array([_CACHE.member(probabilities_file, m) for m in members])
    yhats = np.reshape(yhats, (len(members), len(testy), nb_classes))
    correct = rt.vote_search(torch.from_numpy(np.ascontiguousarray(yhats, np.float64)).cuda(),
                             torch.from_numpy(np.ascontiguousarray(weight_matrix, np.float64)).cuda(),
                             torch.from_numpy(np.asarray(testy, np.int32)).cuda())
    return correct.cpu().numpy() / float(len(testy))


def loss_function(weights, members, probabilities_file, testy):
    normalized = normalize(weights)
    nb_classes = len(np.unique(testy))
    return 1.0 - evaluate_ensemble(members, normalized, probabilities_file, testy, nb_classes)[0]


def apply_differential_evolution(n_members, members_paths, probabilities_file, testy):
    from scipy.optimize import differential_evolution

    def batched_loss(x):            # x: [n_members, S] (vectorized=True)
        w = np.asarray(x, np.float64).T
        norm = np.abs(w).sum(axis=1, keepdims=True)
        w = np.where(norm == 0.0, w, w / np.where(norm == 0.0, 1.0, norm))
        return 1.0 - _search_scores(members_paths, probabilities_file, testy, w)

    bound_w = [(0.0, 1.0) for _ in range(n_members)]
    result = differential_evolution(batched_loss, bound_w, maxiter=20, tol=1e-7, disp=True, vectorized=True,
                                    updating="deferred")
    return normalize(result['x'])


def grid_search(members, testX, testy):
    """All 11^M weight combinations (minus the all-equal ones), scored in one batched kernel; the
    first best in itertools.product order wins, as in the reference's sequential loop."""
    w = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]
    combos = np.array([c for c in product(w, repeat=len(members)) if len(set(c)) != 1], dtype=np.float64)
    norms = np.array([np.linalg.norm(c, 1) for c in combos])
    cand = combos / norms[:, None]
    scores = _search_scores(members, testX, testy, cand)
    best = int(np.argmax(scores))
    print('>%s %.3f' % (cand[best], scores[best]))
    return list(cand[best])


def apply_grid_search(members_paths, probabilities_file, testy):
    weights = grid_search(members_paths, probabilities_file, testy)
    print('Weights: %s' % weights)
    return np.array(weights)


# --------------------------------------------------------------------------- #
# Store_models_probabilities (evaluate_ensemble.py:1002-1109)
# --------------------------------------------------------------------------- #
def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return None, 0, 1


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Contiguous block partition of n clips over the ranks (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return np.arange(start, start + base + (1 if rank < extra else 0))


def shard_members(costs, world: int):
    """Member-sharded partition (alternative): which members each rank owns.  `costs` = one
    cost per member (FLOPs per clip; equal costs for a homogeneous ensemble).  Longest-processing-time
    greedy: members in decreasing cost go to the least loaded rank (ties -> lowest rank), which keeps
    heterogeneous ensembles (I3D-64 : C3D : R3D-34 = 16.7 : 5.8 : 1) balanced.  -> list of sorted index
    lists, one per rank; deterministic, identical on every rank."""
    costs = [float(c) for c in costs]
    load = [0.0] * world
    owned = [[] for _ in range(world)]
    for m in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].append(m)
        load[r] += costs[m]
    return [sorted(o) for o in owned]


class MemberGather:
    """All-gather of the per-rank probability blocks of a member-sharded ensemble into MEMBER order, so
    that the fixed-order fp64 vote that follows is bit-identical to the single-process one (an
    all-reduce would make the summation order depend on the ring).  Buffers and the re-ordering index
    are allocated once and reused every step (no host <-> device traffic in the step)."""

    def __init__(self, owned, m_total: int, dist, world: int):
        self.owned, self.m_total, self.dist, self.world = owned, int(m_total), dist, int(world)
        self.pad = max(len(o) for o in owned)
        self._key = None

    def _prepare(self, local):
        import torch
        key = (tuple(local.shape[1:]), local.dtype, local.device)
        if key == self._key:
            return
        n, c = local.shape[1], local.shape[2]
        self.buf = torch.zeros((self.pad, n, c), dtype=local.dtype, device=local.device)
        self.out = torch.empty((self.world * self.pad, n, c), dtype=local.dtype, device=local.device)
        index = torch.empty((self.m_total,), dtype=torch.long)
        for r, ids in enumerate(self.owned):
            for k, m in enumerate(ids):
                index[m] = r * self.pad + k
        self.index = index.to(local.device)
        self.full = torch.empty((self.m_total, n, c), dtype=local.dtype, device=local.device)
        self._key = key

    def __call__(self, local):
        """local: torch tensor [len(owned[rank]), N, C] -> [m_total, N, C]."""
        import torch
        if self.world == 1:
            return local
        self._prepare(local)
        self.buf[:local.shape[0]].copy_(local)
        if self.dist.get_backend() == "nccl":
            self.dist.all_gather_into_tensor(self.out, self.buf)
            out = self.out
        else:
            parts = [torch.empty_like(self.buf) for _ in range(self.world)]
            self.dist.all_gather(parts, self.buf)
            out = torch.cat(parts, dim=0)
        torch.index_select(out, 0, self.index, out=self.full)
        return self.full


def shard_units(costs, n_clips: int, world: int, chunk: int = 0):
    """2-D partition of an ensemble step over the ranks: work units are (member, clip chunk) pairs, so that heavy
    members (one I3D-64 member is 17.8 % of the global C3D + I3D + R3D-34 step) are split across GPUs instead of
    bounding the step.  Every member's clips are cut into chunks of `chunk` clips (default ceil(n_clips / world));
    units go, in decreasing cost (cost = member FLOPs per clip x clips; ties: member, then chunk order), to the
    least loaded rank (ties -> lowest rank).  With members x chunks that divide evenly this degenerates to plain
    clip sharding (rank r runs chunk r of every member).  -> one list of (member, lo, hi) per rank, sorted;
    deterministic, identical on every rank."""
    costs = [float(c) for c in costs]
    chunk = int(chunk) if chunk else -(-int(n_clips) // int(world))
    chunk = max(1, chunk)
    units = [(m, lo, min(lo + chunk, n_clips)) for m in range(len(costs)) for lo in range(0, n_clips, chunk)]
    load = [0.0] * world
    owned = [[] for _ in range(world)]
    for u in sorted(units, key=lambda u: (-costs[u[0]] * (u[2] - u[1]), u[0], u[1])):
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].append(u)
        load[r] += costs[u[0]] * (u[2] - u[1])
    return [sorted(o) for o in owned]


class UnitGather:
    """Merges the per-rank probability blocks of a unit-sharded step (shard_units) into the full [M, N, C] block on
    every rank.  Each (member, clip) row is produced by exactly one rank and is zero everywhere else, so ONE
    all-reduce(SUM) of the fp32 block is exact in any reduction order (x + 0 + ... + 0 = x bit for bit; probabilities
    are positive).  The soft vote that follows still runs over all M members in member order in fp64, identical to
    the single-process vote.  The buffer is allocated once; no host <-> device traffic in the step."""

    def __init__(self, units, m_total: int, n_clips: int, nb_classes: int, dist, world: int, device):
        import torch
        self.units, self.dist, self.world = units, dist, int(world)
        seen = np.zeros((m_total, n_clips), np.int32)
        for owned in units:
            for m, lo, hi in owned:
                seen[m, lo:hi] += 1
        if not (seen == 1).all():
            raise ValueError("units must cover every (member, clip) exactly once")
        self.full = torch.zeros((m_total, n_clips, nb_classes), dtype=torch.float32, device=device)

    def __call__(self, local):
        """local: [M, N, C] fp32 with this rank's units filled in and zeros elsewhere -> the full block."""
        if self.world == 1:
            return local
        self.full.copy_(local)
        self.dist.all_reduce(self.full)
        return self.full


def gather_member_probs(local, owned, m_total: int, dist, world: int):
    """One-shot form of MemberGather."""
    return MemberGather(owned, m_total, dist, world)(local)


def _gather_rows(local: np.ndarray, n: int, dist, rank: int, world: int) -> np.ndarray:
    """All-gather of the per-rank [M, n_local, C] probability blocks into [M, n, C] (rank order =
    clip order because shards are contiguous)."""
    if world == 1:
        return local
    import torch
    sizes = [len(shard_indices(n, r, world)) for r in range(world)]
    m, _, c = local.shape
    pad = max(sizes)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((m, pad, c), dtype=torch.float32, device=dev)
    buf[:, :local.shape[1]] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:, :s].cpu().numpy() for o, s in zip(out, sizes)], axis=1)


def _gather_device(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _clip_device():
    """Where clips are assembled (select_frames + resize): on the GPU that runs the members
    (cse_assemble_clip, bit-identical to cv2.resize) unless CSE_CPU_RESIZE=1 asks for the cv2 path."""
    if os.environ.get("CSE_CPU_RESIZE", "0") == "1":
        return None
    import torch
    return torch.device("cuda", torch.cuda.current_device())


class _NoMembers:
    """A rank that owns no member of a member-sharded ensemble."""
    M = 0

    def __init__(self, nb_classes):
        import torch
        self.nb_classes, self.device = nb_classes, torch.device("cpu")
        self.probs = torch.zeros((0, 1, nb_classes), dtype=torch.float32)

    def forward_members(self, inputs):
        return inputs[0].shape[0]


def _load_members(model_type, member_paths, input_shape, nb_classes, batch, optical_flow_status="TVL1_precomputed"):
    """All members of one test fold on this GPU, sharing one activation workspace."""
    from .ensemble_runtime import DeviceEnsemble
    from .graph import build_model_graph
    g = build_model_graph(model_type, tuple(input_shape), nb_classes)
    weight_sets = [zoo.load_member_weights(g, p) for p in member_paths]
    kw = {}
    if model_type == "TWOSTREAM_I3D" and optical_flow_status == "FarneBack_onTheFly":
        kw["input_dtypes"] = ("u8", "f32")        # the on-the-fly flow is a float32 volume (train.py:294-332)
    return DeviceEnsemble(g, weight_sets, precision=zoo.DEFAULTS["precision"], max_batch=batch, micro_batch=batch, **kw)


def _predict_members(ens, generator, n_clips, dist_state, chunk, owned=None, m_total=None, workers=1):
    """-> float32 [M, n_clips, C]: every clip decoded once, pushed through every member.
    Default partition: clips sharded over the ranks, members replicated.  owned = shard_members(...)
    switches to the member-sharded partition: `ens` holds only this rank's members, every rank runs all
    clips, and the probability blocks are all-gathered into member order."""
    import torch
    dist, rank, world = dist_state
    if owned is not None and world > 1:
        local = _predict_members(ens, generator, n_clips, (None, 0, 1), chunk, workers=workers)
        full = gather_member_probs(torch.from_numpy(local).to(_gather_device(dist)), owned, m_total, dist, world)
        return full.cpu().numpy()
    mine = shard_indices(n_clips, rank, world)
    bs = generator.batch_size
    out = np.zeros((ens.M, len(mine), ens.nb_classes), np.float32)
    pos = 0
    pend, pend_n = [], 0

    def flush():
        nonlocal pos, pend, pend_n
        if not pend:
            return
        ninp = len(pend[0])
        dev = [torch.cat([p[j] for p in pend]).to(ens.device).contiguous() if torch.is_tensor(pend[0][j])
               else torch.from_numpy(np.ascontiguousarray(np.concatenate([p[j] for p in pend]))).to(ens.device)
               for j in range(ninp)]
        n = ens.forward_members(dev)
        out[:, pos:pos + n] = ens.probs[:, :n].cpu().numpy()
        pos += n
        pend, pend_n = [], 0

    first_batch, last_batch = (mine[0] // bs, mine[-1] // bs) if len(mine) else (0, -1)
    from .clips import iterate_batches
    for b, (x, _) in zip(range(first_batch, last_batch + 1),
                         iterate_batches(generator, first_batch, last_batch, workers)):
        xs = x if isinstance(x, (list, tuple)) else [x]
        lo = b * bs
        keep = [i for i in range(xs[0].shape[0]) if mine[0] <= lo + i <= mine[-1]]
        if not keep:
            continue
        xs = [v[keep] if torch.is_tensor(v) else np.asarray(v)[keep] for v in xs]
        while xs[0].shape[0]:
            take = min(chunk - pend_n, xs[0].shape[0])
            pend.append([v[:take] for v in xs])
            pend_n += take
            xs = [v[take:] for v in xs]
            if pend_n == chunk:
                flush()
    flush()
    return _gather_rows(out, n_clips, dist, rank, world)


def store_probabilities(trained_models_folder, results_folder, involved_sets, batch_size, workers, model_type,
                        training_condition, optical_flow_status, augmentation_status, augmentation_frequency,
                        classes_status, models_name):
    """Predicts every member's probabilities for the test (or train+val) set of every fold and stores
    them in ``<results_folder>/{test|train_val}_predicted_probabilities_<models_name>.csv``."""
    nb_folds = int(os.path.basename(trained_models_folder)[0])
    test_folds_indices = list(range(0, nb_folds))
    rows = []
    dist_state = _dist()
    rank = dist_state[1]
    if not os.path.exists(results_folder):
        os.makedirs(results_folder, exist_ok=True)
    chunk = int(batch_size) if batch_size and int(batch_size) > 1 else zoo.DEFAULTS["max_batch"]
    for test_index in test_folds_indices:
        data_folder = os.path.join(trained_models_folder, "TestSplit" + str(test_index))
        if involved_sets == "test":
            data = pd.read_csv(os.path.join(data_folder, 'test.csv'))
        else:
            data = pd.concat([pd.read_csv(os.path.join(data_folder, 'train.csv')),
                              pd.read_csv(os.path.join(data_folder, 'val.csv'))], ignore_index=True)
        nb_classes = len(set(data['class']))
        sample_input = zoo.define_input(model_type)
        generator = ClipSequence(data, model_type, sample_input.shape, nb_classes, batch_size=1,
                                 optical_flow_status=optical_flow_status, augmentation_status="non_augmented",
                                 augmentation_frequency=0, shuffle=False, device=_clip_device())
        val_folds_indices = [i for i in test_folds_indices if i != test_index]
        member_paths = [os.path.join(data_folder, models_name + "_split_test" + str(test_index) + "_val" +
                                     str(v) + "_weights.hdf5") for v in val_folds_indices]
        owned = None
        if dist_state[2] > 1 and os.environ.get("CSE_SHARD", "clips") == "members":
            # member-sharded partition: this rank loads and runs only its own members on all clips
            owned = shard_members([1.0] * len(member_paths), dist_state[2])
        local_paths = member_paths if owned is None else [member_paths[m] for m in owned[rank]]
        ens = _load_members(model_type, local_paths, sample_input.shape, nb_classes, chunk,
                            optical_flow_status) if local_paths else None
        if ens is None:         # more ranks than members: nothing to run here, but take part in the gather
            ens = _NoMembers(nb_classes)
        probs = _predict_members(ens, generator, generator.n, dist_state, chunk, owned, len(member_paths),
                                 workers=int(workers) if workers else 1)
        del ens
        for j, path in enumerate(member_paths):
            print(probs[j].shape)
            rows.append([os.path.splitext(path)[0], convert_array2listofarrays(probs[j])])
    prefix = "test" if involved_sets == "test" else "train_val"
    csv_file_path = os.path.join(results_folder, prefix + "_predicted_probabilities_" + models_name + ".csv")
    if rank == 0:
        pd.DataFrame(rows, columns=["path", "probabilities"]).to_csv(csv_file_path)
    if dist_state[0] is not None:
        dist_state[0].barrier()
    return csv_file_path


# --------------------------------------------------------------------------- #
# Evaluate_ensembles (evaluate_ensemble.py:1112-1273)
# --------------------------------------------------------------------------- #
def evaluate_ensembles(trained_models_folder, results_folder, weights_type, histories_folder, test_probabilities_file,
                       trainval_probabilities_file, weights_array_file, batch_size, workers, model_type,
                       training_condition, optical_flow_status, augmentation_status, augmentation_frequency,
                       classes_status, models_name):
    nb_folds = int(os.path.basename(trained_models_folder)[0])
    test_folds_indices = list(range(0, nb_folds))
    store_models_predictions = []
    optimization_weights = []
    if not os.path.exists(results_folder):
        os.makedirs(results_folder, exist_ok=True)

    def _store(sets):
        return store_probabilities(trained_models_folder, results_folder, sets, batch_size, workers, model_type,
                                   training_condition, optical_flow_status, augmentation_status,
                                   augmentation_frequency, classes_status, models_name)

    if test_probabilities_file is None:
        test_probabilities_file = _store("test")
    for test_index in test_folds_indices:
        data_folder = os.path.join(trained_models_folder, "TestSplit" + str(test_index))
        test_data = pd.read_csv(os.path.join(data_folder, 'test.csv'))
        trainval_data = pd.concat([pd.read_csv(os.path.join(data_folder, 'train.csv')),
                                   pd.read_csv(os.path.join(data_folder, 'val.csv'))], ignore_index=True)
        nb_classes = len(set(test_data['class']))
        test_labels = test_data['class'].values
        trainval_labels = trainval_data['class'].values
        val_folds_indices = [i for i in test_folds_indices if i != test_index]
        trained_model_paths = []
        for val_index in val_folds_indices:
            model_file = models_name + "_split_test" + str(test_index) + "_val" + str(val_index) + "_weights"
            trained_model_path = os.path.join(data_folder, model_file)
            accuracy, single_model_predictions = evaluate_single_model(trained_model_path, test_labels,
                                                                       test_probabilities_file, nb_classes)
            print("Model val %d : %f" % (val_index, accuracy))
            store_models_predictions.append([trained_model_path, _plain_ints(single_model_predictions)])
            trained_model_paths.append(trained_model_path)
        ensemble_models_number = nb_folds - 1
        weights = None
        if weights_type in ("GRID_SEARCH", "DIFFERENTIAL_EVOLUTION"):
            if trainval_probabilities_file is None:
                trainval_probabilities_file = _store("train_val")
            if weights_array_file is None:
                if weights_type == "GRID_SEARCH":
                    optimization_weights.append(apply_grid_search(trained_model_paths, trainval_probabilities_file,
                                                                  trainval_labels))
                else:
                    optimization_weights.append(apply_differential_evolution(
                        ensemble_models_number, trained_model_paths, trainval_probabilities_file, trainval_labels))
                weights = optimization_weights[test_index]
            else:
                weights = np.load(weights_array_file)[test_index]
        elif weights_type == "SUM":
            weights = np.ones(ensemble_models_number)
        elif weights_type == "VALIDATION_ERROR_INVERSE":
            weights = get_modeltraining_validation_loss(histories_folder, test_index)
        elif weights_type == "MAXIMUM":
            weights = weights_type
        else:
            print("Unknown weighting method.")
        ensemble_model_accuracy, ensemble_model_predictions = evaluate_ensemble(
            trained_model_paths, weights, test_probabilities_file, test_labels, nb_classes)
        print("Fold %d : %f" % (test_index, ensemble_model_accuracy))
        ensemble_model_name = "Ensemble_" + models_name + "_split_test" + str(test_index)
        store_models_predictions.append([ensemble_model_name, _plain_ints(ensemble_model_predictions)])
    csv_file_path = os.path.join(results_folder, "weighted_prediction_results_" + models_name + ".csv")
    if _dist()[1] == 0:
        pd.DataFrame(store_models_predictions, columns=["path", "predictions"]).to_csv(csv_file_path)
        if weights_type in ("GRID_SEARCH", "DIFFERENTIAL_EVOLUTION"):
            np.save(weights_type + "_" + models_name + ".npy", np.array(optimization_weights))
    return csv_file_path


# --------------------------------------------------------------------------- #
# Global_evaluate_models / Combine_ensembles (evaluate_ensemble.py:1280-1474)
# --------------------------------------------------------------------------- #
def compute_combinations(models_list):
    """Every non-empty subset of models_list (evaluate_ensemble.py:1280-1295).  The reference builds each size
    class through set(), whose order depends on per-process string hashing; under torchrun every rank must walk
    the combinations in the SAME order (they meet in collectives inside global_evaluate_ensembles), so the
    deterministic itertools order is kept (duplicates in models_list are dropped, as set() would)."""
    models_list = list(dict.fromkeys(models_list))
    combinations = []
    for k in range(1, len(models_list) + 1):
        combinations.append(list(itertools.combinations(models_list, k)))
    combinations = list(itertools.chain.from_iterable(combinations))
    return len(combinations), combinations


def combine_ensembles(nb_folds, trained_models_parent_folder, models_list, results_folder):
    _, combinations = compute_combinations(models_list)
    acc = {c: global_evaluate_ensembles(nb_folds, trained_models_parent_folder, c, results_folder)
           for c in combinations}
    ordered = dict(sorted(acc.items(), key=lambda item: item[1], reverse=True))
    for combination, accuracy in ordered.items():
        print(combination, accuracy)
    return ordered


def global_evaluate_ensembles(nb_folds, trained_models_parent_folder, models_list, results_folder):
    test_folds_indices = list(range(0, nb_folds))
    store_models_predictions, store_models_accuracies = [], []
    if not os.path.exists(results_folder):
        os.makedirs(results_folder, exist_ok=True)
    model_trainingConditions = createModelsTrainingConditionsDictionary(models_list)
    print(model_trainingConditions)
    all_models_names_string = ""
    for test_index in test_folds_indices:
        all_models_names = []
        frames = []
        trained_model_paths = []
        ensemble_models_number = 0
        for model_type in list(model_trainingConditions.keys()):
            for training_condition in model_trainingConditions[model_type]:
                if model_type + training_condition == "SPECIALCASE_PRETRAINED":
                    all_models_names.append(
                        "TWOSTREAM_I3D_PRETRAINED_OF_FarneBack_onTheFly_AS_augmented_precomputed_Freq3")
                    mt, tc = "TWOSTREAM_I3D", "_PRETRAINED"
                    look = dict(classes_status="unbalanced", optical_flow_status="FarneBack_onTheFly",
                                augmentation_status="augmented_precomputed", augmentation_frequency=3)
                else:
                    all_models_names.append(model_type + training_condition)
                    mt, tc = model_type, training_condition
                    look = dict(classes_status="unbalanced", optical_flow_status="TVL1_precomputed",
                                augmentation_status="non_augmented", augmentation_frequency=0)
                test_probabilities_file = lookFor_probabilitiesFile(nb_folds, results_folder, mt, tc,
                                                                    involved_sets="test", **look)
                models_name, trained_models_subfolder = get_ModelsNameAndTrainedModelsSubfolder(
                    nb_folds, trained_models_parent_folder, mt, tc, **look)
                if test_probabilities_file is None:
                    # the reference recomputes with the TV-L1 / non-augmented settings here (:1408-1419)
                    test_probabilities_file = store_probabilities(
                        trained_models_subfolder, results_folder, "test", batch_size=1, workers=1, model_type=mt,
                        training_condition=tc, optical_flow_status="TVL1_precomputed",
                        augmentation_status="non_augmented", augmentation_frequency=0, classes_status="unbalanced",
                        models_name=models_name)
                frames.append(pd.read_csv(test_probabilities_file))
                ensemble_models_number += 1
                data_folder = os.path.join(trained_models_subfolder, "TestSplit" + str(test_index))
                for val_index in [i for i in test_folds_indices if i != test_index]:
                    model_file = models_name + "_split_test" + str(test_index) + "_val" + str(val_index) + "_weights"
                    trained_model_paths.append(os.path.join(data_folder, model_file))
        all_models_names_string = "_".join(all_models_names)
        merged = pd.concat(frames, ignore_index=True, sort=False)[["path", "probabilities"]]
        merged_file = os.path.join(results_folder, "global_ensemble_probabilities_" + all_models_names_string +
                                   "_TestFold" + str(test_index) + "_" + str(nb_folds) + "folds.csv")
        if _dist()[1] == 0:
            merged.to_csv(merged_file)
        if _dist()[0] is not None:
            _dist()[0].barrier()
        first_models_folder = os.path.dirname(os.path.dirname(trained_model_paths[0]))
        test_data = pd.read_csv(os.path.join(first_models_folder, "TestSplit" + str(test_index), 'test.csv'))
        nb_classes = len(set(test_data['class']))
        test_labels = test_data['class'].values
        weights = np.ones(ensemble_models_number * (nb_folds - 1))
        accuracy, predictions = evaluate_ensemble(trained_model_paths, weights, merged_file, test_labels, nb_classes)
        store_models_accuracies.append(accuracy)
        print("Fold %d : %f" % (test_index, accuracy))
        store_models_predictions.append(["Global_Ensemble_" + all_models_names_string + "_split_test" +
                                         str(test_index), _plain_ints(predictions)])
    csv_file_path = os.path.join(results_folder, "global_ensemble_summed_prediction_results_" + str(nb_folds) +
                                 "_folds_" + all_models_names_string + "_.csv")
    if _dist()[1] == 0:
        pd.DataFrame(store_models_predictions, columns=["path", "predictions"]).to_csv(csv_file_path)
    return np.mean(np.array(store_models_accuracies))


# --------------------------------------------------------------------------- #
# main / argparse (evaluate_ensemble.py:1481-1796)
# --------------------------------------------------------------------------- #
OPERATIONS = ['Confusion_matrices', 'Difference_matrices', 'Evaluate_ensembles', 'Store_models_probabilities',
              'StickDiagrams_wellClassifiedClips_per_numberOfModels', 'Global_evaluate_models', 'Combine_ensembles']
_PLOTS = ('Confusion_matrices', 'Difference_matrices', 'StickDiagrams_wellClassifiedClips_per_numberOfModels')


def main(args):
    try:
        print(args.operation)
        if getattr(args, "precision", None):
            zoo.DEFAULTS["precision"] = args.precision
        if args.operation in _PLOTS:
            print("Operation %s is a matplotlib report outside the accelerated hot path; use the reference's "
                  "plotting code on the CSV files written by this tool." % args.operation)
            return
        if args.operation in ("Evaluate_ensembles", "Store_models_probabilities"):
            models_name, trained_models_subfolder = get_ModelsNameAndTrainedModelsSubfolder(
                args.folds_number, args.trained_models_folder, args.model_type, args.training_condition,
                args.classes_status, args.optical_flow_status, args.augmentation_status, args.augmentation_frequency)
        if args.operation == "Evaluate_ensembles":
            look = (args.folds_number, args.results_folder, args.model_type, args.training_condition,
                    args.classes_status, args.optical_flow_status, args.augmentation_status,
                    args.augmentation_frequency)
            evaluate_ensembles(trained_models_subfolder, args.results_folder, args.weights_type,
                               os.path.join(args.historiesFolder_validationErrorInverse,
                                            os.path.basename(trained_models_subfolder)),
                               lookFor_probabilitiesFile(*look, involved_sets="test"),
                               lookFor_probabilitiesFile(*look, involved_sets="train_val"),
                               args.weights_array_file, args.batch_size, args.workers, args.model_type,
                               args.training_condition, args.optical_flow_status, args.augmentation_status,
                               args.augmentation_frequency, args.classes_status, models_name)
        elif args.operation == "Store_models_probabilities":
            store_probabilities(trained_models_subfolder, args.results_folder, args.involved_sets, args.batch_size,
                                args.workers, args.model_type, args.training_condition, args.optical_flow_status,
                                args.augmentation_status, args.augmentation_frequency, args.classes_status,
                                models_name)
        elif args.operation in ("Global_evaluate_models", "Combine_ensembles"):
            print(args.models_list)
            print("Folds number : " + str(args.folds_number))
            print("Results folder : " + args.results_folder)
            print("Trained models folder : " + args.trained_models_folder)
            fn = global_evaluate_ensembles if args.operation == "Global_evaluate_models" else combine_ensembles
            fn(args.folds_number, args.trained_models_folder, args.models_list, args.results_folder)
        else:
            print("Operation not mentioned")
    except Exception as err:                       # same behaviour as the reference: report and continue
        print('Error:', err)
        traceback.print_tb(err.__traceback__)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument('-op', '--operation', type=str, choices=OPERATIONS, required=True)
    p.add_argument('-et', '--ensemble_type', type=str, choices=['Unique', 'Global'], required=False)
    p.add_argument('-mlist', '--models_list', nargs='+', required=False)
    p.add_argument('-fn', '--folds_number', type=int)
    p.add_argument('-tmf', '--trained_models_folder', type=str)
    p.add_argument('-rf', '--results_folder', type=str, default="Results", required=True)
    p.add_argument('-is', '--involved_sets', type=str, default='test', choices=['train_val', 'test'])
    p.add_argument('-prf', '--prediction_results_file', type=str)
    p.add_argument('-wt', '--weights_type', type=str,
                   choices=['GRID_SEARCH', 'DIFFERENTIAL_EVOLUTION', 'SUM', 'VALIDATION_ERROR_INVERSE', 'MAXIMUM'])
    p.add_argument('-wf', '--weights_array_file', type=str)
    p.add_argument('-hf_vei', '--historiesFolder_validationErrorInverse', default="Data/Weights", type=str)
    p.add_argument('-mt', '--model_type', type=str,
                   choices=['TWOSTREAM_I3D', 'I3D', 'C3D', 'R3D_18', 'R3D_34', 'R3D_50', 'R3D_101', 'R3D_152'])
    p.add_argument('-tc', '--training_condition', type=str, choices=['_SCRATCH', '_PRETRAINED'])
    p.add_argument('-as', '--augmentation_status', type=str, default='non_augmented',
                   choices=['non_augmented', 'augmented_onTheFly', 'augmented_precomputed'])
    p.add_argument('-af', '--augmentation_frequency', type=int, default=0)
    p.add_argument('-ofs', '--optical_flow_status', type=str, choices=['TVL1_precomputed', 'FarneBack_onTheFly'])
    p.add_argument('-cs', '--classes_status', type=str, default='unbalanced', choices=['balanced', 'unbalanced'])
    p.add_argument('-w', '--workers', type=int)
    p.add_argument('-b', '--batch_size', type=int)
    # B200-path additions (defaults reproduce the reference's behaviour at bf16 tensor-core speed)
    p.add_argument('--precision', type=str, default=None, choices=['bf16', 'fp32'],
                   help='bf16 = tcgen05 tensor-core path (default); fp32 = CUDA-core reference-precision path')
    return p
