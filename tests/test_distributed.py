"""N > 1 host logic on the CPU: two `gloo` ranks shard the clips of a fold, run a stand-in
ensemble on their shard, all-gather the per-clip probabilities and must reproduce the
single-process result bit for bit, in clip order (clips sharded, members replicated,
one all-gather).  The device ensemble is replaced by a deterministic CPU stand-in - this test covers
sharding / gathering / ordering, not arithmetic (that is what the `-m gpu` tests are for)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cse_b200 import ensemble as E

N_CLIPS, M, C = 23, 3, 11


class FakeSequence:
    """keras.utils.Sequence protocol with batch_size 1 (what store_probabilities builds)."""
    batch_size = 1

    def __init__(self, n):
        self.n = n
        rng = np.random.default_rng(5)
        self.clips = rng.integers(0, 256, (n, 2, 4, 4, 3), dtype=np.uint8)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.clips[i:i + 1], np.zeros((1, C), np.float32)


class FakeEnsemble:
    """Stand-in for DeviceEnsemble: probabilities are a deterministic function of the clip bytes."""
    def __init__(self):
        self.M, self.nb_classes, self.device = M, C, torch.device("cpu")
        self.probs = torch.zeros((M, 64, C), dtype=torch.float32)
        self.calls = 0

    def forward_members(self, inputs):
        x = inputs[0].to(torch.float32)
        n = x.shape[0]
        feat = x.reshape(n, -1)
        for m in range(M):
            logits = torch.stack([(feat[:, (c + m)::C]).mean(dim=1) * (1 + 0.01 * c) for c in range(C)], dim=1)
            self.probs[m, :n] = torch.softmax(logits / 16.0, dim=1)
        self.calls += 1
        return n


def _single_process_reference():
    return E._predict_members(FakeEnsemble(), FakeSequence(N_CLIPS), N_CLIPS, (None, 0, 1), chunk=4)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ens = FakeEnsemble()
        got = E._predict_members(ens, FakeSequence(N_CLIPS), N_CLIPS, (dist, rank, world), chunk=4)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), got)
        np.save(os.path.join(out_dir, "calls%d.npy" % rank), np.array([ens.calls]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_probabilities_match_single_process(tmp_path, world):
    ref = _single_process_reference()
    assert ref.shape == (M, N_CLIPS, C)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    total_calls = 0
    for r in range(world):
        got = np.load(str(tmp_path / ("rank%d.npy" % r)))
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), "rank %d gathered probabilities differ from the single-process run" % r
        total_calls += int(np.load(str(tmp_path / ("calls%d.npy" % r)))[0])
    # every rank only ran its own shard: ceil(shard / chunk) forward calls each
    expect = sum(-(-len(E.shard_indices(N_CLIPS, r, world)) // 4) for r in range(world))
    assert total_calls == expect


def test_vote_after_gather_is_rank_independent():
    """The vote runs on the gathered [M, N, C] block, so every rank votes on identical bytes; the
    fixed member order keeps the fp64 sum equal to np.tensordot (oracle)."""
    from oracle import vote as OV
    ref = _single_process_reference().astype(np.float64)
    parts = [ref[:, E.shard_indices(N_CLIPS, r, 2)] for r in range(2)]
    glued = np.concatenate(parts, axis=1)
    assert np.array_equal(glued, ref)
    assert np.array_equal(OV.ensemble_predictions(glued, np.ones(M)), OV.ensemble_predictions(ref, np.ones(M)))
