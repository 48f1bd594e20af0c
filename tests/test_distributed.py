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


# --------------------------------------------------------------------------- member-sharded partition
def test_shard_members_balances_by_cost():
    """LPT greedy over FLOPs per clip: the global C3D + I3D-64 + R3D-34 ensemble (4 members each,
    77.1 / 222.3 / 13.3 GFLOP) on 2, 4 and 8 ranks; every member owned exactly once, deterministic."""
    costs = [77.1] * 4 + [222.3] * 4 + [13.3] * 4
    for world in (1, 2, 3, 4, 8, 12):
        owned = E.shard_members(costs, world)
        assert len(owned) == world and sorted(m for o in owned for m in o) == list(range(12))
        assert owned == E.shard_members(costs, world)
        load = [sum(costs[m] for m in o) for o in owned]
        assert max(load) <= sum(costs) / world + max(costs)          # LPT bound
    # 4 ranks: each gets one I3D member; the C3D / R3D members fill up the lightest ranks
    assert all(sum(1 for m in o if 4 <= m < 8) == 1 for o in E.shard_members(costs, 4))
    assert E.shard_members([1.0] * 4, 2) == [[0, 2], [1, 3]]


def _member_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owned = E.shard_members([3.0, 1.0, 2.0, 5.0, 1.0], world)

        class Local(FakeEnsemble):
            """Only this rank's members of the 5-member stand-in ensemble."""
            def __init__(self):
                super().__init__()
                self.ids = owned[rank]
                self.M = len(self.ids)
                self.probs = torch.zeros((self.M, 64, C), dtype=torch.float32)

            def forward_members(self, inputs):
                x = inputs[0].to(torch.float32)
                n = x.shape[0]
                feat = x.reshape(n, -1)
                for k, m in enumerate(self.ids):
                    logits = torch.stack([(feat[:, (c + m)::C]).mean(dim=1) * (1 + 0.01 * c) for c in range(C)], dim=1)
                    self.probs[k, :n] = torch.softmax(logits / 16.0, dim=1)
                return n
        got = E._predict_members(Local(), FakeSequence(N_CLIPS), N_CLIPS, (dist, rank, world), chunk=4,
                                 owned=owned, m_total=5)
        np.save(os.path.join(out_dir, "mrank%d.npy" % rank), got)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_member_sharded_probabilities_match_single_process(tmp_path, world):
    """Members sharded over the ranks (every rank runs all clips through its own members), one
    all-gather into MEMBER order: every rank must hold the single-process [M, N, C] block."""
    class Five(FakeEnsemble):
        def __init__(self):
            super().__init__()
            self.M = 5
            self.probs = torch.zeros((5, 64, C), dtype=torch.float32)

        def forward_members(self, inputs):
            x = inputs[0].to(torch.float32)
            n = x.shape[0]
            feat = x.reshape(n, -1)
            for m in range(5):
                logits = torch.stack([(feat[:, (c + m)::C]).mean(dim=1) * (1 + 0.01 * c) for c in range(C)], dim=1)
                self.probs[m, :n] = torch.softmax(logits / 16.0, dim=1)
            return n
    ref = E._predict_members(Five(), FakeSequence(N_CLIPS), N_CLIPS, (None, 0, 1), chunk=4)
    mp.spawn(_member_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(str(tmp_path / ("mrank%d.npy" % r)))
        assert got.shape == ref.shape == (5, N_CLIPS, C)
        assert np.array_equal(got, ref), "rank %d: member-sharded gather differs" % r


# --------------------------------------------------------------------------- (member, clip-chunk) units
def test_shard_units_balances_and_covers():
    """LPT over (member, clip chunk) units: the global C3D + I3D-64 + R3D-34 ensemble (4 members each) on 256 clips;
    every (member, clip) owned exactly once; load within one unit of perfect; 8 ranks = plain clip sharding."""
    costs = [77.1] * 4 + [222.3] * 4 + [13.3] * 4
    for world in (1, 2, 3, 5, 8):
        units = E.shard_units(costs, 256, world)
        assert units == E.shard_units(costs, 256, world)
        seen = np.zeros((12, 256), int)
        for owned in units:
            for m, lo, hi in owned:
                seen[m, lo:hi] += 1
        assert (seen == 1).all()
        load = [sum(costs[m] * (hi - lo) for m, lo, hi in o) for o in units]
        biggest = max(costs) * -(-256 // world)
        assert max(load) - min(load) <= biggest
        assert max(load) <= sum(costs) * 256 / world + biggest
    u8 = E.shard_units(costs, 256, 8)
    assert all(sorted(o) == [(m, 32 * r, 32 * r + 32) for m in range(12)] for r, o in enumerate(u8))
    # ragged: 23 clips on 3 ranks, one heavy member is split over all ranks
    u3 = E.shard_units([10.0, 1.0], 23, 3)
    assert all(any(m == 0 for m, _, _ in o) for o in u3)


def _unit_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [3.0, 1.0, 2.0, 5.0, 1.0]
        units = E.shard_units(costs, N_CLIPS, world)
        seq = FakeSequence(N_CLIPS)
        x = torch.from_numpy(seq.clips).to(torch.float32)
        feat = x.reshape(N_CLIPS, -1)
        local = torch.zeros((5, N_CLIPS, C), dtype=torch.float32)
        for m, lo, hi in units[rank]:                       # only this rank's (member, clip-chunk) units
            logits = torch.stack([(feat[lo:hi, (c + m)::C]).mean(dim=1) * (1 + 0.01 * c) for c in range(C)], dim=1)
            local[m, lo:hi] = torch.softmax(logits / 16.0, dim=1)
        gather = E.UnitGather(units, 5, N_CLIPS, C, dist, world, torch.device("cpu"))
        np.save(os.path.join(out_dir, "urank%d.npy" % rank), gather(local).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_unit_sharded_probabilities_match_single_process(tmp_path, world):
    """(member, clip-chunk) units sharded over the ranks, merged by one all-reduce of disjoint blocks: every rank must
    hold the single-process [M, N, C] block bit for bit, hence vote identically."""
    seq = FakeSequence(N_CLIPS)
    feat = torch.from_numpy(seq.clips).to(torch.float32).reshape(N_CLIPS, -1)
    ref = np.stack([torch.softmax(torch.stack([(feat[:, (c + m)::C]).mean(dim=1) * (1 + 0.01 * c) for c in range(C)],
                                              dim=1) / 16.0, dim=1).numpy() for m in range(5)])
    mp.spawn(_unit_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(str(tmp_path / ("urank%d.npy" % r)))
        assert np.array_equal(got, ref), "rank %d: unit-sharded merge differs" % r
