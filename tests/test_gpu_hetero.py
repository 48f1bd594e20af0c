"""The class bench.py times: HeteroEnsemble (global heterogeneous ensemble of global_evaluate_ensembles,
evaluate_ensemble.py:1329-1474, weights = ones over all members :1455) - resident path, pipelined host path and the
per-member path + oracle vote must give bit-identical predictions; the member-sharded gather on NCCL."""
import os
import socket

import numpy as np
import pytest
import torch

from cse_b200 import graph as G
from cse_b200.ensemble_runtime import HeteroEnsemble
from cse_b200.model import Member
from cse_b200.weights import synthetic_weights
from oracle import models as OM
from oracle import vote as OV

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

GROUPS = [("C3D", (16, 48, 48, 3), 2, 3), ("I3D", (16, 64, 64, 3), 2, 2), ("R3D_18", (16, 48, 48, 3), 3, 5)]
BATCH = 5


def _soft_head(g, w, mt, x0):
    """Random-weight members give logits in the thousands (one-hot softmax); rescale the head so the members'
    probabilities are soft and the vote is decided by their sum."""
    lg = OM.forward(mt, w, x0, torch.float32)[0]
    s = np.float32(3.0 / float(lg.abs().max()))
    w[g.output] = [w[g.output][0] * s, w[g.output][1] * s]
    return w


def _build():
    rng = np.random.default_rng(11)
    groups, graphs, weights = [], [], []
    for gi, (mt, shape, members, mb) in enumerate(GROUPS):
        g = G.build_model_graph(mt, shape, 11)
        x0 = rng.integers(0, 256, (1,) + shape, dtype=np.uint8)
        ws = [_soft_head(g, synthetic_weights(g, seed=500 + 10 * gi + j, nontrivial=True), mt, x0) for j in range(members)]
        groups.append((g, ws, mb))
        graphs.append(g)
        weights.append(ws)
    return groups, graphs, weights


def _batches(graphs, n_batches, n=BATCH, seed=21):
    rng = np.random.default_rng(seed)
    # smooth-ish content so members of different architectures disagree on some clips
    return [[[torch.from_numpy((rng.integers(0, 256, (n,) + tuple(g.shape(name)), dtype=np.uint8) // (1 + b % 3)
                                ).astype(np.uint8)).pin_memory() for name in g.inputs] for g in graphs]
            for b in range(n_batches)]


def test_hetero_resident_stream_and_per_member_paths_agree():
    groups, graphs, weights = _build()
    ens = HeteroEnsemble(groups, precision="bf16", max_batch=BATCH)
    assert ens.M == 7 and ens.micro_batch == [3, 2, 5]
    host = _batches(graphs, 4)
    # (1) resident path, batch by batch
    resident, probs_dev = [], []
    for b in host:
        dev = [[h.cuda() for h in hs] for hs in b]
        resident.append(ens.predict_device(dev).cpu().numpy().copy())
        probs_dev.append(ens.probs[:, :BATCH].cpu().numpy().copy())
    # (2) pipelined host path (3-deep), 4 different batches; then again on the same object (persisting buffers/events),
    #     and once more after abandoning a generator early
    streamed = [p.cpu().numpy().copy() for p in ens.stream_host(iter(host))]
    assert len(streamed) == len(host)
    for a, b in zip(resident, streamed):
        assert np.array_equal(a, b)
    gen = ens.stream_host(iter(host))
    first = next(gen).cpu().numpy().copy()
    del gen                                    # abandoned after one batch: uploads of batches 2..3 are in flight
    assert np.array_equal(first, resident[0])
    again = [p.cpu().numpy().copy() for p in ens.stream_host(iter(host[::-1]), depth=2)]
    for a, b in zip(resident[::-1], again):
        assert np.array_equal(a, b)
    # (3) every member on its own (whole plan, own workspace) + the oracle's fp64 vote on the GPU's probabilities
    for bi, b in enumerate(host):
        rows = []
        for (g, ws, mb), hs in zip(groups, b):
            xs = [h.numpy() for h in hs]
            for w in ws:
                m = Member(g, w, precision="bf16", max_batch=mb)
                rows.append(m.predict(xs if len(xs) > 1 else xs[0]))
                del m
        stack = np.stack(rows)
        assert np.array_equal(stack, probs_dev[bi]), "ensemble probabilities differ from the per-member path"
        exp = OV.ensemble_predictions(stack.astype(np.float64), np.ones(len(rows)))
        assert np.array_equal(resident[bi], exp.astype(np.int32))
    # the vote is not trivial: members disagree somewhere
    single = np.stack([p.argmax(-1) for p in probs_dev])            # [batch, M, n]
    assert (single != single[:, :1]).any()


def test_hetero_vote_matches_oracle_members():
    """Ensemble probabilities against the oracle's own fp32 forward pass of every member (bf16 tolerance)."""
    groups, graphs, weights = _build()
    ens = HeteroEnsemble(groups, precision="bf16", max_batch=BATCH)
    host = _batches(graphs, 1, seed=5)[0]
    pred = ens.predict_host(host)
    got = ens.probs[:, :BATCH].cpu().numpy()
    m = 0
    for (mt, shape, members, mb), ws, hs in zip(GROUPS, weights, host):
        xs = [h.numpy() for h in hs]
        for w in ws:
            exp = OM.forward(mt, w, xs if len(xs) > 1 else xs[0], torch.float32)[1].numpy()
            np.testing.assert_allclose(got[m], exp, rtol=0, atol=2e-2)
            m += 1
    assert pred.shape == (BATCH,) and pred.dtype == np.int32


# --------------------------------------------------------------------------- NCCL (needs >= 2 GPUs)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from cse_b200 import ensemble as E
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        groups, graphs, weights = _build()
        costs = [g.total_flops() for g, ws, _ in groups for _ in ws]
        units = E.shard_units(costs, BATCH, world)
        ens = HeteroEnsemble(groups, precision="bf16", max_batch=BATCH)
        host = _batches(graphs, 2)
        dev = [[[h.cuda() for h in hs] for hs in b] for b in host]
        full, probs_full = [], []
        for d in dev:                                   # every member, every clip, on this GPU alone
            full.append(ens.predict_device(d).cpu().numpy().copy())
            probs_full.append(ens.probs[:, :BATCH].clone())
        # the same steps with only this rank's (member, clip-chunk) units, merged over NCCL
        ens.set_units(units[rank], E.UnitGather(units, len(costs), BATCH, 11, dist, world, torch.device("cuda", rank)))
        for bi, d in enumerate(dev):
            pred = ens.predict_device(d).cpu().numpy()
            assert torch.equal(ens.gather.full, probs_full[bi]), "rank %d: gathered probabilities differ" % rank
            assert np.array_equal(pred, full[bi])
            np.save(os.path.join(out_dir, "nccl_rank%d_b%d.npy" % (rank, bi)), pred)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_unit_gather_on_nccl(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for bi in range(2):
        a = np.load(str(tmp_path / ("nccl_rank0_b%d.npy" % bi)))
        b = np.load(str(tmp_path / ("nccl_rank1_b%d.npy" % bi)))
        assert np.array_equal(a, b)
