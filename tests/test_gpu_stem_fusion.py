"""Stem fusion across the members of one ensemble (DeviceEnsemble(fuse_stems=True), the default): members 2k / 2k+1
share their pre-processed clips, so their 64-filter 7x7x7 stems (I3D Conv3d_1a_7x7 train.py:1026, the TwoStream flow
tower's :999-1009, the R3D stem :1481) run as one tcgen05 GEMM with N = 128 through the CTA-pair kernel.  Every member's
logits must be bit-identical to the unfused ensemble's, for whole-ensemble steps and for the member subsets of
unit-sharded steps (a member whose partner is absent runs its stand-alone plan)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from cse_b200 import graph as G, runtime as rt                  # noqa: E402
from cse_b200.ensemble_runtime import DeviceEnsemble            # noqa: E402
from cse_b200.lowering import Lowerer                           # noqa: E402
from cse_b200.weights import synthetic_weights                  # noqa: E402

CASES = [("I3D", (16, 64, 64, 3)), ("R3D_18", (16, 56, 56, 3)), ("TWOSTREAM_I3D", (16, 64, 64, 0))]


def clips_for(g, n, seed):
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.integers(0, 256, (n,) + tuple(g.shape(name)), dtype=np.uint8)).cuda() for name in g.inputs]


@pytest.mark.parametrize("mt,shape", CASES)
def test_fused_stems_bit_identical(mt, shape):
    g = G.build_model_graph(mt, shape, 11)
    assert Lowerer.stem_fusable(g)
    ws = [synthetic_weights(g, seed=10 + j) for j in range(3)]          # a pair and an unpaired third member
    x = clips_for(g, 6, 3)
    plain = DeviceEnsemble(g, ws, max_batch=8, micro_batch=4, fuse_stems=False)
    fused = DeviceEnsemble(g, ws, max_batch=8, micro_batch=4)
    assert not plain.fuse_stems and plain.roles == [None, None, None]
    assert fused.fuse_stems and fused.roles == ["lead", "follow", None]
    lead_ops = [o for o in fused.members[0].plan.ops if o.name.endswith("+peer")]
    assert len(lead_ops) == len(g.inputs) and all(o.bn == 128 and o.out_split == 64 and o.halo == 3 for o in lead_ops)
    assert len(fused.members[1].plan.ops) == len(plain.members[1].plan.ops) - len(g.inputs)
    plain.forward_members(x)
    fused.forward_members(x)
    torch.cuda.synchronize()
    assert torch.equal(plain.logits[:, :6], fused.logits[:, :6])
    assert torch.equal(plain.probs[:, :6], fused.probs[:, :6])
    assert float(plain.logits[:, :6].abs().max()) > 0
    assert torch.equal(plain.predict_device(x), fused.predict_device(x))
    # member subsets (unit-sharded steps): a leader without its follower, a follower without its leader, any order
    for ids in ([1], [0], [1, 2], [0, 2], [0, 1, 2], [2, 1], [2, 0, 1]):
        fused.logits.zero_()
        fused.forward_subset(x, ids, 0)
        torch.cuda.synchronize()
        for j in range(3):
            want = plain.logits[j, :6] if j in ids else torch.zeros_like(plain.logits[j, :6])
            assert torch.equal(fused.logits[j, :6], want), (ids, j)
    # and the whole step again after the stand-alone members were built on the shared workspace
    fused.forward_members(x)
    torch.cuda.synchronize()
    assert torch.equal(plain.logits[:, :6], fused.logits[:, :6])


def test_fused_stems_float_flow_and_profile():
    """TwoStream with the FarneBack float flow input: the fused flow stem reads the float-sourced layout; profile_ops
    attributes the fused stem's FLOPs to both members."""
    g = G.build_model_graph("TWOSTREAM_I3D", (16, 64, 64, 0), 11)
    ws = [synthetic_weights(g, seed=20 + j) for j in range(2)]
    rng = np.random.default_rng(4)
    x = [torch.from_numpy(rng.integers(0, 256, (4, 16, 64, 64, 3), dtype=np.uint8)).cuda(),
         torch.from_numpy(rng.standard_normal((4, 16, 64, 64, 2)).astype(np.float32) * 3).cuda()]
    plain = DeviceEnsemble(g, ws, max_batch=4, micro_batch=4, fuse_stems=False, input_dtypes=("u8", "f32"))
    fused = DeviceEnsemble(g, ws, max_batch=4, micro_batch=4, input_dtypes=("u8", "f32"))
    assert fused.roles == ["lead", "follow"]
    plain.forward_members(x)
    fused.forward_members(x)
    torch.cuda.synchronize()
    assert torch.equal(plain.logits, fused.logits)
    pa = {r["name"]: r for r in plain.profile_ops(x, iters=1)}
    pf = {r["name"]: r for r in fused.profile_ops(x, iters=1)}
    assert set(pa) == set(pf)
    for k in pa:
        assert pa[k]["flops"] == pytest.approx(pf[k]["flops"]) and pa[k]["bytes"] == pytest.approx(pf[k]["bytes"]), k


@pytest.mark.parametrize("mt,shape", [("I3D", (64, 224, 224, 3)), ("TWOSTREAM_I3D", (20, 224, 224, 0)), ("R3D_34", (16, 112, 112, 3))])
def test_fused_stems_bit_identical_baseline_shapes(mt, shape):
    """BASELINE.json's own clip geometries (configs[2], the reference-true T = 20 TwoStream, the R3D-34 of configs[4])."""
    g = G.build_model_graph(mt, shape, 11)
    ws = [synthetic_weights(g, seed=30 + j) for j in range(2)]
    x = clips_for(g, 3, 5)
    plain = DeviceEnsemble(g, ws, max_batch=4, micro_batch=2, fuse_stems=False)
    fused = DeviceEnsemble(g, ws, max_batch=4, micro_batch=2)
    assert fused.roles == ["lead", "follow"]
    plain.forward_members(x)
    fused.forward_members(x)
    torch.cuda.synchronize()
    assert torch.equal(plain.logits[:, :3], fused.logits[:, :3]) and torch.equal(plain.probs[:, :3], fused.probs[:, :3])
    assert torch.isfinite(fused.logits[:, :3]).all() and float(fused.logits[:, :3].abs().max()) > 0
