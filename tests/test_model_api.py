"""cse_model_* (include/cse.h): graph construction and lowering inside the library.

CPU part: for every architecture family the native lowering (csrc/model_graph.h, model_lower.h, model.cu) must produce the
SAME plan as the Python lowering (cse_b200/graph.py + lowering.py) - every field of every `cse_op`, the packed weight arena
byte for byte, the workspace size - from the same Keras-layout weights handed over in `model.layers` order (the positional
rule of model.load_weights, train.py:1731-1769).  The GPU part runs a native model end to end."""
import ctypes as C

import numpy as np
import pytest

from cse_b200 import graph as G, lowering as L, runtime as rt
from cse_b200.weights import synthetic_weights

CASES = [("C3D", (16, 112, 112, 3), "bf16", 8), ("C3D", (16, 112, 112, 3), "bf16", 128), ("C3D", (16, 48, 48, 3), "fp32", 4),
         ("I3D", (20, 224, 224, 3), "bf16", 16), ("I3D", (64, 224, 224, 3), "bf16", 4), ("I3D", (20, 96, 96, 3), "fp32", 2),
         ("TWOSTREAM_I3D", (20, 224, 224, 0), "bf16", 8), ("R3D_18", (16, 64, 64, 3), "bf16", 3),
         ("R3D_34", (16, 112, 112, 3), "bf16", 256), ("R3D_34", (16, 112, 112, 3), "bf16", 32),
         ("R3D_50", (16, 48, 48, 3), "bf16", 2), ("R3D_50", (16, 48, 48, 3), "fp32", 2)]


def _fields(s):
    out = {}
    for name, typ in rt.CseOp._fields_:
        v = getattr(s, name)
        out[name] = tuple(v) if hasattr(v, "__len__") else v
    return out


@pytest.mark.parametrize("mt,shape,precision,nb", CASES)
def test_native_lowering_equals_python_lowering(mt, shape, precision, nb):
    g = G.build_model_graph(mt, shape, 11)
    w = synthetic_weights(g, seed=11, nontrivial=True)
    m = rt.NativeModel(mt, shape, 11, precision, nb)
    # the library lists the weighted layers in Keras model.layers order, like the graph the HDF5 reader pairs with
    assert [l for l, _ in m.layers()] == [n.name for n in g.weighted_layers()]
    for (lname, tensors), node in zip(m.layers(), g.weighted_layers()):
        assert [(t, tuple(s)) for t, s in tensors] == [(t, tuple(s)) for t, s in node.weights]
    m.set_weights(w)
    m.lower()
    plan = L.lower(g, w, precision, nb)
    ref_ops = plan.to_structs()
    ops = m.ops()
    assert len(ops) == len(ref_ops)
    for i, (a, b) in enumerate(zip(ops, ref_ops)):
        fa, fb = _fields(a), _fields(b)
        diff = {k: (fa[k], fb[k]) for k in fa if fa[k] != fb[k]}
        assert not diff, "op %d (%s): native vs python %r" % (i, plan.ops[i].name, diff)
    assert m.workspace_bytes() == plan.workspace_bytes
    arena = m.weight_arena()
    assert arena.shape == plan.weight_arena.shape
    assert np.array_equal(arena, plan.weight_arena), "first differing byte at %d" % int(np.argmax(arena != plan.weight_arena))


def test_native_options_and_errors():
    g = G.build_model_graph("TWOSTREAM_I3D", (20, 96, 96, 0), 11)
    w = synthetic_weights(g, seed=3)
    m = rt.NativeModel("TWOSTREAM_I3D", (20, 96, 96, 0), 11, "bf16", 2, persist_input=True, flow_input_f32=True)
    m.set_weights(w)
    m.lower()
    plan = L.lower(g, w, "bf16", 2, persist_input=True, input_dtypes=("u8", "f32"))
    for a, b in zip(m.ops(), plan.to_structs()):
        assert _fields(a) == _fields(b)
    assert m.workspace_bytes() == plan.workspace_bytes
    with pytest.raises(rt.CseError):
        rt.NativeModel("VGG", (16, 112, 112, 3))
    m2 = rt.NativeModel("C3D", (16, 48, 48, 3), 11, "bf16", 2)
    with pytest.raises(rt.CseError):
        m2.lower()                                   # weights were never set
    bad = np.zeros((3, 3, 3, 3, 32), np.float32)
    dims = (C.c_int64 * 5)(*bad.shape)
    assert m2.lib.cse_model_set_weight(m2.handle, 0, 0, bad.ctypes.data, dims, 5) != 0      # conv1 has 64 filters


@pytest.mark.parametrize("mt,shape,nb,f32flow", [("I3D", (64, 224, 224, 3), 4, False), ("R3D_34", (16, 112, 112, 3), 32, False),
                                                 ("TWOSTREAM_I3D", (20, 224, 224, 0), 8, False), ("TWOSTREAM_I3D", (20, 96, 96, 0), 2, True)])
def test_native_stem_pair_equals_python_roles(mt, shape, nb, f32flow):
    """cse_model_pair_stems: the leader's and the follower's native plans equal lowering.lower(stem_role='lead' / 'follow')
    field by field and byte by byte (the N = 128 stem op with the peer's columns, the persistent peer buffer's offset)."""
    g = G.build_model_graph(mt, shape, 11)
    w0, w1 = synthetic_weights(g, seed=5, nontrivial=True), synthetic_weights(g, seed=6, nontrivial=True)
    kw = dict(flow_input_f32=True) if f32flow else {}
    lead, fol = rt.NativeModel(mt, shape, 11, "bf16", nb, **kw), rt.NativeModel(mt, shape, 11, "bf16", nb, **kw)
    lead.set_weights(w0)
    fol.set_weights(w1)
    lead.pair_stems(fol)
    lead.lower()
    fol.lower()
    pkw = dict(persist_input=True, **(dict(input_dtypes=("u8", "f32")) if f32flow else {}))
    for m, plan in ((lead, L.lower(g, w0, "bf16", nb, stem_role="lead", stem_peer=w1, **pkw)),
                    (fol, L.lower(g, w1, "bf16", nb, stem_role="follow", **pkw))):
        ref_ops, ops = plan.to_structs(), m.ops()
        assert len(ops) == len(ref_ops)
        for i, (a, b) in enumerate(zip(ops, ref_ops)):
            fa, fb = _fields(a), _fields(b)
            diff = {k: (fa[k], fb[k]) for k in fa if fa[k] != fb[k]}
            assert not diff, "op %d (%s): native vs python %r" % (i, plan.ops[i].name, diff)
        assert m.workspace_bytes() == plan.workspace_bytes
        assert np.array_equal(m.weight_arena(), plan.weight_arena)
    # pairing rules
    c1, c2 = rt.NativeModel("C3D", (16, 48, 48, 3), 11, "bf16", 2), rt.NativeModel("C3D", (16, 48, 48, 3), 11, "bf16", 2)
    cg = G.build_model_graph("C3D", (16, 48, 48, 3), 11)
    c1.set_weights(synthetic_weights(cg, seed=1))
    c2.set_weights(synthetic_weights(cg, seed=2))
    with pytest.raises(rt.CseError):
        c1.pair_stems(c2)                                # C3D has no 7x7x7 stem
    a, b = rt.NativeModel(mt, shape, 11, "bf16", nb), rt.NativeModel(mt, shape, 11, "bf16", nb)
    a.set_weights(w0)
    with pytest.raises(rt.CseError):
        a.pair_stems(b)                                  # the follower's weights are not set yet
    with pytest.raises(rt.CseError):
        lead.pair_stems(fol)                             # already lowered


@pytest.mark.gpu
@pytest.mark.parametrize("mt,shape,n", [("I3D", (20, 96, 96, 3), 2), ("R3D_18", (16, 64, 64, 3), 3)])
def test_native_stem_pair_forward(mt, shape, n):
    """A paired leader / follower on one shared workspace: bit-identical logits to the stand-alone Python members."""
    import torch
    from cse_b200.model import Member
    g = G.build_model_graph(mt, shape, 11)
    ws = [synthetic_weights(g, seed=31 + j, nontrivial=True) for j in range(2)]
    x = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (n,) + shape, dtype=np.uint8)).cuda()
    refs = [Member(g, w, precision="bf16", max_batch=n).forward_device([x])[0] for w in ws]
    lead, fol = rt.NativeModel(mt, shape, 11, "bf16", n), rt.NativeModel(mt, shape, 11, "bf16", n)
    lead.set_weights(ws[0])
    fol.set_weights(ws[1])
    lead.pair_stems(fol)
    lead.lower()
    fol.lower()
    shared = torch.empty(max(lead.workspace_bytes(), fol.workspace_bytes()) + 1024, dtype=torch.uint8, device="cuda")
    lead.finalize(shared)
    fol.finalize(shared)
    for _ in range(2):                                   # twice: the peer buffer is rewritten by every leader pass
        l0, _ = lead.forward(x)
        l1, _ = fol.forward(x, shared_input=True)
        torch.cuda.synchronize()
        assert torch.equal(l0, refs[0]) and torch.equal(l1, refs[1])


@pytest.mark.gpu
@pytest.mark.parametrize("mt,shape,n", [("C3D", (16, 112, 112, 3), 3), ("I3D", (20, 96, 96, 3), 2), ("R3D_18", (16, 64, 64, 3), 3)])
def test_native_model_forward_matches_member(mt, shape, n):
    """The native model end to end on the GPU: bit-identical logits / probabilities to the Python-lowered Member (same
    plan), and a second member of the fold re-using the pre-processed clips through a shared workspace."""
    import torch
    from cse_b200.model import Member
    g = G.build_model_graph(mt, shape, 11)
    w = synthetic_weights(g, seed=21, nontrivial=True)
    x = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (n,) + shape, dtype=np.uint8)).cuda()
    ref_logits, ref_probs = Member(g, w, precision="bf16", max_batch=n).forward_device([x])
    m = rt.NativeModel(mt, shape, 11, "bf16", n, persist_input=True)
    m.set_weights(w)
    m.lower()
    shared = torch.empty(m.workspace_bytes() + 1024, dtype=torch.uint8, device="cuda")
    m.finalize(shared)
    logits, probs = m.forward(x)
    torch.cuda.synchronize()
    assert torch.equal(logits, ref_logits) and torch.equal(probs, ref_probs)
    w2 = synthetic_weights(g, seed=22, nontrivial=True)
    m2 = rt.NativeModel(mt, shape, 11, "bf16", n, persist_input=True)
    m2.set_weights(w2)
    m2.finalize(shared)
    l2, _ = m2.forward(x, shared_input=True)          # clips pre-processed by m stay valid in the shared workspace
    ref2, _ = Member(g, w2, precision="bf16", max_batch=n).forward_device([x])
    torch.cuda.synchronize()
    assert torch.equal(l2, ref2)


@pytest.mark.gpu
def test_plain_c_host_builds_and_runs(tmp_path):
    """tools/c_abi_demo.c: a plain C program (no Python, no torch, no CUDA headers) builds a 2-member C3D ensemble through
    cse_model_* / cse_vote and checks the vote on the host."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "crowded-scenes-ensemble-classification_b200")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.run(["gcc", "-O2", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "tools", "c_abi_demo.c"),
                    "-L", libdir, "-lcse_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "c_abi_demo ok" in out.stdout, out.stdout + out.stderr
