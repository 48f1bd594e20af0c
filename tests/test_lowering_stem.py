"""CPU check of the stride-2 stem lowering (no GPU): the space-to-depth clip layout written by the pre-processing
kernel + the regrouped, zero-extended kernel of lowering._s2d_stem_conv are the same sum of products as the
reference's Conv3D(64, 7x7x7, strides 2, 'same') (train.py:1026, 1481), for 2x2 and 2x2x2 cells, even and odd
extents."""
import numpy as np
import pytest
import torch

from cse_b200 import graph as G, lowering as L
from cse_b200.weights import synthetic_weights
from oracle import ops as O


def s2d_cells(x, s2d, cell, wpitch, wpad):
    """numpy restatement of the MODE 20 / 21 layouts of csrc/ops.cu::preprocess_rows_kernel."""
    n, t, h, w, c = x.shape
    t2 = (t + 1) // 2 if s2d == 2 else t
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    out = np.zeros((n, t2, h2, wpitch, cell), np.float64)
    for pd in range(2 if s2d == 2 else 1):
        for ph in range(2):
            for pw in range(2):
                src = x[:, pd::2] if s2d == 2 else x
                src = src[:, :, ph::2, pw::2]
                k0 = ((pd * 2 + ph) * 2 + pw) * c
                out[:, :src.shape[1], :src.shape[2], wpad:wpad + src.shape[3], k0:k0 + c] = src
    return out


@pytest.mark.parametrize("depth", [True, False, "always"])
@pytest.mark.parametrize("dhw,c", [((8, 16, 16), 3), ((9, 21, 19), 3), ((7, 10, 34), 1), ((6, 20, 28), 2), ((5, 9, 12), 3)])
def test_s2d_stem_regrouping_equals_reference_conv(dhw, c, depth):
    g = G.Graph("t", "functional")
    x = g.input(dhw + (c,), name="in")
    g.conv3d(x, 16, (7, 7, 7), (2, 2, 2), "same", True, None, name="c")
    w = synthetic_weights(g, seed=3, nontrivial=True)
    low = L.Lowerer(g, w, "bf16", 2, s2d_depth=depth)
    captured = {}
    orig = low._conv_like

    def spy(name, view, k2, bias, k, s, pads, out_dims, *a, **kw):
        captured.update(k2=k2, k=tuple(k), s=tuple(s), pads=tuple(pads), view=view, out_dims=tuple(out_dims))
        return orig(name, view, k2, bias, k, s, pads, out_dims, *a, **kw)
    low._conv_like = spy
    low.lower()
    view, k2 = captured["view"], captured["k2"]
    pre = low.ops[0].out0
    assert pre.s2d == (2 if (depth == "always" or (depth and c in (1, 2))) else 1)
    clips = np.random.default_rng(0).integers(0, 256, (2,) + dhw + (c,)).astype(np.float64)
    cells = s2d_cells(clips, pre.s2d, pre.ld, pre.wpitch, pre.wpad)
    # the overlapping-stride TMA view: position w2 sees the 4 cells w2 .. w2+3 of the padded row
    w2 = pre.dims[2]
    win = np.concatenate([cells[:, :, :, j:j + w2, :] for j in range(4)], axis=-1)
    assert win.shape[-1] == view.C
    kd, kh = captured["k"][:2]
    sd = captured["s"][0]
    pd_, ph_ = captured["pads"][:2]
    od, oh, ow = captured["out_dims"]
    xp = np.zeros((2, (od - 1) * sd + kd, oh - 1 + kh, w2, view.C))
    d_hi = min(win.shape[1], xp.shape[1] - pd_)
    h_hi = min(win.shape[2], xp.shape[2] - ph_)
    xp[:, pd_:pd_ + d_hi, ph_:ph_ + h_hi] = win[:, :d_hi, :h_hi]
    got = np.zeros((2, od, oh, ow, 16))
    for fd in range(kd):
        for fh in range(kh):
            got += np.einsum("ndhwc,co->ndhwo", xp[:, fd:fd + (od - 1) * sd + 1:sd, fh:fh + oh, :ow], k2[fd, fh, 0].astype(np.float64))
    kern, bias = w["c"]
    exp = O.conv3d(torch.as_tensor(clips), torch.as_tensor(kern, dtype=torch.float64), None, (2, 2, 2), "same").numpy()
    assert got.shape == exp.shape
    np.testing.assert_allclose(got, exp, rtol=1e-9, atol=1e-6)
    # executed K of the regrouped conv
    assert kd * kh * view.C == (16 * 32 * c if pre.s2d == 2 else 28 * 4 * pre.ld)


@pytest.mark.parametrize("mt,shape", [("I3D", (16, 64, 64, 3)), ("R3D_18", (16, 56, 56, 3)), ("TWOSTREAM_I3D", (16, 64, 64, 0))])
def test_stem_roles_share_buffers(mt, shape):
    """Stem fusion across two members (DeviceEnsemble(fuse_stems=True)): the leader's, the follower's and the stand-alone
    lowering of one ensemble must place the pre-processed clips at the same workspace offsets, leader and follower must
    agree on the peer buffer, the leader's stem is one N = 128 op with a column split, the follower has no stem op and
    its first consumers read the peer buffer."""
    from cse_b200 import runtime as rt
    from cse_b200.lowering import Lowerer, lower
    g = G.build_model_graph(mt, shape, 11)
    assert Lowerer.stem_fusable(g) and not Lowerer.stem_fusable(g, "fp32")
    w0, w1 = synthetic_weights(g, seed=1), synthetic_weights(g, seed=2)
    solo = lower(g, w0, "bf16", 4, persist_input=True)
    lead = lower(g, w0, "bf16", 4, persist_input=True, stem_role="lead", stem_peer=w1)
    fol = lower(g, w1, "bf16", 4, persist_input=True, stem_role="follow")

    def pre(p):
        return [(o.name, o.out0.buf.offset, o.out0.buf.nbytes) for o in p.ops if o.kind == rt.OP_PREPROCESS]

    def peer(p):
        return sorted((b.name, b.offset, b.nbytes) for b in p.buffers if b.name.endswith(":peer"))

    assert pre(solo) == pre(lead) == pre(fol)
    assert peer(lead) == peer(fol) and len(peer(lead)) == len(g.inputs) and not peer(solo)
    top_pre = max(o + n for _, o, n in pre(lead))
    assert min(o for _, o, _ in peer(lead)) == top_pre                 # right behind the clips, before any plan-local buffer
    stems = [o for o in lead.ops if o.name.endswith("+peer")]
    assert len(stems) == len(g.inputs)
    for o in stems:
        assert (o.bn, o.out_split, o.halo, o.kc) == (128, 64, 3, 64) and o.out1.buf.name.endswith(":peer")
        ref = [s for s in solo.ops if s.name == o.name[:-5]][0]
        assert o.flops == 2 * ref.flops and o.k == ref.k and o.s == ref.s and o.pad == ref.pad and o.brick == ref.brick
    assert len(fol.ops) == len(solo.ops) - len(g.inputs) and len(lead.ops) == len(solo.ops)
    readers = [o for o in fol.ops if o.in0 is not None and o.in0.buf.name.endswith(":peer")]
    assert len(readers) >= len(g.inputs)
    assert lead.workspace_bytes >= max(fol.workspace_bytes, solo.workspace_bytes)
    with pytest.raises(ValueError):
        lower(g, w0, "bf16", 4, stem_role="lead")                      # a leader needs its peer's weights
    cg = G.build_model_graph("C3D", (16, 56, 56, 3), 11)
    assert not Lowerer.stem_fusable(cg)
