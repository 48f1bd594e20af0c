"""Per-kernel parity on the B200: every CUDA kernel (through the C ABI) vs the CPU oracle on the
same seeded inputs.  Shapes cover the (k, stride, pad, C, N) tuples of SURVEY App. B.

Tolerances (written here, per the task contract):
  * fp32 path  : |gpu - oracle_fp64| <= 1e-4 * max|oracle|   (north-star: <= 1e-4 relative)
  * bf16 path  : op input is read back from the GPU (already bf16) and the oracle is run in fp64
                 on those exact values with bf16-rounded weights, so the only differences are the
                 fp32 accumulation order and the final bf16 rounding: <= 2^-7 relative per element
                 (+ tiny absolute slack).
  * integer / byte kernels (preprocess, vote) : bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import ops as O
from oracle import vote as OV
from cse_b200 import graph as G
from cse_b200 import runtime as rt
from cse_b200.model import Member
from cse_b200.weights import synthetic_weights

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

T64 = torch.float64

_tuned = set()


def tune(key, value):
    """cse_tune with automatic reset to the built-in default after the test (see _reset_tuning)."""
    _tuned.add(key)
    rt.tune(key, value)


@pytest.fixture(autouse=True)
def _reset_tuning():
    yield
    while _tuned:
        rt.tune(_tuned.pop(), -1)


def bf16_round(a):
    return torch.as_tensor(np.asarray(a, np.float32)).to(torch.bfloat16).to(torch.float64)


def make_member(build, precision, nb, seed=0, **kw):
    g = G.Graph("test", "functional")
    build(g)
    w = synthetic_weights(g, seed=seed, nontrivial=True)
    m = Member(g, w, precision=precision, max_batch=nb, keep_all=True, **kw)
    return g, w, m


def run(m, x_u8_list):
    dev = [torch.from_numpy(x).cuda() for x in x_u8_list]
    m.run_ops(dev, 0, m.num_ops)
    torch.cuda.synchronize()


def clips(seed, n, shape):
    return np.random.default_rng(seed).integers(0, 256, (n,) + tuple(shape), dtype=np.uint8)


# --------------------------------------------------------------------------- library
def test_library_and_device():
    lib = rt.load_library()
    assert lib.cse_abi_version() == rt.ABI_VERSION == 2
    sm, mj, mn = rt.device_info()
    assert sm > 0 and mj == 10, "expected a Blackwell (sm_100) device, got cc %d.%d" % (mj, mn)


# --------------------------------------------------------------------------- preprocess
@pytest.mark.parametrize("c,dtype,ld", [(3, "bf16", 8), (2, "bf16", 8), (3, "fp32", 3), (2, "fp32", 2), (3, "fp32", 4)])
def test_preprocess_exact(c, dtype, ld):
    x = clips(1, 3, (5, 12, 10, c))
    out = rt.preprocess(torch.from_numpy(x).cuda(), dtype, ld).float().cpu().numpy()
    exp = np.zeros(x.shape[:-1] + (ld,), np.float32)
    exp[..., :c] = x.astype(np.float32)          # train.py:466-478: raw frames stored as float32
    assert np.array_equal(out, exp)              # 0..255 are exact in bf16 and fp32


def test_preprocess_crop_mean_scale():
    x = clips(2, 2, (6, 16, 14, 3))
    mean, scale = [10.0, 20.0, 30.0], [0.5, 0.25, 2.0]
    out = rt.preprocess(torch.from_numpy(x).cuda(), "fp32", 3, crop=(1, 2, 3, 4, 10, 8), mean=mean,
                        scale=scale).cpu().numpy()
    exp = (x[:, 1:5, 2:12, 3:11, :].astype(np.float32) - np.float32(mean)) * np.float32(scale)
    assert np.array_equal(out, exp)


# --------------------------------------------------------------------------- conv, fp32 direct engine
CONV_CASES = [  # (in dhw, cin_mid, cout, k, s, padding)
    ((6, 12, 12), 0, 24, (3, 3, 3), (1, 1, 1), "same"),
    ((9, 20, 20), 0, 16, (7, 7, 7), (2, 2, 2), "same"),
    ((6, 12, 12), 16, 32, (3, 3, 3), (2, 2, 2), "same"),
    ((5, 7, 7), 16, 40, (1, 1, 1), (1, 2, 2), "valid"),
    ((4, 8, 8), 24, 16, (1, 1, 1), (2, 2, 2), "valid"),
    ((4, 9, 9), 16, 11, (1, 1, 1), (1, 1, 1), "same"),
]


@pytest.mark.parametrize("dhw,cmid,cout,k,s,pad", CONV_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_direct_vs_oracle(dhw, cmid, cout, k, s, pad, precision):
    def build(g):
        x = g.input(dhw + (3,), name="in")
        if cmid:
            x = g.conv3d(x, cmid, (1, 1, 1), (1, 1, 1), "same", True, "relu", name="pre")
        x = g.conv3d(x, cout, k, s, pad, True, None, name="c")
        x = g.bn(x, scale=True, name="b")
        g.relu(x, name="r")
    g, w, m = make_member(build, precision, 2, tc=False, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    xs = clips(3, 2, dhw + (3,))
    run(m, [xs])
    src = "pre" if cmid else "in"
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors[src], 2), dtype=T64)
    kern, bias = w["c"]
    if precision == "bf16":
        kern = bf16_round(kern).numpy()
    y = O.conv3d(xin, torch.as_tensor(kern, dtype=T64), torch.as_tensor(bias, dtype=T64), s, pad)
    y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]]))
    got = m.read_tensor(m.plan.tensors["r"], 2)
    exp = y.numpy()
    tol = 1e-4 if precision == "fp32" else 2.0 ** -7
    err = np.abs(got - exp).max() / max(np.abs(exp).max(), 1e-6)
    assert got.shape == exp.shape
    assert err <= tol, "rel err %g" % err


# --------------------------------------------------------------------------- conv, tcgen05 engine
TC_CASES = [  # (in dhw, cin, cout, k, nb)
    ((4, 8, 8), 64, 128, (3, 3, 3), 2),       # C3D conv2 family, one K chunk of 64
    ((4, 8, 8), 128, 256, (3, 3, 3), 2),      # two chunks, N=256
    ((2, 7, 7), 64, 512, (3, 3, 3), 3),       # two N tiles, ragged 7x7 brick
    ((4, 14, 14), 32, 96, (3, 3, 3), 2),      # kc=32 (SWIZZLE_64B)
    ((3, 6, 6), 16, 48, (3, 3, 3), 2),        # kc=16 (SWIZZLE_32B)
    ((3, 6, 6), 24, 208, (3, 3, 3), 2),       # Cin=24 -> OOB-filled channel tail; N=208
    ((5, 9, 9), 192, 64, (1, 1, 1), 2),       # 1x1x1 (Inception branches)
    ((2, 5, 5), 480, 24, (1, 1, 1), 2),       # Cin=480 (kc=32), Cout=24 -> N tile 32 overhang
    ((1, 1, 1), 512, 4096, (1, 1, 1), 5),     # Dense as 1x1x1 conv on [n,1,1,1,K]
]


@pytest.mark.parametrize("dhw,cin,cout,k,nb", TC_CASES)
def test_conv_tcgen05_vs_oracle(dhw, cin, cout, k, nb):
    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, cin, (1, 1, 1), (1, 1, 1), "same", True, "relu", name="pre")
        x = g.conv3d(x, cout, k, (1, 1, 1), "same", True, None, name="c")
        x = g.bn(x, scale=True, name="b")
        g.relu(x, name="r")
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    op = [o for o in m.plan.ops if o.name == "c"][0]
    assert op.engine == rt.ENGINE_TCGEN05
    xs = clips(4, nb, dhw + (3,))
    run(m, [xs])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], nb), dtype=T64)
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same")
    y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]]))
    got = m.read_tensor(m.plan.tensors["r"], nb)
    exp = y.numpy()
    err = np.abs(got - exp).max() / max(np.abs(exp).max(), 1e-6)
    assert err <= 2.0 ** -7, "rel err %g (kc=%d bn=%d brick=%s)" % (err, op.kc, op.bn, op.brick)


SIBLING_CASES = [((4, 14, 14), 64, (32, 48, 16, 24, 8, 16), 3),      # split 32 -> 64-wide chunks shrink to 32
                 ((3, 9, 11), 40, (112, 48, 32, 8, 16, 8), 2),      # split 112 -> 16-wide chunks, ragged bricks
                 ((2, 7, 7), 256, (256, 160, 64, 32, 32, 16), 5)]   # two N tiles, split on a tile boundary or not


@pytest.mark.parametrize("dhw,cin,f,nb", SIBLING_CASES)
def test_conv_tcgen05_fused_siblings(dhw, cin, f, nb):
    """Inception block (train.py:1048-1064): the three 1x1x1 convs reading the block input run as ONE
    tcgen05 GEMM whose output columns are stored to three tensors (branch 0 into its concat slice, 1a and
    2a into dense buffers of their own); every branch and the concat must match the oracle."""
    from cse_b200.graph import _conv3d_bn
    k1, k3, s1 = (1, 1, 1), (3, 3, 3), (1, 1, 1)

    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, cin, k1, s1, "same", True, "relu", name="pre")
        b0 = _conv3d_bn(g, x, f[0], k1, s1, "b0")
        b1 = _conv3d_bn(g, x, f[1], k1, s1, "b1a")
        b1 = _conv3d_bn(g, b1, f[2], k3, s1, "b1b")
        b2 = _conv3d_bn(g, x, f[3], k1, s1, "b2a")
        b2 = _conv3d_bn(g, b2, f[4], k3, s1, "b2b")
        b3 = g.maxpool(x, k3, s1, "same", name="b3a")
        b3 = _conv3d_bn(g, b3, f[5], k1, s1, "b3b")
        g.concat([b0, b1, b2, b3], name="cat")
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    fused = [o for o in m.plan.ops if o.out_split > 0]
    assert len(fused) == 1 and fused[0].out_split == f[0] and fused[0].engine == rt.ENGINE_TCGEN05
    assert fused[0].out0.C == f[0] + f[1] + f[3]
    run(m, [clips(9, nb, dhw + (3,))])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], nb), dtype=T64)

    def cbr(x, name, k):
        y = O.conv3d(x, bf16_round(w[name + "_conv"][0]), None, s1, "same")
        return O.relu(O.batchnorm(y, None, *[torch.as_tensor(a, dtype=T64) for a in w[name + "_bn"]]))
    exp = {"b0": cbr(xin, "b0", k1), "b1a": cbr(xin, "b1a", k1), "b2a": cbr(xin, "b2a", k1)}
    for name, e in exp.items():
        got = m.read_tensor(m.plan.tensors[name], nb)
        e = e.numpy()
        err = np.abs(got - e).max() / max(np.abs(e).max(), 1e-6)
        assert err <= 2.0 ** -7, "%s rel err %g" % (name, err)
    # downstream: the 3x3x3 convs read the scratch slices, the concat holds all four branches
    a1 = torch.as_tensor(m.read_tensor(m.plan.tensors["b1a"], nb), dtype=T64)
    a2 = torch.as_tensor(m.read_tensor(m.plan.tensors["b2a"], nb), dtype=T64)
    p3 = O.maxpool3d(xin, k3, s1, "same")
    assert np.array_equal(m.read_tensor(m.plan.tensors["b3a"], nb), p3.numpy().astype(np.float32))
    cat = np.concatenate([m.read_tensor(m.plan.tensors["b0"], nb), cbr(a1, "b1b", k3).numpy(),
                          cbr(a2, "b2b", k3).numpy(), cbr(p3, "b3b", k1).numpy()], axis=-1)
    got = m.read_tensor(m.plan.tensors["cat"], nb)
    assert got.shape == cat.shape
    err = np.abs(got - cat).max() / np.abs(cat).max()
    assert err <= 2.0 ** -7, "concat rel err %g" % err
    # and the unfused lowering gives bit-identical activations (same MMA order per output column)
    g2, w2, m2 = make_member(build, "bf16", nb, scale=[1 / 64.0] * 3, mean=[128.0] * 3, fuse_siblings=False)
    assert not [o for o in m2.plan.ops if o.out_split > 0]
    run(m2, [clips(9, nb, dhw + (3,))])
    assert np.array_equal(m2.read_tensor(m2.plan.tensors["cat"], nb), got)


TWIN_CASES = [((4, 8, 8), 64, 128, (3, 3, 3), 2), ((5, 9, 9), 192, 64, (1, 1, 1), 2), ((3, 6, 6), 16, 48, (3, 3, 3), 3),
              ((4, 14, 14), 32, 96, (3, 3, 3), 1), ((2, 5, 5), 480, 24, (1, 1, 1), 3)]


@pytest.mark.parametrize("dhw,cin,cout,k,nb", TWIN_CASES)
def test_conv_tcgen05_twin_tiles(dhw, cin, cout, k, nb):
    """Twin-tile mode (two M tiles share every weight stage, four TMEM accumulators), forced on small
    shapes so that even and odd tile counts, several tiles per CTA and the residual / second-output
    epilogue all run through it."""
    tune("twin_min_tiles", 2)

    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, cin, (1, 1, 1), (1, 1, 1), "same", True, "relu", name="pre")
        x = g.conv3d(x, cout, k, (1, 1, 1), "same", True, None, name="c")
        x = g.bn(x, scale=True, name="b")
        g.relu(x, name="r")
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    op = [o for o in m.plan.ops if o.name == "c"][0]
    assert op.engine == rt.ENGINE_TCGEN05 and op.bn <= 128
    xs = clips(4, nb, dhw + (3,))
    run(m, [xs])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], nb), dtype=T64)
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same")
    y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]]))
    got = m.read_tensor(m.plan.tensors["r"], nb)
    err = np.abs(got - y.numpy()).max() / max(np.abs(y.numpy()).max(), 1e-6)
    assert err <= 2.0 ** -7, "rel err %g (kc=%d bn=%d brick=%s)" % (err, op.kc, op.bn, op.brick)


def test_twin_tiles_whole_r3d_matches_default():
    """R3D_18 (residual + second-output epilogues, strided convs) with the twin path forced everywhere
    it is legal gives bit-identical logits to the default tiling: same products, same per-tile
    accumulation order."""
    shape = (16, 64, 64, 3)
    g = G.build_model_graph("R3D_18", shape, 11)
    w = synthetic_weights(g, seed=5, nontrivial=True)
    x = torch.from_numpy(clips(30, 3, shape)).cuda()
    tune("twin_min_tiles", 0)
    base, _ = Member(g, w, precision="bf16", max_batch=3).forward_device([x])
    base = base.cpu().numpy()
    tune("twin_min_tiles", 2)
    twin, _ = Member(g, w, precision="bf16", max_batch=3).forward_device([x])
    assert np.array_equal(base, twin.cpu().numpy())


@pytest.mark.parametrize("unroll", [True, False])
@pytest.mark.parametrize("dhw,c,cout,nb", [((4, 16, 16), 3, 64, 2), ((3, 9, 13), 3, 24, 3), ((5, 8, 8), 2, 64, 2)])
def test_conv_tcgen05_packed_stem(dhw, c, cout, nb, unroll):
    """First-layer 3x3x3 conv on the C<=8 uint8 clip (C3D conv1): kw taps folded into a 32-wide K
    chunk read through an overlapping-stride TMA view of the W-padded pre-processed clip."""
    def build(g):
        x = g.input(dhw + (c,), name="in")
        g.conv3d(x, cout, (3, 3, 3), (1, 1, 1), "same", True, "relu", name="c")
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * c, mean=[128.0] * c, stem_unroll=unroll)
    op = [o for o in m.plan.ops if o.name == "c"][0]
    assert op.engine == rt.ENGINE_TCGEN05 and (op.in0.ld == 16 if unroll else op.in0.wpitch == dhw[2] + 4)
    xs = clips(8, nb, dhw + (c,))
    run(m, [xs])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["in"], nb), dtype=T64)
    assert np.array_equal(xin.numpy(), (xs.astype(np.float64) - 128.0) / 64.0)
    kern, bias = w["c"]
    y = O.relu(O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same"))
    got = m.read_tensor(m.plan.tensors["c"], nb)
    err = np.abs(got - y.numpy()).max() / np.abs(y.numpy()).max()
    assert err <= 2.0 ** -7, "rel err %g" % err


S2D_STEM = [((8, 16, 16), 3, 64, 2), ((9, 21, 19), 3, 64, 2), ((6, 20, 28), 2, 64, 3), ((4, 12, 12), 3, 32, 2),
            ((7, 10, 34), 1, 16, 1)]


@pytest.mark.parametrize("halo,depth", [(True, True), (True, False), (False, False)])
@pytest.mark.parametrize("dhw,c,cout,nb", S2D_STEM)
def test_conv_tcgen05_s2d_stem(dhw, c, cout, nb, halo, depth):
    """7x7x7 / stride 2 'same' stem on the raw clip (I3D Conv3d_1a_7x7, R3D stem) on the tcgen05 engine:
    2x2 space-to-depth cells written by the pre-processing kernel, 4-cell window through an
    overlapping-stride TMA view, k=(7,4,1) stride (2,1,1) with a zero-extended regrouped kernel.
    Even and odd extents (TF 'same' pads (2,3) resp. (3,3)) and C = 1, 2, 3.  depth: 2x2x2 cells (8*C channels without
    padding, k=(4,4,1) stride 1) for C = 1, 2 (see lowering._input for the measured choice)."""
    def build(g):
        x = g.input(dhw + (c,), name="in")
        x = g.conv3d(x, cout, (7, 7, 7), (2, 2, 2), "same", True, None, name="c")
        x = g.bn(x, scale=True, name="b")
        g.relu(x, name="r")
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * c, mean=[128.0] * c, stem_halo=halo, s2d_depth=depth)
    op = [o for o in m.plan.ops if o.name == "c"][0]
    if depth and c in (1, 2):
        assert op.engine == rt.ENGINE_TCGEN05 and op.k == (4, 4, 1) and op.s == (1, 1, 1) and op.in0.C == 32 * c
    else:
        assert op.engine == rt.ENGINE_TCGEN05 and op.k == (7, 4, 1) and op.s == (2, 1, 1)
    assert op.halo == (2 if halo else 0)      # h-halo: one A box per input plane feeds the 4 kh taps
    xs = clips(11, nb, dhw + (c,))
    run(m, [xs])
    xin = torch.as_tensor((xs.astype(np.float64) - 128.0) / 64.0, dtype=T64)      # exact in bf16
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (2, 2, 2), "same")
    y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]])).numpy()
    got = m.read_tensor(m.plan.tensors["r"], nb)
    assert got.shape == y.shape
    err = np.abs(got - y).max() / np.abs(y).max()
    assert err <= 2.0 ** -7, "rel err %g (kc=%d bn=%d brick=%s)" % (err, op.kc, op.bn, op.brick)


@pytest.mark.parametrize("dhw,c,cout,nb", S2D_STEM + [((16, 48, 48), 2, 64, 5), ((10, 30, 44), 1, 32, 3), ((12, 40, 40), 2, 128, 2),
                                            ((16, 56, 56), 3, 64, 4), ((11, 30, 30), 3, 64, 3)])
def test_conv_tcgen05_s2d_stem_shared_weights(dhw, c, cout, nb):
    """Shared-B h-halo mode (groups of 4 tiles of a CTA use one weight block per (fd, chunk), 8 TMEM
    accumulators; taken by the 1- and 2-channel stems whose K chunk is 32), forced on small shapes: full / partial tile groups, several groups per CTA, ragged
    bricks.  Must be bit-identical to the per-tile weight streaming (same MMA sequence per tile)."""
    def build(g):
        x = g.input(dhw + (c,), name="in")
        x = g.conv3d(x, cout, (7, 7, 7), (2, 2, 2), "same", True, None, name="c")
        x = g.bn(x, scale=True, name="b")
        g.relu(x, name="r")
    xs = clips(11, nb, dhw + (c,))
    tune("bshare_min_tiles", 0)          # off
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * c, mean=[128.0] * c)
    run(m, [xs])
    ref = m.read_tensor(m.plan.tensors["r"], nb)
    del m
    tune("bshare_min_tiles", 1)          # always
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * c, mean=[128.0] * c)
    assert [o for o in m.plan.ops if o.name == "c"][0].halo == 2
    run(m, [xs])
    got = m.read_tensor(m.plan.tensors["r"], nb)
    assert np.array_equal(got, ref)
    xin = torch.as_tensor((xs.astype(np.float64) - 128.0) / 64.0, dtype=T64)
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (2, 2, 2), "same")
    y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]])).numpy()
    assert np.abs(got - y).max() / np.abs(y).max() <= 2.0 ** -7


@pytest.mark.parametrize("dhw,c,cout,nb,depth", [((8, 16, 16), 3, 64, 2, False), ((9, 21, 19), 3, 64, 3, False),
                                                 ((6, 20, 28), 2, 64, 3, True), ((4, 12, 12), 3, 32, 2, False),
                                                 ((16, 48, 48), 2, 64, 5, False), ((11, 30, 30), 3, 64, 3, False),
                                                 ((7, 10, 34), 1, 16, 1, True), ((12, 40, 40), 2, 128, 2, True)])
def test_conv_tcgen05_stem_cta_pair(dhw, c, cout, nb, depth):
    """CTA-pair mode (tcgen05 cta_group::2: one 256 x N x 16 MMA per two CTAs, every CTA stages its own A box and half
    of the weight rows; conv_tc2.cu), forced on small shapes: even and odd tile counts (the last CTA of an odd count
    re-runs a tile without storing it), several tile pairs per cluster, ragged bricks, kc = 32 and 64, N = 16 .. 128.
    Every tile accumulates the same products in the same order as the single-CTA h-halo path -> bit-identical."""
    def build(g):
        x = g.input(dhw + (c,), name="in")
        x = g.conv3d(x, cout, (7, 7, 7), (2, 2, 2), "same", True, None, name="c")
        x = g.bn(x, scale=True, name="b")
        g.relu(x, name="r")
    xs = clips(11, nb, dhw + (c,))
    tune("bshare_min_tiles", 0)
    tune("pair_min_tiles", 0)            # off
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * c, mean=[128.0] * c, s2d_depth=depth)
    run(m, [xs])
    ref = m.read_tensor(m.plan.tensors["r"], nb)
    del m
    tune("pair_min_tiles", 1)            # always
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * c, mean=[128.0] * c, s2d_depth=depth)
    assert [o for o in m.plan.ops if o.name == "c"][0].halo == 2
    run(m, [xs])
    got = m.read_tensor(m.plan.tensors["r"], nb)
    xin = torch.as_tensor((xs.astype(np.float64) - 128.0) / 64.0, dtype=T64)
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (2, 2, 2), "same")
    y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]])).numpy()
    assert np.abs(got - y).max() / np.abs(y).max() <= 2.0 ** -7
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("dhw,cin,cout,k,nb,res", [((2, 7, 7), 512, 512, (3, 3, 3), 2, False), ((1, 4, 4), 256, 512, (3, 3, 3), 3, True),
                                                   ((1, 1, 1), 4096, 4096, (1, 1, 1), 5, False), ((2, 7, 7), 96, 208, (3, 3, 3), 1, True)])
def test_conv_tcgen05_split_k(dhw, cin, cout, k, nb, res):
    """Split-K (layers with fewer tiles than SMs: C3D conv5 at small batches, the Dense layers, R3D stages 3-4): every
    tile's K loop is spread over several CTAs, fp32 partial tiles are summed in order by splitk_reduce_kernel, which
    also applies BN / ReLU / the residual add and the second output.  Against the oracle, and against the unsplit
    lowering (same products, different fp32 summation order)."""
    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, cin, (1, 1, 1), (1, 1, 1), "same", True, "relu", name="pre")
        y = g.conv3d(x, cout, k, (1, 1, 1), "same", True, None, name="c")
        if res:
            sc = g.conv3d(x, cout, (1, 1, 1), (1, 1, 1), "same", True, None, name="sc")
            y = g.add_([sc, y], name="add")
        y = g.bn(y, scale=True, name="b")
        g.relu(y, name="r")
    outs = {}
    for split in (True, False):
        g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * 3, mean=[128.0] * 3, split_k=split)
        op = [o for o in m.plan.ops if o.name == "c"][0]
        assert op.engine == rt.ENGINE_TCGEN05 and (op.ksplit > 1) == split, (op.ksplit, op.bn, op.brick)
        xs = clips(4, nb, dhw + (3,))
        run(m, [xs])
        outs[split] = m.read_tensor(m.plan.tensors["r"], nb)
        if split:
            xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], nb), dtype=T64)
            kern, bias = w["c"]
            y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same")
            if res:
                ks, bs = w["sc"]
                y = y + bf16_round(O.conv3d(xin, bf16_round(ks), torch.as_tensor(bs, dtype=T64), (1, 1, 1), "same").numpy())
            y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]])).numpy()
            err = np.abs(outs[True] - y).max() / max(np.abs(y).max(), 1e-6)
            assert err <= 2.0 ** -7, "rel err %g (ksplit=%d bn=%d brick=%s)" % (err, op.ksplit, op.bn, op.brick)
        del m
    assert np.abs(outs[True] - outs[False]).max() <= 2.0 ** -7 * np.abs(outs[False]).max()


@pytest.mark.parametrize("dhw,cin,cout,nb,res", [((4, 28, 28), 64, 64, 2, True), ((3, 12, 20), 64, 128, 3, False),
                                                 ((2, 9, 13), 128, 64, 2, True), ((5, 16, 16), 64, 32, 1, True)])
def test_conv_tcgen05_pair_halo_3x3x3(dhw, cin, cout, nb, res):
    """3x3x3 / stride-1 convs with one small N tile on the CTA-pair kernel (tc_halo = 3: one haloed A box per (fd, fw)
    feeds the kh taps, every CTA loads half of the weight rows, 256 x N x 16 MMAs) incl. the residual add + second
    BN-ReLU output of the pre-activation ResNet blocks (R3D stage 1, train.py:1368-1393).  Ragged bricks, odd tile
    counts, 1 and 2 channel chunks.  Against the oracle and against the generic lowering."""
    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, cin, (1, 1, 1), (1, 1, 1), "same", True, "relu", name="pre")
        y = g.conv3d(x, cout, (3, 3, 3), (1, 1, 1), "same", True, None, name="c")
        if res:
            sc = g.conv3d(x, cout, (1, 1, 1), (1, 1, 1), "same", True, None, name="sc")
            y = g.add_([sc, y], name="add")
        y = g.bn(y, scale=True, name="b")
        g.relu(y, name="r")
    outs = {}
    xs = clips(4, nb, dhw + (3,))
    for mode in ("always", False):
        g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * 3, mean=[128.0] * 3, pair_halo=mode, split_k=False, balance_n=False)
        op = [o for o in m.plan.ops if o.name == "c"][0]
        assert op.engine == rt.ENGINE_TCGEN05 and op.halo == (3 if mode else 0), (op.halo, op.kc, op.bn)
        run(m, [xs])
        outs[mode] = m.read_tensor(m.plan.tensors["r"], nb)
        if res:
            outs[(mode, "add")] = m.read_tensor(m.plan.tensors["add"], nb)
        if mode:
            xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], nb), dtype=T64)
            kern, bias = w["c"]
            y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same")
            if res:
                ks, bs = w["sc"]
                y = y + bf16_round(O.conv3d(xin, bf16_round(ks), torch.as_tensor(bs, dtype=T64), (1, 1, 1), "same").numpy())
            y = O.relu(O.batchnorm(y, *[torch.as_tensor(a, dtype=T64) for a in w["b"]])).numpy()
            err = np.abs(outs[mode] - y).max() / max(np.abs(y).max(), 1e-6)
            assert err <= 2.0 ** -7, "rel err %g (bn=%d brick=%s)" % (err, op.bn, op.brick)
        del m
    # same products, same K order per tile (fd, fh, fw, chunk vs fd, fw, chunk, fh differ) -> equal up to fp32 summation order
    assert np.abs(outs["always"] - outs[False]).max() <= 2.0 ** -7 * np.abs(outs[False]).max()
    if res:
        assert np.abs(outs[("always", "add")] - outs[(False, "add")]).max() <= 2.0 ** -7 * np.abs(outs[(False, "add")]).max()


PAIR_POOL = [((4, 16, 16), 3, True), ((3, 12, 40), 3, True), ((5, 8, 24), 2, False), ((2, 34, 18), 3, True),
             ((3, 16, 112), 4, True)]


@pytest.mark.parametrize("dhw,c,relu", PAIR_POOL)
def test_conv_tcgen05_pair_pool_stem(dhw, c, relu):
    """C3D conv1 + pool1 (train.py:1230-1233) as one op: pair-unrolled clip, GEMM row = 2 output pixels
    (N = 128), MaxPooling3D (1,2,2) in registers (max over the pair's column halves, shuffle with the
    row partner).  Ragged tiles in H and W, C = 2/3/4, with and without ReLU (signed outputs)."""
    def build(g):
        x = g.input(dhw + (c,), name="in")
        x = g.conv3d(x, 64, (3, 3, 3), (1, 1, 1), "same", True, "relu" if relu else None, name="c")
        g.maxpool(x, (1, 2, 2), (1, 2, 2), "valid", name="p")
    g, w, m = make_member(build, "bf16", 3, scale=[1 / 64.0] * c, mean=[128.0] * c)
    assert len(m.plan.ops) == 2
    op = m.plan.ops[1]
    assert op.engine == rt.ENGINE_TCGEN05 and op.pair_pool == 1 and op.bn == 128 and op.halo == 1
    xs = clips(21, 3, dhw + (c,))
    run(m, [xs])
    xin = torch.as_tensor((xs.astype(np.float64) - 128.0) / 64.0, dtype=T64)      # exact in bf16
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same")
    if relu:
        y = O.relu(y)
    exp = O.maxpool3d(bf16_round(y.numpy()), (1, 2, 2), (1, 2, 2), "valid").numpy()
    got = m.read_tensor(m.plan.tensors["p"], 3)
    assert got.shape == exp.shape
    err = np.abs(got - exp).max() / np.abs(exp).max()
    assert err <= 2.0 ** -7, "rel err %g brick=%s" % (err, op.brick)
    # four epilogue groups (four 128-column accumulators drained in turn; opt-in) vs the default two: same bits
    del m
    tune("stem_groups", 4)
    g, w2, m2 = make_member(build, "bf16", 3, scale=[1 / 64.0] * c, mean=[128.0] * c)
    run(m2, [xs])
    assert np.array_equal(m2.read_tensor(m2.plan.tensors["p"], 3), got)


def test_pair_pool_stem_falls_back_on_odd_width():
    def build(g):
        x = g.input((4, 10, 15, 3), name="in")
        x = g.conv3d(x, 64, (3, 3, 3), (1, 1, 1), "same", True, "relu", name="c")
        g.maxpool(x, (1, 2, 2), (1, 2, 2), "valid", name="p")
    g, w, m = make_member(build, "bf16", 2, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    op = [o for o in m.plan.ops if o.name == "c"][0]
    assert op.pair_pool == 0 and op.engine == rt.ENGINE_TCGEN05
    xs = clips(22, 2, (4, 10, 15, 3))
    run(m, [xs])
    xin = torch.as_tensor((xs.astype(np.float64) - 128.0) / 64.0, dtype=T64)
    kern, bias = w["c"]
    y = O.relu(O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same"))
    exp = O.maxpool3d(bf16_round(y.numpy()), (1, 2, 2), (1, 2, 2), "valid").numpy()
    got = m.read_tensor(m.plan.tensors["p"], 2)
    assert np.abs(got - exp).max() / np.abs(exp).max() <= 2.0 ** -7


POOL_FUSED = [  # (in dhw, cin, cout, pool k, zeropad, nb)
    ((4, 8, 8), 64, 128, (2, 2, 2), False, 2),
    ((4, 16, 16), 64, 64, (1, 2, 2), False, 2),
    ((2, 7, 7), 64, 512, (2, 2, 2), True, 3),       # C3D conv5b -> zeropad5 -> pool5
    ((4, 14, 14), 32, 96, (2, 2, 2), False, 2),
    ((5, 9, 9), 16, 48, (2, 2, 2), False, 2),       # odd dims: 'valid' pooling drops the last row / plane
]


@pytest.mark.parametrize("dhw,cin,cout,pk,zp,nb", POOL_FUSED)
def test_conv_tcgen05_fused_maxpool(dhw, cin, cout, pk, zp, nb):
    """MaxPooling3D (window == stride, 'valid') fused into the conv epilogue; conv outputs are signed
    (no ReLU) so the 0-valued padding of ZeroPadding3D is distinguishable from -inf padding."""
    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, cin, (1, 1, 1), (1, 1, 1), "same", True, "relu", name="pre")
        x = g.conv3d(x, cout, (3, 3, 3), (1, 1, 1), "same", True, None, name="c")
        if zp:
            x = g.zeropad(x, ((0, 0), (0, 1), (0, 1)), name="z")
        g.maxpool(x, pk, pk, "valid", name="p")
    # (split_k off: at these test sizes the lowering would otherwise prefer split-K over the fused pool)
    g, w, m = make_member(build, "bf16", nb, scale=[1 / 64.0] * 3, mean=[128.0] * 3, split_k=False)
    op = [o for o in m.plan.ops if o.name == "c"][0]
    assert op.engine == rt.ENGINE_TCGEN05 and op.pool_k == tuple(pk) and len(m.plan.ops) == 3
    run(m, [clips(10, nb, dhw + (3,))])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], nb), dtype=T64)
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same")
    y = bf16_round(y.numpy())                       # the kernel pools bf16-rounded conv outputs
    if zp:
        y = O.zeropad3d(y, ((0, 0), (0, 1), (0, 1)))
    exp = O.maxpool3d(y, pk, pk, "valid").numpy()
    got = m.read_tensor(m.plan.tensors["p"], nb)
    assert got.shape == exp.shape
    err = np.abs(got - exp).max() / np.abs(exp).max()
    assert err <= 2.0 ** -7, "rel err %g" % err


STRIDED_TC = [((6, 12, 12), 16, 32, (3, 3, 3), (2, 2, 2), "same"), ((4, 8, 8), 64, 128, (1, 1, 1), (2, 2, 2), "valid"),
              ((1, 7, 7), 64, 128, (1, 1, 1), (1, 2, 2), "valid"), ((5, 9, 9), 32, 64, (3, 3, 3), (2, 2, 2), "same")]


@pytest.mark.parametrize("dhw,cin,cout,k,s,pad", STRIDED_TC)
def test_conv_tcgen05_strided(dhw, cin, cout, k, s, pad):
    """Strided convs (R3D stage transitions / projection shortcuts, train.py:1338-1345, 1355-1357) on
    the tcgen05 engine: TMA elementStrides pick every s-th input position."""
    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, cin, (1, 1, 1), (1, 1, 1), "same", True, "relu", name="pre")
        g.conv3d(x, cout, k, s, pad, True, None, name="c")
    g, w, m = make_member(build, "bf16", 2, scale=[1 / 64.0] * 3, mean=[128.0] * 3, tc_strided=True)
    op = [o for o in m.plan.ops if o.name == "c"][0]
    assert op.engine == rt.ENGINE_TCGEN05
    run(m, [clips(9, 2, dhw + (3,))])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], 2), dtype=T64)
    kern, bias = w["c"]
    y = O.conv3d(xin, bf16_round(kern), torch.as_tensor(bias, dtype=T64), s, pad).numpy()
    got = m.read_tensor(m.plan.tensors["c"], 2)
    err = np.abs(got - y).max() / np.abs(y).max()
    assert err <= 2.0 ** -7, "rel err %g" % err


def test_conv_tcgen05_partial_batch_and_residual():
    """n < max_batch, residual add + second BN-ReLU output (pre-activation ResNet epilogue)."""
    dhw = (2, 6, 6)

    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, 64, (1, 1, 1), (1, 1, 1), "same", True, None, name="pre")
        a = g.relu(g.bn(x, name="bn1"), name="a1")
        c1 = g.conv3d(a, 64, (3, 3, 3), (1, 1, 1), "same", True, None, name="c1")
        a2 = g.relu(g.bn(c1, name="bn2"), name="a2")
        c2 = g.conv3d(a2, 64, (3, 3, 3), (1, 1, 1), "same", True, None, name="c2")
        s = g.add_([x, c2], name="sum")
        g.relu(g.bn(s, name="bn3"), name="a3")
    g, w, m = make_member(build, "bf16", 4, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    op = [o for o in m.plan.ops if o.name == "c2"][0]
    assert op.engine == rt.ENGINE_TCGEN05 and op.in1 is not None and op.out1 is not None
    n = 3
    xs = clips(5, n, dhw + (3,))
    run(m, [xs])
    pre = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], n), dtype=T64)
    a2 = torch.as_tensor(m.read_tensor(m.plan.tensors["a2"], n), dtype=T64)
    kern, bias = w["c2"]
    raw = O.conv3d(a2, bf16_round(kern), torch.as_tensor(bias, dtype=T64), (1, 1, 1), "same") + pre
    act = O.relu(O.batchnorm(raw, *[torch.as_tensor(a, dtype=T64) for a in w["bn3"]]))
    for name, exp in (("sum", raw.numpy()), ("a3", act.numpy())):
        got = m.read_tensor(m.plan.tensors[name], n)
        err = np.abs(got - exp).max() / np.abs(exp).max()
        assert err <= 2.0 ** -7, "%s rel err %g" % (name, err)


# --------------------------------------------------------------------------- pooling
POOL_CASES = [((1, 3, 3), (1, 2, 2), "same"), ((3, 3, 3), (1, 1, 1), "same"), ((3, 3, 3), (2, 2, 2), "same"),
              ((2, 2, 2), (2, 2, 2), "same"), ((2, 2, 2), (2, 2, 2), "valid"), ((1, 2, 2), (1, 2, 2), "valid")]


@pytest.mark.parametrize("k,s,pad", POOL_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_maxpool_exact(k, s, pad, precision):
    dhw = (5, 7, 9)

    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, 16, (1, 1, 1), (1, 1, 1), "same", True, None, name="pre")   # signed values
        g.maxpool(x, k, s, pad, name="p")
    g, w, m = make_member(build, precision, 2, tc=False, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    run(m, [clips(6, 2, dhw + (3,))])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], 2), dtype=T64)
    exp = O.maxpool3d(xin, k, s, pad).numpy()
    got = m.read_tensor(m.plan.tensors["p"], 2)
    assert np.array_equal(got, exp.astype(np.float32))       # max is exact in any precision


@pytest.mark.parametrize("dhw,c,nb", [((9, 28, 28), 48, 2),      # 2 vectors / pixel, 4 items / thread
                                      ((33, 28, 28), 16, 1),     # split into D segments (halo planes re-read)
                                      ((20, 14, 14), 40, 3),     # 4 vectors / pixel, ragged last channel chunk
                                      ((8, 7, 7), 72, 2),        # 8 vectors / pixel
                                      ((1, 5, 3), 8, 2),         # single plane
                                      ((2, 40, 40), 8, 1),       # 1 vector / pixel, one 40-row strip
                                      ((5, 30, 30), 64, 1),      # 28 + 2 rows -> strips of 7/8 rows with halo rows
                                      ((2, 3, 230), 64, 1)])     # row too long for shared memory -> register-blocked kernel
def test_maxpool3s1_plane_sweep_exact(dhw, c, nb):
    """3x3x3 / stride-1 'same' max-pool (Inception branch 3, train.py:1066): plane-sweep kernel, signed inputs."""
    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, c, (1, 1, 1), (1, 1, 1), "same", True, None, name="pre")
        g.maxpool(x, (3, 3, 3), (1, 1, 1), "same", name="p")
    g, w, m = make_member(build, "bf16", nb, tc=False, scale=[1 / 64.0] * 3, mean=[128.0] * 3)
    run(m, [clips(16, nb, dhw + (3,))])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], nb), dtype=T64)
    exp = O.maxpool3d(xin, (3, 3, 3), (1, 1, 1), "same").numpy()
    assert np.array_equal(m.read_tensor(m.plan.tensors["p"], nb), exp.astype(np.float32))


def test_zeropad_maxpool_and_avgpool():
    dhw = (2, 7, 7)

    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, 16, (1, 1, 1), (1, 1, 1), "same", True, None, name="pre")
        z = g.zeropad(x, ((0, 0), (0, 1), (0, 1)), name="z")
        g.maxpool(z, (2, 2, 2), (2, 2, 2), "valid", name="p")
        g.avgpool(x, (2, 7, 7), (1, 1, 1), "valid", name="avg")
    g, w, m = make_member(build, "fp32", 2, scale=[1 / 64.0] * 3, mean=[200.0] * 3)   # mostly negative
    run(m, [clips(7, 2, dhw + (3,))])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], 2), dtype=T64)
    exp = O.maxpool3d(O.zeropad3d(xin, ((0, 0), (0, 1), (0, 1))), (2, 2, 2), (2, 2, 2), "valid").numpy()
    assert np.array_equal(m.read_tensor(m.plan.tensors["p"], 2), exp.astype(np.float32))
    avg = O.avgpool3d(xin, (2, 7, 7)).numpy()
    np.testing.assert_allclose(m.read_tensor(m.plan.tensors["avg"], 2), avg, rtol=1e-5, atol=1e-6)


def test_zeropad_maxpool_bf16_wblocked():
    """ZeroPadding3D + MaxPooling3D on the stand-alone W-blocked bf16 kernel (0-valued padding must win
    over negative activations), ragged W (9 outputs -> partial last block)."""
    dhw = (3, 6, 17)

    def build(g):
        x = g.input(dhw + (3,), name="in")
        x = g.conv3d(x, 24, (1, 1, 1), (1, 1, 1), "same", True, None, name="pre")
        z = g.zeropad(x, ((0, 0), (0, 1), (0, 1)), name="z")
        g.maxpool(z, (2, 2, 2), (2, 2, 2), "valid", name="p")
    g, w, m = make_member(build, "bf16", 2, tc=False, scale=[1 / 64.0] * 3, mean=[200.0] * 3)   # mostly negative
    run(m, [clips(7, 2, dhw + (3,))])
    xin = torch.as_tensor(m.read_tensor(m.plan.tensors["pre"], 2), dtype=T64)
    exp = O.maxpool3d(O.zeropad3d(xin, ((0, 0), (0, 1), (0, 1))), (2, 2, 2), (2, 2, 2), "valid").numpy()
    assert np.array_equal(m.read_tensor(m.plan.tensors["p"], 2), exp.astype(np.float32))


# --------------------------------------------------------------------------- vote
def test_vote_bit_exact_vs_oracle():
    rng = np.random.default_rng(11)
    for (mm, n, c) in [(4, 300, 11), (12, 1000, 11), (1, 7, 11), (32, 129, 11), (4, 128, 3)]:
        z = rng.standard_normal((mm, n, c)) * 3
        p = np.exp(z - z.max(-1, keepdims=True))
        p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
        p64 = np.stack([OV.csv_roundtrip(p[j]) for j in range(mm)])       # what the reference votes on
        wts = rng.uniform(0.1, 1, mm)
        wts /= wts.sum()
        d64 = torch.from_numpy(p64).cuda()
        for mode, wv in (("SUM", None), ("WEIGHTED", wts), ("MAXIMUM", None)):
            ref = OV.ensemble_predictions(p64, "MAXIMUM" if mode == "MAXIMUM" else (np.ones(mm) if wv is None else wv))
            got = rt.vote(d64, None if wv is None else torch.from_numpy(wv), mode).cpu().numpy()
            assert np.array_equal(got, ref.astype(np.int32)), (mm, n, c, mode)
        # fp32 probabilities straight from the model (in-memory fast path)
        got32 = rt.vote(torch.from_numpy(p).cuda(), None, "SUM").cpu().numpy()
        ref32 = OV.ensemble_predictions(p.astype(np.float64), np.ones(mm))
        assert np.array_equal(got32, ref32.astype(np.int32))


def test_vote_ties_and_summed():
    rng = np.random.default_rng(12)
    p = (rng.integers(0, 4, (4, 257, 11)) / 8.0)
    pred, summed = rt.vote(torch.from_numpy(p).cuda(), None, "SUM", return_summed=True)
    assert np.array_equal(pred.cpu().numpy(), OV.ensemble_predictions(p, np.ones(4)).astype(np.int32))
    assert np.array_equal(summed.cpu().numpy(), OV.summed_probabilities(p, np.ones(4)))
    got = rt.vote(torch.from_numpy(p).cuda(), None, "MAXIMUM").cpu().numpy()
    assert np.array_equal(got, OV.ensemble_predictions(p, "MAXIMUM").astype(np.int32))


def test_vote_search_counts():
    rng = np.random.default_rng(13)
    m, n, c, wn = 4, 500, 11, 64
    p = rng.random((m, n, c))
    p /= p.sum(-1, keepdims=True)
    labels = rng.integers(0, c, n).astype(np.int32)
    wm = rng.random((wn, m))
    got = rt.vote_search(torch.from_numpy(p).cuda(), torch.from_numpy(wm).cuda(),
                         torch.from_numpy(labels).cuda()).cpu().numpy()
    exp = np.array([(OV.ensemble_predictions(p, wm[i]) == labels).sum() for i in range(wn)])
    assert np.array_equal(got, exp)


# --------------------------------------------------------------------------- clip assembly
def test_assemble_clip_bit_exact_vs_oracle_and_cv2():
    """cse_assemble_clip (select_frames + cv2.resize INTER_LINEAR on uint8, train.py:132-145, 286) vs the
    numpy restatement (itself pinned against cv2 in tests/test_oracle_resize.py): bit-exact."""
    from oracle import resize as R
    rng = np.random.default_rng(21)
    cases = [(37, 90, 122, 3, 16, 112, 112), (16, 360, 640, 3, 16, 112, 112), (70, 120, 160, 3, 64, 224, 224),
             (25, 48, 64, 1, 20, 224, 224), (9, 33, 17, 2, 4, 7, 150), (5, 1, 1, 4, 5, 3, 3), (40, 224, 224, 3, 20, 224, 224),
             (33, 448, 448, 3, 16, 224, 224), (7, 50, 300, 3, 7, 131, 9)]
    for (n, hs, ws, c, t, h, w) in cases:
        frames = rng.integers(0, 256, (n, hs, ws, c) if c > 1 else (n, hs, ws), dtype=np.uint8)
        got = rt.assemble_clip(torch.from_numpy(frames).cuda(), t, h, w).cpu().numpy()
        exp = R.assemble_clip(list(frames), t, h, w)
        assert got.shape == exp.shape and np.array_equal(got, exp), (n, hs, ws, c, t, h, w)
    with pytest.raises(rt.CseError):          # select_frames would keep 3 < 16 frames
        rt.assemble_clip(torch.zeros((3, 8, 8, 3), dtype=torch.uint8).cuda(), 16, 8, 8)


def test_assemble_clip_equals_reference_golden_and_feeds_predict():
    """The committed videos, decoded on the host, assembled on the GPU == the reference's own
    get_onestream_videoclip / get_twostream_videoclip outputs (tests/golden/clips_golden.npz)."""
    import hashlib
    import os
    from cse_b200 import clips
    gold = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(gold, "clips_golden.npz"))
    dev = torch.device("cuda", torch.cuda.current_device())
    for tag in ("small", "up", "c3d", "i3d"):
        t, h, w = (int(v) for v in g["shape_" + tag])
        rgb = clips.load_rgb_clip(os.path.join(gold, "clip_rgb.avi"), t, h, w, device=dev)
        flow = clips.load_flow_clip(os.path.join(gold, "clip_flow_x.avi"), os.path.join(gold, "clip_flow_y.avi"), t, h, w,
                                    device=dev)
        assert rgb.is_cuda and flow.is_cuda
        for arr, key in ((rgb, "sha_rgb_"), (flow, "sha_flow_")):
            digest = np.frombuffer(hashlib.sha256(arr.cpu().numpy().tobytes()).digest(), np.uint8)
            assert np.array_equal(digest, g[key + tag]), (tag, key)
    # a GPU-assembled clip goes straight into Member.predict and gives the same probabilities as the numpy clip
    m = Member(G.build_model_graph("C3D", (16, 112, 112, 3), 11), synthetic_weights(G.build_model_graph("C3D", (16, 112, 112, 3), 11), seed=3),
               precision="bf16", max_batch=2)
    clip_dev = clips.load_rgb_clip(os.path.join(gold, "clip_rgb.avi"), 16, 112, 112, device=dev)[None]
    clip_np = clips.load_rgb_clip(os.path.join(gold, "clip_rgb.avi"), 16, 112, 112)[None]
    assert np.array_equal(m.predict(clip_dev), m.predict(clip_np))
