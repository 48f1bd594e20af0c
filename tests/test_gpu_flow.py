"""GPU Farneback flow (csrc/flow.cu) through the C ABI: bit-identical to oracle/farneback.py, within 1e-4
pixel of cv2.calcOpticalFlowFarneback as the reference calls it (train.py:320-322), and the whole loader branch
(scale, gray, flow, select_frames, resize) against the golden made by the reference's own extractor."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")

from cse_b200 import clips as CL            # noqa: E402
from cse_b200 import runtime as rt          # noqa: E402
from oracle import farneback as FB          # noqa: E402
from test_oracle_farneback import GOLD, VIDEO, moving_scene, video_frames       # noqa: E402

DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_integer_stages_bit_exact():
    frames = video_frames()
    factor = 224 / max(frames[0].shape)
    f = dev(np.stack(frames[:4]))
    gray = rt.bgr2gray(f)
    assert np.array_equal(gray.cpu().numpy(), np.stack([cv2.cvtColor(x, cv2.COLOR_BGR2GRAY) for x in frames[:4]]))
    assert np.array_equal(rt.resize_u8(f, fx=factor, fy=factor).cpu().numpy(),
                          np.stack([cv2.resize(x, None, fx=factor, fy=factor) for x in frames[:4]]))
    assert np.array_equal(rt.resize_u8(gray, fx=factor, fy=factor).cpu().numpy(),
                          np.stack([cv2.resize(cv2.cvtColor(x, cv2.COLOR_BGR2GRAY), None, fx=factor, fy=factor) for x in frames[:4]]))
    img = np.random.default_rng(5).integers(0, 256, (2, 360, 640, 3)).astype(np.uint8)
    factor = 224 / 640
    assert np.array_equal(rt.resize_u8(dev(img), fx=factor, fy=factor).cpu().numpy(),
                          np.stack([cv2.resize(x, None, fx=factor, fy=factor) for x in img]))
    assert np.array_equal(rt.resize_u8(dev(img), 112, 112).cpu().numpy(), np.stack([cv2.resize(x, (112, 112)) for x in img]))
    with pytest.raises(rt.CseError):
        rt.check(rt.load_library().cse_resize_u8(f.data_ptr(), 4, 90, 122, 3, f.data_ptr(), 100, 100, 0.5, 0.5, None))


@pytest.mark.parametrize("shape,dst", [((3, 165, 224, 2), (224, 224)), ((2, 165, 224, 2), (28, 36)), ((1, 56, 56, 2), (112, 112)),
                                       ((2, 63, 112), (126, 224)), ((2, 224, 224), (112, 112))])
def test_float_resize_bit_exact(shape, dst):
    img = np.random.default_rng(1).standard_normal(shape).astype(np.float32) * 3
    got = rt.resize_linear_f32(dev(img), *dst).cpu().numpy()
    assert np.array_equal(got, np.stack([FB.resize_linear(x, *dst) for x in img]))
    if len(shape) == 4:                    # the 2-channel path of cv2 is reproducible, its 1-channel (IPP) path is not
        assert np.array_equal(got, np.stack([cv2.resize(x, (dst[1], dst[0])) for x in img]))


@pytest.mark.parametrize("h,w,shift", [(126, 224, (1, 2)), (224, 224, (3, -2)), (64, 80, (0, 1)), (45, 61, (1, 0)), (165, 224, (2, -3))])
def test_flow_pair_matches_oracle_and_cv2(h, w, shift):
    a, b = moving_scene(np.random.default_rng(h * w), h, w, shift)
    got = rt.farneback(dev(np.stack([a, b]))).cpu().numpy()
    assert got.shape == (1, h, w, 2) and got.dtype == np.float32
    want = FB.calc_optical_flow_farneback(a, b)
    assert np.array_equal(got[0], want), "max |gpu - oracle| = %g" % np.abs(got[0] - want).max()
    ref = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 5, 11, 5, 5, 1.1, 0)
    assert np.abs(ref).max() > 0.5
    assert np.abs(got[0] - ref).max() <= (1e-4 if h % 4 == 0 and w % 4 == 0 or min(h, w) < 64 else 1e-3), np.abs(got[0] - ref).max()


def test_flow_batch_equals_pairs_and_other_parameters():
    """One call over F frames = F - 1 independent pair calls; non-default pyramid / window / polynomial parameters."""
    rng = np.random.default_rng(21)
    a, b = moving_scene(rng, 96, 128, (1, -1))
    c, d = moving_scene(rng, 96, 128, (0, 2))
    frames = np.stack([a, b, c, d, a])
    full = rt.farneback(dev(frames)).cpu().numpy()
    assert full.shape == (4, 96, 128, 2)
    for i in range(4):
        assert np.array_equal(full[i], rt.farneback(dev(frames[i:i + 2])).cpu().numpy()[0])
    assert np.array_equal(full[2], FB.calc_optical_flow_farneback(c, d))
    for kw in (dict(pyr_scale=0.5, levels=1, winsize=7, iterations=2, poly_n=5, poly_sigma=1.1),
               dict(pyr_scale=0.8, levels=3, winsize=15, iterations=3, poly_n=7, poly_sigma=1.5)):
        got = rt.farneback(dev(np.stack([a, b])), **kw).cpu().numpy()[0]
        assert np.array_equal(got, FB.calc_optical_flow_farneback(a, b, **kw))
        ref = cv2.calcOpticalFlowFarneback(a, b, None, kw["pyr_scale"], kw["levels"], kw["winsize"], kw["iterations"], kw["poly_n"],
                                           kw["poly_sigma"], 0)
        assert np.abs(got - ref).max() <= 1e-4


def test_flow_long_video_matches_cv2():
    """One call over a 33-frame 224 x 224 video (the I3D working size: three pyramid levels, 32 pairs on blockIdx.y):
    spot-checked pairs against cv2 itself and against single-pair calls."""
    rng = np.random.default_rng(7)
    base = cv2.GaussianBlur(rng.integers(0, 256, (224 + 80, 224 + 80)).astype(np.float32), (9, 9), 2.5)
    gray = np.stack([np.clip(base[i:i + 224, 2 * i:2 * i + 224], 0, 255).astype(np.uint8) for i in range(33)])
    flow = rt.farneback(dev(gray)).cpu().numpy()
    assert flow.shape == (32, 224, 224, 2)
    for i in (0, 13, 31):
        ref = cv2.calcOpticalFlowFarneback(gray[i], gray[i + 1], None, 0.5, 5, 11, 5, 5, 1.1, 0)
        assert np.abs(ref).max() > 1.0 and np.abs(flow[i] - ref).max() <= 1e-4, (i, np.abs(flow[i] - ref).max())
        assert np.array_equal(flow[i], rt.farneback(dev(gray[i:i + 2])).cpu().numpy()[0])


def test_flow_argument_errors():
    lib = rt.load_library()
    g = dev(np.zeros((2, 64, 64), np.uint8))
    out = torch.empty((1, 64, 64, 2), dtype=torch.float32, device=DEV)
    work = torch.empty(lib.cse_farneback_workspace_bytes(2, 64, 64) // 8 + 1, dtype=torch.float64, device=DEV)
    assert lib.cse_farneback_workspace_bytes(1, 64, 64) == 0
    with pytest.raises(rt.CseError):        # one frame: no pair
        rt.check(lib.cse_farneback(g.data_ptr(), 1, 64, 64, 0.5, 5, 11, 5, 5, 1.1, out.data_ptr(), work.data_ptr(), work.numel() * 8, None))
    with pytest.raises(rt.CseError):        # workspace too small
        rt.check(lib.cse_farneback(g.data_ptr(), 2, 64, 64, 0.5, 5, 11, 5, 5, 1.1, out.data_ptr(), work.data_ptr(), 1024, None))
    with pytest.raises(rt.CseError):        # polynomial neighbourhood beyond the kernel's tap arrays
        rt.check(lib.cse_farneback(g.data_ptr(), 2, 64, 64, 0.5, 5, 11, 5, 40, 1.1, out.data_ptr(), work.data_ptr(), work.numel() * 8, None))
    rt.check(lib.cse_farneback(g.data_ptr(), 2, 64, 64, 0.5, 5, 11, 5, 5, 1.1, out.data_ptr(), work.data_ptr(), work.numel() * 8, None))
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0    # two identical flat frames: no motion


def test_flow_clip_matches_reference_golden():
    """load_farneback_twostream_clip(device=cuda): rgb bit-exact, flow within 1e-3 pixel of what the reference's own
    get_twostream_videoclip(..., 'FarneBack_onTheFly') returned (flows reach 8.6 pixels on this video)."""
    t, h, w = (int(v) for v in GOLD["shape_small"])
    rgb, flow = CL.load_farneback_twostream_clip(VIDEO, t, h, w, device=DEV)
    assert rgb.is_cuda and flow.is_cuda and flow.dtype == torch.float32 and tuple(flow.shape) == (t, h, w, 2)
    assert np.array_equal(rgb.cpu().numpy(), GOLD["rgb_small"])
    err = np.abs(flow.cpu().numpy() - GOLD["flow_small"]).max()
    assert err <= 1e-3, err
    t, h, w = (int(v) for v in GOLD["shape_i3d"])
    rgb, flow = CL.load_farneback_twostream_clip(VIDEO, t, h, w, device=DEV)
    err = np.abs(flow.cpu().numpy()[:, ::8, ::8] - GOLD["flow_i3d_sub"]).max()
    assert err <= 1e-3, err
    # and bit-identical to the oracle's restatement of the same branch on the first fields
    frames = video_frames()
    flows = FB.farneback_flow(frames[:3])
    host = CL.farneback_flow_clip_device(frames[:3], 2, h, w, DEV).cpu().numpy()
    assert np.array_equal(host, np.stack([FB.resize_linear(f, h, w) for f in flows]))


def test_clip_sequence_farneback_on_device():
    df = pd.DataFrame({"rgbclips_path": [VIDEO, VIDEO, VIDEO], "x_axis_flowclips_path": ["", "", ""],
                       "y_axis_flowclips_path": ["", "", ""], "class": [0, 3, 5]})
    seq = CL.ClipSequence(df, "TWOSTREAM_I3D", (8, 28, 36, 0), 11, batch_size=2, optical_flow_status="FarneBack_onTheFly", device=DEV)
    got = [b for b, _ in CL.iterate_batches(seq, 0, 1, workers=2)]
    rgb, flow = got[0]
    assert rgb.is_cuda and flow.is_cuda and tuple(rgb.shape) == (2, 8, 28, 36, 3) and tuple(flow.shape) == (2, 8, 28, 36, 2)
    assert np.array_equal(rgb[1].cpu().numpy(), GOLD["rgb_small"])
    assert np.abs(flow[1].cpu().numpy() - GOLD["flow_small"]).max() <= 1e-3
    assert tuple(got[1][1].shape) == (1, 8, 28, 36, 2)
