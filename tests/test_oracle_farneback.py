"""oracle/farneback.py (the restatement of cv2.calcOpticalFlowFarneback as the reference calls it, train.py:294-332)
pinned against cv2 itself, stage by stage where cv2 exposes the stage, and against the golden made by the reference's
own extractor (tools/make_golden_farneback.py).  Integer stages and the 2-channel float resize are bit-exact; the flow
is held to 1e-4 pixel where the pyramid ratios are integers (measured <= 6e-6) and 1e-3 where they are not (measured
6.5e-5 on the 165 x 224 MJPG golden video): cv2's IPP blur and single-channel resize sum in an order - and with
coefficients - no plain restatement reproduces."""
import os

import numpy as np
import pytest

from oracle import farneback as FB
from oracle import resize as RZ

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "farneback_golden.npz"))
VIDEO = os.path.join(HERE, "golden", "clip_rgb.avi")
FLOW_TOL = 1e-4            # pixels; flows are several pixels


def moving_scene(rng, h, w, shift):
    """Two uint8 frames of one smooth random texture, the second displaced by `shift` = (dy, dx) pixels."""
    base = rng.integers(0, 256, (h // 4 + 8, w // 4 + 8)).astype(np.float32)
    big = cv2.GaussianBlur(cv2.resize(base, (w + 32, h + 32), interpolation=cv2.INTER_CUBIC), (7, 7), 2.0)
    dy, dx = shift
    a = np.clip(big[16:16 + h, 16:16 + w], 0, 255).astype(np.uint8)
    b = np.clip(big[16 + dy:16 + dy + h, 16 + dx:16 + dx + w], 0, 255).astype(np.uint8)
    return a, b


def video_frames():
    cap = cv2.VideoCapture(VIDEO)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            return frames
        frames.append(f)


@pytest.mark.parametrize("ksize,sigma", [(3, 0.0), (3, 0.5), (5, 0.0), (9, 1.5), (21, 3.5), (7, 0.0), (9, 0.0)])
def test_gaussian_kernel_matches_cv2(ksize, sigma):
    assert np.array_equal(FB.gaussian_kernel(ksize, sigma), cv2.getGaussianKernel(ksize, sigma, cv2.CV_32F).ravel())


def test_integer_stages_bit_exact():
    frames = video_frames()
    factor = 224 / max(frames[0].shape)
    for f in frames[:3]:
        assert np.array_equal(FB.bgr2gray(f), cv2.cvtColor(f, cv2.COLOR_BGR2GRAY))
        assert np.array_equal(RZ.resize_linear_u8(f, fx=factor, fy=factor), cv2.resize(f, None, fx=factor, fy=factor))
        g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(RZ.resize_linear_u8(g, fx=factor, fy=factor), cv2.resize(g, None, fx=factor, fy=factor))
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (360, 640, 3)).astype(np.uint8)           # a Crowd-11-sized frame, scaled down
    factor = 224 / 640
    assert np.array_equal(RZ.resize_linear_u8(img, fx=factor, fy=factor), cv2.resize(img, None, fx=factor, fy=factor))


@pytest.mark.parametrize("shape,dst", [((165, 224, 2), (224, 224)), ((165, 224, 2), (28, 36)), ((56, 56, 2), (112, 112)),
                                       ((41, 56, 2), (82, 112)), ((126, 224, 2), (224, 224))])
def test_float_resize_two_channels_bit_exact(shape, dst):
    """The flow fields (2 channels: pyramid upsampling and the final resize to the network's size) go through OpenCV's
    own bilinear code, which the restatement reproduces bit for bit."""
    img = np.random.default_rng(1).standard_normal(shape).astype(np.float32) * 3
    assert np.array_equal(FB.resize_linear(img, *dst), cv2.resize(img, (dst[1], dst[0])))


@pytest.mark.parametrize("shape,dst", [((224, 224), (112, 112)), ((126, 224), (63, 112)), ((165, 224), (82, 112)), ((165, 224), (41, 56))])
def test_float_resize_one_channel_close(shape, dst):
    """Single-channel float images (the blurred pyramid levels) are resized by IPP inside cv2: same taps, another
    summation (1 ulp on 40-50 % of the pixels for integer ratios, 3e-5 relative for fractional ones)."""
    img = np.random.default_rng(2).uniform(0, 255, shape).astype(np.float32)
    assert np.abs(FB.resize_linear(img, *dst) - cv2.resize(img, (dst[1], dst[0]))).max() <= 255 * 1e-4


def test_pyramid_levels():
    # 224 x 224: 0.25, 0.5, 1 (0.125 would be 28 < 32 pixels); 126 x 224: 0.5, 1
    assert [(lv[3], lv[4], lv[2]) for lv in FB.pyramid_levels(224, 224, 0.5, 5)] == [(56, 56, 9), (112, 112, 3), (224, 224, 3)]
    assert [(lv[3], lv[4]) for lv in FB.pyramid_levels(126, 224, 0.5, 5)] == [(112, 63), (224, 126)]
    assert [(lv[3], lv[4]) for lv in FB.pyramid_levels(40, 40, 0.5, 5)] == [(40, 40)]


@pytest.mark.parametrize("h,w,shift,tol", [(126, 224, (1, 2), FLOW_TOL), (224, 224, (3, -2), FLOW_TOL), (64, 80, (0, 1), FLOW_TOL),
                                           (224, 168, (-2, 4), FLOW_TOL), (45, 61, (1, 0), FLOW_TOL), (165, 224, (2, -3), 1e-3)])
def test_flow_matches_cv2(h, w, shift, tol):
    a, b = moving_scene(np.random.default_rng(h * w), h, w, shift)
    ref = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 5, 11, 5, 5, 1.1, 0)
    got = FB.calc_optical_flow_farneback(a, b)
    assert got.dtype == np.float32 and got.shape == (h, w, 2)
    assert np.abs(ref).max() > 0.5                                      # there is motion to measure
    assert np.abs(got - ref).max() <= tol, np.abs(got - ref).max()


def test_flow_other_parameters():
    a, b = moving_scene(np.random.default_rng(11), 96, 128, (1, -1))
    for kw in (dict(pyr_scale=0.5, levels=1, winsize=7, iterations=2, poly_n=5, poly_sigma=1.1),
               dict(pyr_scale=0.8, levels=3, winsize=15, iterations=3, poly_n=7, poly_sigma=1.5)):
        ref = cv2.calcOpticalFlowFarneback(a, b, None, kw["pyr_scale"], kw["levels"], kw["winsize"], kw["iterations"], kw["poly_n"],
                                           kw["poly_sigma"], 0)
        assert np.abs(FB.calc_optical_flow_farneback(a, b, **kw) - ref).max() <= FLOW_TOL


def test_flow_clip_matches_reference_golden():
    """The whole loader branch (scale, gray, flow of consecutive frames, select_frames, float resize) on the golden
    video against what the reference's own get_twostream_videoclip returned; 9 of the 37 frames keep the test short."""
    frames = video_frames()
    t, h, w = (int(v) for v in GOLD["shape_small"])
    step = max(1, (len(frames) - 1) // t)
    flows = FB.farneback_flow(frames[:2 * step + 2])                    # flow fields 0 .. 2 * step
    for k in range(3):
        got = FB.resize_linear(flows[k * step], h, w)
        assert np.abs(got - GOLD["flow_small"][k]).max() <= 1e-3, (k, np.abs(got - GOLD["flow_small"][k]).max())
    assert float(GOLD["absmax_small"][0]) > 1.0
