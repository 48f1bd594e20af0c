"""Clip assembly oracle (oracle/resize.py) pinned against (1) the real cv2.resize of the container's
OpenCV and (2) outputs of the reference's own get_onestream_videoclip / get_twostream_videoclip /
select_frames (tests/golden/clips_golden.npz, made by tools/make_golden_clips.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import resize as R
from cse_b200 import clips

GOLD = os.path.join(os.path.dirname(__file__), "golden")
cv2 = pytest.importorskip("cv2")


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_resize_restatement_equals_cv2_bit_exact():
    rng = np.random.default_rng(5)
    for _ in range(120):
        hs, ws = (int(v) for v in rng.integers(1, 300, 2))
        h, w = (int(v) for v in rng.integers(1, 260, 2))
        c = int(rng.choice([1, 2, 3, 4]))
        img = rng.integers(0, 256, (hs, ws, c) if c > 1 else (hs, ws), dtype=np.uint8)
        assert np.array_equal(R.resize_linear_u8(img, w, h), cv2.resize(img, (w, h))), (hs, ws, h, w, c)


@pytest.mark.parametrize("src,dst", [((360, 640), (112, 112)), ((360, 640), (224, 224)), ((240, 320), (112, 112)),
                                     ((224, 224), (224, 224)), ((448, 448), (224, 224)), ((112, 112), (224, 224)),
                                     ((2, 2), (7, 5)), ((1, 1), (4, 4)), ((300, 5), (3, 300))])
def test_resize_model_shapes_and_edges(src, dst):
    img = np.random.default_rng(7).integers(0, 256, src + (3,), dtype=np.uint8)
    assert np.array_equal(R.resize_linear_u8(img, dst[1], dst[0]), cv2.resize(img, (dst[1], dst[0])))


def test_select_frames_golden():
    g = np.load(os.path.join(GOLD, "clips_golden.npz"))
    flat, pos = g["select_flat"], 0
    for ln in g["select_cases"]:
        row = flat[pos:pos + ln]
        pos += ln
        n, t, kept = int(row[0]), int(row[1]), [int(v) for v in row[2:]]
        assert R.select_frame_indices(n, t) == kept
        assert list(clips.select_frames(list(range(n)), t)) == kept


@pytest.mark.parametrize("tag", ["small", "up", "c3d", "i3d"])
def test_clip_assembly_equals_reference_functions(tag):
    """Decode the committed videos with OpenCV, assemble with the oracle and with the product's host
    loader; both must equal what the reference's own functions returned."""
    g = np.load(os.path.join(GOLD, "clips_golden.npz"))
    t, h, w = (int(v) for v in g["shape_" + tag])
    rgb_path = os.path.join(GOLD, "clip_rgb.avi")
    fx, fy = os.path.join(GOLD, "clip_flow_x.avi"), os.path.join(GOLD, "clip_flow_y.avi")
    frames = clips.decode_frames(rgb_path)
    rgb = R.assemble_clip(frames, t, h, w)
    flow = np.stack([R.assemble_clip(clips.decode_frames(p, gray=True), t, h, w) for p in (fx, fy)], axis=-1)
    assert np.array_equal(sha(rgb), g["sha_rgb_" + tag]) and np.array_equal(sha(flow), g["sha_flow_" + tag])
    if "rgb_" + tag in g:
        assert np.array_equal(rgb, g["rgb_" + tag]) and np.array_equal(flow, g["flow_" + tag])
    assert np.array_equal(clips.load_rgb_clip(rgb_path, t, h, w), rgb)
    assert np.array_equal(clips.load_flow_clip(fx, fy, t, h, w), flow)


def test_parallel_decode_keeps_order_and_bytes(tmp_path):
    """clips.iterate_batches with worker threads (the reference's `workers`): same batches, same order."""
    import pandas as pd
    rgb = os.path.join(GOLD, "clip_rgb.avi")
    fx, fy = os.path.join(GOLD, "clip_flow_x.avi"), os.path.join(GOLD, "clip_flow_y.avi")
    # different clips = different pre-decoded frame ranges of the same video
    frames = np.asarray(clips.decode_frames(rgb))
    paths = []
    for k in range(7):
        p = str(tmp_path / ("clip%d.npy" % k))
        np.save(p, frames[k:k + 20])
        paths.append(p)
    data = pd.DataFrame({"rgbclips_path": paths, "x_axis_flowclips_path": [fx] * 7, "y_axis_flowclips_path": [fy] * 7,
                         "class": list(range(7))})
    for mt, shape in (("C3D", (16, 40, 52, 3)), ("TWOSTREAM_I3D", (8, 32, 32, 0))):
        seq = clips.ClipSequence(data, mt, shape, 11, batch_size=2)
        serial = [seq[i] for i in range(len(seq))]
        threaded = list(clips.iterate_batches(seq, 0, len(seq) - 1, workers=3))
        assert len(threaded) == len(serial) == 4
        for (xa, ya), (xb, yb) in zip(serial, threaded):
            xa, xb = (xa if isinstance(xa, list) else [xa]), (xb if isinstance(xb, list) else [xb])
            assert all(np.array_equal(a, b) for a, b in zip(xa, xb)) and np.array_equal(ya, yb)
        assert list(clips.iterate_batches(seq, 2, 1, workers=3)) == []
