"""On-the-fly Farneback flow of the TwoStream variant (train.py:294-332, 223-239): the loader-side restatement in
cse_b200/clips.py against outputs of the reference's own functions (tools/make_golden_farneback.py, same OpenCV)."""
import hashlib
import os

import numpy as np
import pandas as pd
import pytest

from cse_b200 import clips as CL

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "farneback_golden.npz"))
VIDEO = os.path.join(HERE, "golden", "clip_rgb.avi")


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_farneback_clip_matches_reference_small():
    t, h, w = (int(v) for v in GOLD["shape_small"])
    rgb, flow = CL.load_farneback_twostream_clip(VIDEO, t, h, w)
    assert rgb.dtype == np.uint8 and flow.dtype == np.float32
    assert np.array_equal(rgb, GOLD["rgb_small"])
    assert np.array_equal(flow, GOLD["flow_small"]), "max |diff| %g" % np.abs(flow - GOLD["flow_small"]).max()
    assert float(np.abs(flow).max()) == float(GOLD["absmax_small"][0]) > 0.0


def test_farneback_clip_matches_reference_i3d_shape():
    t, h, w = (int(v) for v in GOLD["shape_i3d"])
    rgb, flow = CL.load_farneback_twostream_clip(VIDEO, t, h, w)
    assert np.array_equal(_sha(rgb), GOLD["sha_rgb_i3d"])
    assert np.array_equal(_sha(flow), GOLD["sha_flow_i3d"])


def test_clip_sequence_farneback_batches():
    df = pd.DataFrame({"rgbclips_path": [VIDEO, VIDEO, VIDEO], "x_axis_flowclips_path": ["", "", ""],
                       "y_axis_flowclips_path": ["", "", ""], "class": [0, 3, 5]})
    seq = CL.ClipSequence(df, "TWOSTREAM_I3D", (8, 28, 36, 0), 11, batch_size=2, optical_flow_status="FarneBack_onTheFly")
    assert len(seq) == 2
    (rgb, flow), y = seq[0]
    assert rgb.shape == (2, 8, 28, 36, 3) and rgb.dtype == np.uint8
    assert flow.shape == (2, 8, 28, 36, 2) and flow.dtype == np.float32
    assert np.array_equal(flow[1], GOLD["flow_small"]) and y.shape == (2, 11) and y[1, 3] == 1.0
    got = [b for b, _ in CL.iterate_batches(seq, 0, 1, workers=2)]
    assert np.array_equal(got[1][1][0], GOLD["flow_small"])
    with pytest.raises(ValueError):
        CL.ClipSequence(df, "TWOSTREAM_I3D", (8, 28, 36, 0), 11, optical_flow_status="TVL1_onTheFly")
