#!/usr/bin/env python
"""Generate tests/golden/farneback_golden.npz by running the REFERENCE's own on-the-fly flow path:
``opticalflow_FarneBack_extractor`` and ``get_twostream_videoclip(..., optical_flow_status='FarneBack_onTheFly')``
are extracted - unmodified - with ``ast`` from /root/reference/train.py and executed with the container's OpenCV on
the small MJPG video tests/golden/clip_rgb.avi (written by tools/make_golden_clips.py).  Only ``cv2.waitKey`` is
stubbed (headless OpenCV).

Run in the build container (needs /root/reference).  Usage: python tools/make_golden_farneback.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_clips as MG        # noqa: E402

MG.WANTED = ["select_frames", "get_twostream_videoclip", "opticalflow_TVL1_retriever", "opticalflow_FarneBack_extractor"]


def main():
    ns = MG.load_reference_functions()
    video = os.path.join(MG.GOLD, "clip_rgb.avi")
    out = {}
    for tag, (t, h, w) in {"small": (8, 28, 36), "i3d": (20, 224, 224)}.items():
        rgb, flow = ns["get_twostream_videoclip"](video, [None, None], t, h, w, optical_flow_status="FarneBack_onTheFly",
                                                  augmentation_status="non_augmented")
        assert rgb.dtype == np.uint8 and rgb.shape == (t, h, w, 3)
        assert flow.dtype == np.float32 and flow.shape == (t, h, w, 2)
        if tag == "small":
            out["rgb_small"], out["flow_small"] = rgb, flow
        else:                                   # every 8th pixel: enough to compare values, not only the hash
            out["flow_i3d_sub"] = np.ascontiguousarray(flow[:, ::8, ::8])
        out["sha_rgb_" + tag] = np.frombuffer(hashlib.sha256(rgb.tobytes()).digest(), np.uint8)
        out["sha_flow_" + tag] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(flow).tobytes()).digest(), np.uint8)
        out["shape_" + tag] = np.array([t, h, w])
        out["absmax_" + tag] = np.array([float(np.abs(flow).max())])
    path = os.path.join(MG.GOLD, "farneback_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
