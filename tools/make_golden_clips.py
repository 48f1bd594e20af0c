#!/usr/bin/env python
"""Generate tests/golden/clips_golden.npz (+ the tiny videos it reads) by running the REFERENCE's own
clip assembly: ``select_frames``, ``get_onestream_videoclip``, ``get_twostream_videoclip`` and
``opticalflow_TVL1_retriever`` are extracted - unmodified - with ``ast`` from /root/reference/train.py
and executed with the container's OpenCV.  Only ``cv2.waitKey`` is stubbed (the headless OpenCV build
raises "not implemented"; the reference only polls it for a 'q' key press).

Run in the build container (needs /root/reference).  Usage: python tools/make_golden_clips.py
"""
import ast
import hashlib
import os

import cv2
import numpy as np

REF = "/root/reference/train.py"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "..", "tests", "golden")
WANTED = ["select_frames", "get_onestream_videoclip", "get_twostream_videoclip", "opticalflow_TVL1_retriever"]


class Cv2Proxy:
    """cv2 with waitKey stubbed; everything else is the real module."""

    def __getattr__(self, name):
        if name == "waitKey":
            return lambda *_: -1
        return getattr(cv2, name)


def load_reference_functions():
    tree = ast.parse(open(REF).read())
    ns = {"np": np, "cv2": Cv2Proxy()}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    missing = [w for w in WANTED if w not in ns]
    if missing:
        raise RuntimeError("reference functions not found: %s" % missing)
    return ns


def write_video(path, frames, color=True):
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (w, h), color)
    if not vw.isOpened():
        raise RuntimeError("cannot open a MJPG writer for %s" % path)
    for f in frames:
        vw.write(f)
    vw.release()


def synthetic_frames(rng, n, h, w, c):
    """Smooth moving blobs + texture so that the JPEG-coded frames are not flat."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = []
    for i in range(n):
        f = np.zeros((h, w, c), np.float32)
        for k in range(c):
            f[..., k] = 127 + 90 * np.sin((xx + 3 * i * (k + 1)) / (7.0 + k)) * np.cos((yy - 2 * i) / (5.0 + 2 * k))
        f += rng.normal(0, 4, f.shape)
        out.append(np.clip(f, 0, 255).astype(np.uint8))
    return out


def main():
    ns = load_reference_functions()
    rng = np.random.default_rng(20261018)
    os.makedirs(GOLD, exist_ok=True)
    out = {}
    # select_frames on plain lists: (n, T) -> kept indices
    sel = []
    for n, t in [(16, 16), (17, 16), (31, 16), (32, 16), (33, 16), (100, 16), (100, 20), (64, 64), (250, 64), (7, 16), (1, 16),
                 (40, 20), (39, 20)]:
        sel.append([n, t] + list(ns["select_frames"](list(range(n)), t)))
    out["select_cases"] = np.array([len(s) for s in sel])
    out["select_flat"] = np.concatenate([np.asarray(s) for s in sel])
    # videos: 37 frames of 90x122 (odd sizes) colour, 2 gray flow videos
    rgb = synthetic_frames(rng, 37, 90, 122, 3)
    fx = [f[..., 0] for f in synthetic_frames(rng, 37, 90, 122, 1)]
    fy = [f[..., 0] for f in synthetic_frames(rng, 37, 90, 122, 1)]
    paths = {k: os.path.join(GOLD, "clip_%s.avi" % k) for k in ("rgb", "flow_x", "flow_y")}
    write_video(paths["rgb"], rgb)
    write_video(paths["flow_x"], [cv2.cvtColor(f, cv2.COLOR_GRAY2BGR) for f in fx])
    write_video(paths["flow_y"], [cv2.cvtColor(f, cv2.COLOR_GRAY2BGR) for f in fy])
    # small targets are stored in full, model-sized ones as sha256 of the bytes
    for tag, (t, h, w) in {"small": (8, 28, 36), "up": (3, 100, 130), "c3d": (16, 112, 112), "i3d": (20, 224, 224)}.items():
        one = ns["get_onestream_videoclip"](paths["rgb"], t, h, w)
        r2, f2 = ns["get_twostream_videoclip"](paths["rgb"], [paths["flow_x"], paths["flow_y"]], t, h, w,
                                               optical_flow_status="TVL1_precomputed")
        assert one.dtype == np.uint8 and one.shape == (t, h, w, 3) and f2.shape == (t, h, w, 2)
        assert np.array_equal(one, r2)
        if tag in ("small", "up"):
            out["rgb_" + tag], out["flow_" + tag] = one, f2
        out["sha_rgb_" + tag] = np.frombuffer(hashlib.sha256(one.tobytes()).digest(), np.uint8)
        out["sha_flow_" + tag] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(f2).tobytes()).digest(), np.uint8)
        out["shape_" + tag] = np.array([t, h, w])
    np.savez_compressed(os.path.join(GOLD, "clips_golden.npz"), **out)
    print("wrote", os.path.join(GOLD, "clips_golden.npz"), {k: os.path.getsize(p) for k, p in paths.items()})


if __name__ == "__main__":
    main()
