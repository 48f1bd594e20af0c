#!/usr/bin/env python
"""End-to-end throughput of the drop-in path INCLUDING video decoding (outside bench.py's contract):
N synthetic MJPG videos on disk -> ClipSequence (cv2 decode, select_frames, resize on the CPU or on the
GPU) -> 4-member C3D DeviceEnsemble -> probabilities, for several `workers` settings.  Shows where the
time goes once the network itself runs at thousands of clips/s (decode is next).

--flow: the FarneBack_onTheFly TwoStream-I3D variant instead (T = 20, 224 x 224, 4 members): the dense flow of every
video computed by the loader threads with OpenCV like the reference (CSE_CPU_FLOW=1) or on the GPU (cse_farneback).

Usage: python tools/bench_ingest.py [--clips 96] [--size 320x240] [--frames 48] [--flow]"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=96)
    ap.add_argument("--size", default="320x240")
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--flow", action="store_true")
    args = ap.parse_args()
    import cv2
    import pandas as pd
    import torch
    from cse_b200 import clips, ensemble as E, graph as G
    from cse_b200.ensemble_runtime import DeviceEnsemble
    from cse_b200.weights import synthetic_weights

    w, h = (int(v) for v in args.size.split("x"))
    rng = np.random.default_rng(0)
    tmp = tempfile.mkdtemp(prefix="cse_ingest_")
    paths = []
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for k in range(args.clips):
        p = os.path.join(tmp, "clip%03d.avi" % k)
        vw = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"MJPG"), 25, (w, h))
        for i in range(args.frames):
            f = np.stack([127 + 90 * np.sin((xx + 3 * i + 11 * k) / (9.0 + c)) * np.cos((yy - 2 * i) / (7.0 + c)) for c in range(3)], -1)
            vw.write(np.clip(f + rng.normal(0, 3, f.shape), 0, 255).astype(np.uint8))
        vw.release()
        paths.append(p)
    data = pd.DataFrame({"rgbclips_path": paths, "class": [k % 11 for k in range(args.clips)]})
    if args.flow:
        return flow_mode(args, paths, w, h)
    shape = (16, 112, 112, 3)
    g = G.build_model_graph("C3D", shape, 11)
    ens = DeviceEnsemble(g, [synthetic_weights(g, seed=100 + j) for j in range(4)], max_batch=32, micro_batch=32)
    dev = torch.device("cuda", torch.cuda.current_device())
    out = {}
    ref = None
    for resize, device in (("cpu", None), ("gpu", dev)):
        for workers in (1, 4, 16):
            seq = clips.ClipSequence(data, "C3D", shape, 11, batch_size=1, device=device)
            E._predict_members(ens, seq, min(8, seq.n), (None, 0, 1), 32, workers=workers)        # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            probs = E._predict_members(ens, seq, seq.n, (None, 0, 1), 32, workers=workers)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if ref is None:
                ref = probs
            assert np.array_equal(probs, ref), "probabilities differ between ingest settings"
            out["resize=%s workers=%d" % (resize, workers)] = args.clips / dt
    t0 = time.perf_counter()
    for p in paths[:16]:
        clips.decode_frames(p)
    dec = 16 / (time.perf_counter() - t0)
    print("videos: %d x %d frames of %dx%d MJPG; host cores: %d" % (args.clips, args.frames, w, h, os.cpu_count()))
    print("decode only, one thread: %.1f clips/s" % dec)
    for k, v in out.items():
        print("%-28s %8.1f clips/s" % (k, v))


def flow_mode(args, paths, w, h):
    import pandas as pd
    import torch
    from cse_b200 import clips, ensemble as E, graph as G
    from cse_b200.ensemble_runtime import DeviceEnsemble
    from cse_b200.weights import synthetic_weights
    data = pd.DataFrame({"rgbclips_path": paths, "x_axis_flowclips_path": [""] * len(paths), "y_axis_flowclips_path": [""] * len(paths),
                         "class": [k % 11 for k in range(len(paths))]})
    shape = (20, 224, 224, 0)
    g = G.build_model_graph("TWOSTREAM_I3D", shape, 11)
    ens = DeviceEnsemble(g, [synthetic_weights(g, seed=100 + j) for j in range(4)], max_batch=16, micro_batch=16, input_dtypes=("u8", "f32"))
    dev = torch.device("cuda", torch.cuda.current_device())
    out, probs = {}, {}
    for flow, workers in (("cpu", 4), ("cpu", 16), ("gpu", 4), ("gpu", 16)):
        os.environ["CSE_CPU_FLOW"] = "1" if flow == "cpu" else "0"
        seq = clips.ClipSequence(data, "TWOSTREAM_I3D", shape, 11, batch_size=1, optical_flow_status="FarneBack_onTheFly", device=dev)
        E._predict_members(ens, seq, min(4, seq.n), (None, 0, 1), 16, workers=workers)           # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        probs[flow] = E._predict_members(ens, seq, seq.n, (None, 0, 1), 16, workers=workers)
        torch.cuda.synchronize()
        out["flow=%s workers=%d" % (flow, workers)] = len(paths) / (time.perf_counter() - t0)
    agree = float((probs["cpu"].sum(0).argmax(-1) == probs["gpu"].sum(0).argmax(-1)).mean())
    print("videos: %d x %d frames of %dx%d MJPG -> TwoStream-I3D 20x224x224, 4 members, FarneBack_onTheFly; host cores: %d"
          % (len(paths), args.frames, w, h, os.cpu_count()))
    for k, v in out.items():
        print("%-28s %8.1f clips/s" % (k, v))
    print("max |p_gpu_flow - p_cpu_flow| = %.3g, soft-vote agreement %.4f" % (float(np.abs(probs["cpu"] - probs["gpu"]).max()), agree))


if __name__ == "__main__":
    main()
