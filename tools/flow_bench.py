#!/usr/bin/env python
"""Throughput of the on-the-fly Farneback branch: cse_farneback on the GPU against cv2 on the host cores,
on one synthetic video of F frames at the extractor's working size (longest side 224).  Prints one JSON line.
Usage: python tools/flow_bench.py [--frames 65] [--height 126] [--width 224] [--reps 5]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=65)
    ap.add_argument("--height", type=int, default=224)
    ap.add_argument("--width", type=int, default=224)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-pairs", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import cv2
    import torch
    from cse_b200 import runtime as rt
    rng = np.random.default_rng(0)
    base = cv2.GaussianBlur(rng.integers(0, 256, (a.height + 2 * a.frames + 8, a.width + 2 * a.frames + 8)).astype(np.float32), (9, 9), 2.5)
    gray = np.stack([np.clip(base[i:i + a.height, 2 * i:2 * i + a.width], 0, 255).astype(np.uint8) for i in range(a.frames)])
    dev = torch.from_numpy(gray).to("cuda:0")
    flow = rt.farneback(dev)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(a.reps):
        flow = rt.farneback(dev)
    ev[1].record()
    torch.cuda.synchronize()
    gpu_ms = ev[0].elapsed_time(ev[1]) / a.reps
    if a.no_cpu:
        print(json.dumps({"gpu_ms_per_video": round(gpu_ms, 3)}))
        return
    n = min(a.cpu_pairs, a.frames - 1)
    cv2.setNumThreads(1)                       # one loader worker = one thread (the reference's `workers` are processes)
    t0 = time.perf_counter()
    ref = [cv2.calcOpticalFlowFarneback(gray[i], gray[i + 1], None, 0.5, 5, 11, 5, 5, 1.1, 0) for i in range(n)]
    cpu_ms = (time.perf_counter() - t0) * 1e3 / n
    cv2.setNumThreads(-1)                      # and with OpenCV's own thread pool over all host cores
    t0 = time.perf_counter()
    for i in range(n):
        cv2.calcOpticalFlowFarneback(gray[i], gray[i + 1], None, 0.5, 5, 11, 5, 5, 1.1, 0)
    cpu_mt_ms = (time.perf_counter() - t0) * 1e3 / n
    err = float(np.abs(flow[:n].cpu().numpy() - np.stack(ref)).max())
    px = a.height * a.width
    print(json.dumps({"frames": a.frames, "size": [a.height, a.width], "gpu_ms_per_video": round(gpu_ms, 3),
                      "gpu_pairs_per_s": round((a.frames - 1) / gpu_ms * 1e3, 1), "gpu_mpix_per_s": round((a.frames - 1) * px / gpu_ms / 1e3, 1),
                      "cv2_ms_per_pair": round(cpu_ms, 3), "cv2_pairs_per_s": round(1e3 / cpu_ms, 1), "cv2_all_threads_ms_per_pair": round(cpu_mt_ms, 3), "cv2_threads": cv2.getNumThreads(), "host_cores": os.cpu_count(),
                      "speedup": round(cpu_ms * (a.frames - 1) / gpu_ms, 1), "max_abs_diff_vs_cv2": err,
                      "max_abs_flow": float(np.abs(np.stack(ref)).max())}))


if __name__ == "__main__":
    main()
