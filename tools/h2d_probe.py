"""Pinned-host -> device copy bandwidth of this box, alone and under a running ensemble step (explains the gap
between the resident and the end-to-end clips/s of the short-step workloads)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    torch.cuda.set_device(0)
    n = 154 * 1024 * 1024
    host = torch.randint(0, 256, (n,), dtype=torch.uint8).pin_memory()
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    copy = torch.cuda.Stream()
    out = {}
    for name, chunks in (("one_copy", 1), ("8_chunks", 8)):
        step = n // chunks
        for _ in range(3):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy):
            a.record(copy)
            for _ in range(10):
                for c in range(chunks):
                    dev[c * step:(c + 1) * step].copy_(host[c * step:(c + 1) * step], non_blocking=True)
            b.record(copy)
        torch.cuda.synchronize()
        out[name + "_gbs"] = 10 * n / (a.elapsed_time(b) / 1e3) / 1e9
    # pageable host memory for comparison
    pag = torch.randint(0, 256, (n,), dtype=torch.uint8)
    t0 = time.perf_counter()
    for _ in range(3):
        dev.copy_(pag)
    torch.cuda.synchronize()
    out["pageable_gbs"] = 3 * n / (time.perf_counter() - t0) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
