"""Times the stride-2 7x7x7 stem conv (I3D Conv3d_1a_7x7 / R3D stem, train.py:1026, 1481) alone under the lowering
variants: 2x2 vs 2x2x2 space-to-depth cells, shared-B mode on / off.  CUDA events around the op, one member."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cse_b200 import graph as G, runtime as rt           # noqa: E402
from cse_b200.model import Member                        # noqa: E402
from cse_b200.weights import synthetic_weights           # noqa: E402


def stem_graph(shape):
    g = G.Graph("stem", "functional")
    x = g.input(shape, name="in")
    x = g.conv3d(x, 64, (7, 7, 7), (2, 2, 2), "same", False, None, name="c")
    x = g.bn(x, scale=False, name="b")
    g.relu(x, name="r")
    return g


def time_op(m, x, idx, iters=5):
    m.run_ops([x], 0, m.num_ops)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.run_ops([x], idx, idx + 1)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    out = []
    for c, shape, n in ((3, (64, 224, 224, 3), 32), (2, (64, 224, 224, 2), 32), (3, (16, 112, 112, 3), 256)):
        g = stem_graph(shape)
        w = synthetic_weights(g, seed=1)
        x = torch.randint(0, 256, (n,) + shape, dtype=torch.uint8, device="cuda")
        flops = 2.0 * n * np.prod([s // 2 for s in shape[:3]]) * 64 * 343 * c
        for depth, bshare, pair in ((False, -1, 0), (False, 0, -1), ("always", -1, 0), ("always", 0, -1)):
            if True:
                rt.tune("bshare_min_tiles", bshare)
                rt.tune("pair_min_tiles", pair)
                m = Member(g, w, precision="bf16", max_batch=n, s2d_depth=depth)
                op = [o for o in m.plan.ops if o.name == "c"][0]
                idx = m.plan.ops.index(op)
                ms = time_op(m, x, idx)
                pre = time_op(m, x, 0)
                rec = {"C": c, "shape": shape, "n": n, "s2d_depth": depth, "bshare_min_tiles": bshare, "pair_min_tiles": pair, "k": op.k, "kc": op.kc,
                       "Cin": op.in0.C, "brick": op.brick, "halo": op.halo, "stem_ms": round(ms, 3),
                       "alg_tflops": round(flops / ms / 1e9, 1), "preprocess_ms": round(pre, 3)}
                print(json.dumps(rec), flush=True)
                out.append(rec)
                del m
    rt.tune("bshare_min_tiles", -1)
    rt.tune("pair_min_tiles", -1)


if __name__ == "__main__":
    main()
