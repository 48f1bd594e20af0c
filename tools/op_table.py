#!/usr/bin/env python
"""Per-op table of one ensemble step (DeviceEnsemble.profile_ops): ms, algorithmic TFLOP/s, activation bytes moved
(input + output elements x 2 B, per member) and the resulting GB/s - to see which layers sit far from BOTH bounds.
Usage: python tools/op_table.py [I3D|R3D_34|C3D|TWOSTREAM_I3D] [T H W] [clips] [members]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cse_b200 import graph as G                                 # noqa: E402
from cse_b200.ensemble_runtime import DeviceEnsemble            # noqa: E402
from cse_b200.weights import synthetic_weights                  # noqa: E402


def main():
    mt = sys.argv[1] if len(sys.argv) > 1 else "I3D"
    t, h, w = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (64, 224, 224)
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 32
    members = int(sys.argv[6]) if len(sys.argv) > 6 else 4
    g = G.build_model_graph(mt, (t, h, w, 0 if mt == "TWOSTREAM_I3D" else 3), 11)
    import json
    kw = json.loads(os.environ.get("CSE_LOWER_KW", "{}"))       # lowering experiments, as in bench.py
    ens = DeviceEnsemble(g, [synthetic_weights(g, seed=1 + j) for j in range(members)], max_batch=n, micro_batch=n, **kw)
    torch.manual_seed(0)
    x = [torch.randint(0, 256, (n,) + tuple(g.shape(i)), dtype=torch.uint8, device="cuda") for i in g.inputs]
    ens.forward_members(x)
    torch.cuda.synchronize()
    print("logits checksum %.9g  max |logit| %.6g" % (float(ens.logits[:, :n].double().sum()), float(ens.logits[:, :n].abs().max())))
    rows = ens.profile_ops(x, iters=2)
    ops = {}
    for m in ens.members:
        for o in m.plan.ops:
            ops.setdefault(o.name[:-5] if o.name.endswith("+peer") else o.name, o)
    tot = sum(r["ms"] for r in rows)
    print("%-58s %-9s %8s %6s %8s %8s" % ("op", "engine", "ms", "share", "TFLOP/s", "GB/s"))
    for r in sorted(rows, key=lambda r: -r["ms"]):
        o = ops[r["name"]]
        act = 0.0
        for ref in (o.in0, o.in1, o.out0, o.out1, o.out2):
            if ref is not None:
                act += float(np.prod(ref.dims)) * ref.C * 2
        act *= n * members
        print("%-58s %-9s %8.3f %5.1f%% %8.1f %8.0f" % (r["name"][:58], r["engine"] or r["kind"], r["ms"], 100 * r["ms"] / tot,
                                                       r["flops"] / r["ms"] / 1e9 if r["ms"] else 0, act / r["ms"] / 1e6 if r["ms"] else 0))
    print("total %.2f ms" % tot)


if __name__ == "__main__":
    main()
