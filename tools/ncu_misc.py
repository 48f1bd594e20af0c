"""Small driver for ncu (--profile-from-start off): the horizontally fused sibling 1x1x1 GEMM of I3D Mixed_3b and one
3x3x3 conv of R3D-34 stage 1 (CTA-pair h-halo mode), each launched twice inside the profiler range."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cse_b200 import graph as G                                 # noqa: E402
from cse_b200.model import Member                               # noqa: E402
from cse_b200.weights import synthetic_weights                  # noqa: E402


def main():
    cases = [("I3D", (64, 224, 224, 3), 16, lambda o: "+" in o.name and "Conv3d_3b_0a" in o.name),
             ("R3D_34", (16, 112, 112, 3), 128, lambda o: o.name == "conv3d_2")]
    if len(sys.argv) > 1 and sys.argv[1] == "small":        # the small-Cin 3x3x3 convs of Inception branch 2b
        cases = [("I3D", (64, 224, 224, 3), 16, lambda o: o.name.startswith("Conv3d_3c_2b_3x3")),
                 ("I3D", (64, 224, 224, 3), 16, lambda o: o.name.startswith("Conv3d_3b_2b_3x3"))]
    if len(sys.argv) > 1 and sys.argv[1] == "pool":         # the 3x3x3 / stride-1 'same' pools of the Inception branches
        cases = [("I3D", (64, 224, 224, 3), 32, lambda o: o.name.startswith("MaxPool2d_3c_3a")),
                 ("I3D", (64, 224, 224, 3), 32, lambda o: o.name.startswith("MaxPool2d_4c_3a"))]
    for mt, shape, n, pick in cases:
        g = G.build_model_graph(mt, shape, 11)
        m = Member(g, synthetic_weights(g, seed=1), precision="bf16", max_batch=n)
        x = [torch.randint(0, 256, (n,) + shape, dtype=torch.uint8, device="cuda")]
        m.forward_device(x)
        k = [i for i, o in enumerate(m.plan.ops) if pick(o)][0]
        op = m.plan.ops[k]
        print(mt, k, op.name, "bn", op.bn, "kc", op.kc, "halo", op.halo, "split", op.out_split, op.out_split2, "flops/clip %.3g" % op.flops)
        m.run_ops(x, k, k + 1)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(2):
            m.run_ops(x, k, k + 1)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        del m


if __name__ == "__main__":
    main()
