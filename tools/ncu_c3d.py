"""Small driver for ncu: one C3D member (train.py:1224-1273) on 256 clips, two forward passes (profile the second)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cse_b200 import graph as G                          # noqa: E402
from cse_b200.model import Member                        # noqa: E402
from cse_b200.weights import synthetic_weights           # noqa: E402


def main():
    shape = (16, 112, 112, 3)
    g = G.build_model_graph("C3D", shape, 11)
    m = Member(g, synthetic_weights(g, seed=100), precision="bf16", max_batch=256)
    x = torch.randint(0, 256, (256,) + shape, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        m.forward_device([x])
    torch.cuda.synchronize()
    print([(o.name, o.ksplit) for o in m.plan.ops if o.engine == 2])


if __name__ == "__main__":
    main()
