"""Small driver for ncu: the I3D / R3D 7x7x7 stride-2 stem conv (CTA-pair kernel) a few times, one member."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cse_b200.model import Member                        # noqa: E402
from cse_b200.weights import synthetic_weights           # noqa: E402
from tools.stem_bench import stem_graph                  # noqa: E402


def main():
    for shape, n in (((64, 224, 224, 3), 32), ((64, 224, 224, 2), 32), ((16, 112, 112, 3), 256)):
        g = stem_graph(shape)
        m = Member(g, synthetic_weights(g, seed=1), precision="bf16", max_batch=n)
        x = torch.randint(0, 256, (n,) + shape, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            m.run_ops([x], 0, m.num_ops)
        torch.cuda.synchronize()
        del m


if __name__ == "__main__":
    main()
