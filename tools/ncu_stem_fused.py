"""Small driver for ncu: the member-fused 7x7x7 stride-2 stem (two members, N = 128, CTA-pair kernel) of I3D-64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cse_b200 import graph as G                                 # noqa: E402
from cse_b200.ensemble_runtime import DeviceEnsemble            # noqa: E402
from cse_b200.weights import synthetic_weights                  # noqa: E402


def main():
    shape, n = (64, 224, 224, 3), 32
    g = G.build_model_graph("I3D", shape, 11)
    ens = DeviceEnsemble(g, [synthetic_weights(g, seed=1 + j) for j in range(2)], max_batch=n, micro_batch=n)
    assert ens.roles == ["lead", "follow"]
    x = [torch.randint(0, 256, (n,) + shape, dtype=torch.uint8, device="cuda")]
    lead = ens.members[0]
    stem = [k for k, o in enumerate(lead.plan.ops) if o.name.endswith("+peer")][0]
    lead.run_ops(x, 0, stem)
    for _ in range(3):
        lead.run_ops(x, stem, stem + 1)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
