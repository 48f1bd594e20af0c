#!/usr/bin/env python
# coding=utf-8
"""Drop-in for the reference's ``python evaluate_ensemble.py -op ... `` on the ensemble-inference
hot path (same flags; evaluate_ensemble.py:1676-1794 of the reference), running on libcse_b200.

Single GPU:   python evaluate_ensemble.py -op Evaluate_ensembles -mt C3D -tc _SCRATCH -wt SUM ...
Multi GPU:    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
                  evaluate_ensemble.py -op Store_models_probabilities ...
              (clips are sharded over the ranks, probabilities all-gathered over NCCL)
"""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _init_distributed():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    return dist


if __name__ == '__main__':
    from cse_b200 import runtime as rt
    from cse_b200.ensemble import build_parser, main

    args = build_parser().parse_args()
    import torch
    if not torch.cuda.is_available():          # the reference exits the same way (evaluate_ensemble.py:1672-1674)
        print('error-no-gpu')
        sys.exit()
    rt.load_library()
    dist = _init_distributed()
    main(args)
    if dist is not None:
        dist.destroy_process_group()
