#!/usr/bin/env python
"""bench.py - ensemble clips/sec of the B200-native hot path (and of the CPU reference arm).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic clips per GPU: every member
of the ensemble runs its forward pass on the batch (pre-processing included) and the soft vote
produces one prediction per clip.  Default workload = BASELINE.json configs[1]: the homogeneous
C3D fold ensemble on [256,16,112,112,3] uint8 clips, 11 classes, M = folds-1 = 4 members
(evaluate_ensemble.py:1042-1049), SUM vote.  Prints ONE JSON line (rank 0).

* value      : ensemble clips/s over all GPUs, inputs resident in HBM (uint8), CUDA events,
               max over ranks.
* e2e        : same metric through Member/Ensemble's host API: pinned host uint8 batch -> H2D,
               forward, vote, D2H of the int32 predictions, all inside the timed region.
* roofline   : the tcgen05 conv kernel: algorithmic conv FLOPs (un-padded reference shapes) of
               the layers it ran / its summed launch durations (CUDA events around each launch in
               a separate profiling pass on the same stream), vs the measured dense-bf16 peak.
* cpu_baseline / --impl reference : the oracle's torch-CPU fp32 restatement of the same members
               (Keras 2.2.4 / TF 1.15 cannot be installed offline) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

# name: (groups, clips per GPU per step); group = (model_type, clip shape, members, micro-batch)
# members = folds - 1 = 4 per architecture, as the reference builds them for `-fn 5`
# (evaluate_ensemble.py:1042-1049).
WORKLOADS = {
    "c3d_ens": ([("C3D", (16, 112, 112, 3), 4, 256)], 256),                       # BASELINE configs[1]
    "c3d_single_b8": ([("C3D", (16, 112, 112, 3), 1, 8)], 8),                     # configs[0]
    "r3d34_ens": ([("R3D_34", (16, 112, 112, 3), 4, 256)], 256),
    "i3d20_ens": ([("I3D", (20, 224, 224, 3), 4, 128)], 128),                     # reference-true T=20 (train.py:1573)
    "i3d64_ens": ([("I3D", (64, 224, 224, 3), 4, 64)], 64),                       # configs[2]
    "twostream20_ens": ([("TWOSTREAM_I3D", (20, 224, 224, 0), 4, 64)], 64),
    "twostream64_ens": ([("TWOSTREAM_I3D", (64, 224, 224, 0), 4, 32)], 32),       # configs[3]
    "global_hetero": ([("C3D", (16, 112, 112, 3), 4, 256), ("I3D", (64, 224, 224, 3), 4, 32),
                       ("R3D_34", (16, 112, 112, 3), 4, 256)], 256),              # configs[4] (per GPU)
    "global_hetero_t20": ([("C3D", (16, 112, 112, 3), 4, 256), ("I3D", (20, 224, 224, 3), 4, 128),
                           ("R3D_34", (16, 112, 112, 3), 4, 256)], 256),
}


def workload_config(name, groups, batch, world, by_members=False):
    par = ("members sharded over %d GPU(s) by FLOPs, every GPU runs all %d clips, 1 all-gather of the fp32 member "
           "probabilities per step, vote in member order" % (world, batch)) if by_members else (
           "clips sharded over %d GPU(s), members replicated, 1 all-gather of int32 predictions per step" % world)
    return {"workload": name, "models": [g[0] for g in groups], "clips": [list(g[1]) for g in groups],
            "members": [g[2] for g in groups], "batch_per_gpu": batch, "vote": "SUM", "classes": 11,
            "parallelism": par}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU during the timed region: NVML in-thread
    (no fork in the timed region), `nvidia-smi` as the fallback."""

    REASONS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4)]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            self.rows.append((sm, self.max_sm, [name for name, bit in self.REASONS if mask & bit]))
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 6 and parts[0].isdigit():
            self.rows.append((int(parts[0]), int(parts[1]) if parts[1].isdigit() else None,
                              [name for (name, _), v in zip(self.REASONS, parts[2:6]) if v.lower().startswith("active")]))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop_evt.wait(0.05 if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows if r[1]]
        reasons = sorted({name for r in self.rows for name in r[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_clips_per_s(groups, clips_per_step, steps, warmup, threads=None):
    """Oracle port (torch CPU fp32) of the same ensemble: every member forward + numpy vote."""
    import torch
    from cse_b200 import graph as G
    from cse_b200.weights import synthetic_weights
    from oracle import models as OM, vote as OV
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    rng = np.random.default_rng(1234)
    prepared = []
    for gi, (mt, shape, members, _) in enumerate(groups):
        g = G.build_model_graph(mt, shape, 11)
        ws = [synthetic_weights(g, seed=100 + 10 * gi + m) for m in range(members)]
        xs = [rng.integers(0, 256, (clips_per_step,) + tuple(g.shape(n)), dtype=np.uint8) for n in g.inputs]
        prepared.append((mt, ws, xs if len(xs) > 1 else xs[0]))

    def step():
        probs = [OM.forward(mt, w, x, torch.float32)[1].numpy() for mt, ws, x in prepared for w in ws]
        return OV.ensemble_predictions(np.stack(probs).astype(np.float64), np.ones(len(probs)))

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return clips_per_step * steps / dt, dt / steps, threads


def cpu_sample_size(groups, budget_gflop=3000.0):
    """Clips per CPU step: bounded so a step is a few seconds of host work."""
    gflop = sum(m * {"C3D": 77.1, "R3D_34": 13.3}.get(mt, 222.3 * shape[0] / 64 * (2 if mt == "TWOSTREAM_I3D" else 1))
                for mt, shape, m, _ in groups)
    return max(1, min(32, int(budget_gflop / gflop)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    groups, batch = WORKLOADS[args.workload]
    sample = cpu_sample_size(groups)
    v, sec, threads = cpu_reference_clips_per_s(groups, sample, args.steps, args.warmup)
    members = sum(g[2] for g in groups)
    line = {
        "impl": "reference", "metric": "ensemble clips/sec", "value": v, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, groups, batch, args.gpus),
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": "%d clips x %d members per step, torch CPU fp32 oracle restatement "
                                   "(Keras 2.2.4/TF 1.15 not installable offline)" % (sample, members)},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cse_b200 import graph as G, runtime as rt
    from cse_b200.ensemble_runtime import HeteroEnsemble
    from cse_b200.weights import synthetic_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt.load_library()
    groups, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    if args.members:
        groups = [(mt, shape, args.members, mb) for mt, shape, _, mb in groups]
    if args.micro_batch:
        groups = [(mt, shape, m, args.micro_batch) for mt, shape, m, _ in groups]
    groups = [(mt, shape, m, min(mb, batch)) for mt, shape, m, mb in groups]
    lower_kw = json.loads(os.environ.get("CSE_LOWER_KW", "{}"))      # lowering experiments (e.g. {"fuse_pool": false})
    by_members = args.shard == "members" and world > 1
    all_graphs = [G.build_model_graph(mt, shape, 11) for mt, shape, _, _ in groups]
    members = sum(m for _, _, m, _ in groups)
    owned_all = None
    mine = None
    if by_members:
        # member-sharded partition: every rank sees all clips and runs only the members it
        # owns (balanced by FLOPs per clip); probabilities are all-gathered into member order
        from cse_b200.ensemble import shard_members, MemberGather
        if world > members:
            raise SystemExit("--shard members needs at least one member per rank (%d members, %d ranks)" % (members, world))
        costs = [g.total_flops() for g, (_, _, m, _) in zip(all_graphs, groups) for _ in range(m)]
        owned_all = shard_members(costs, world)
        mine = set(owned_all[rank])
    built, graphs = [], []
    flat = 0
    for gi, ((mt, shape, m, mb), g) in enumerate(zip(groups, all_graphs)):
        ws = [synthetic_weights(g, seed=100 + 10 * gi + j) for j in range(m) if mine is None or flat + j in mine]
        flat += m
        if ws:
            graphs.append(g)
            built.append((g, ws, mb))
    ens = HeteroEnsemble(built, precision=args.precision, max_batch=batch, **lower_kw)
    del built
    if by_members:
        ens.gather = MemberGather(owned_all, members, dist, world)
    gen = torch.Generator(device="cpu").manual_seed(1234 + (0 if by_members else rank))
    host = [[torch.randint(0, 256, (batch,) + tuple(g.shape(n)), dtype=torch.uint8, generator=gen).pin_memory()
             for n in g.inputs] for g in graphs]
    dev_in = [[h.cuda() for h in hs] for hs in host]
    in_bytes = sum(h.numel() for hs in host for h in hs)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: clips are sharded over the ranks (members replicated); the one exchange of the path is
    # the all-gather of the per-clip predictions (evaluate_ensemble.py:1262-1268 writes them all)
    gathered = torch.empty((world * batch,), dtype=torch.int32, device="cuda") if world > 1 and not by_members else None

    def step_resident():
        pred = ens.predict_device(dev_in)
        if by_members:
            return pred                  # every rank voted on the gathered [M, batch, C] block
        if world > 1:
            dist.all_gather_into_tensor(gathered, pred)
            return gathered
        return pred

    pinned_pred = torch.empty((world * batch,), dtype=torch.int32).pin_memory()

    def run_e2e(nsteps):
        """Host API: every step uploads its pinned uint8 batch (H2D on a copy stream, overlapped with
        the previous step's compute), runs members + vote (+ all-gather), and reads the predictions
        back (D2H); all of it inside the timed region."""
        out = None
        for pred in ens.stream_host(host for _ in range(nsteps)):
            if by_members:
                pinned_pred[:batch].copy_(pred, non_blocking=True)
            elif world > 1:
                dist.all_gather_into_tensor(gathered, pred)
                pinned_pred.copy_(gathered, non_blocking=True)
            else:
                pinned_pred[:batch].copy_(pred, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pinned_pred

    # ---- kernel-resident timing ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        pred = step_resident()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ens.last_launches * args.steps
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end timing (host buffers, H2D + D2H inside) ----
    run_e2e(max(1, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    run_e2e(args.steps)
    e3.record(stream)
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    total_clips = batch * (1 if by_members else world) * args.steps
    value = total_clips / (ms / 1e3)
    e2e_value = total_clips / (ms_e2e / 1e3)

    if rank == 0:
        peaks = load_peaks()
        prof = ens.profile_ops(dev_in, iters=2)
        tc = [p for p in prof if p["engine"] == "tcgen05"]
        tc_flops = sum(p["flops"] for p in tc)
        tc_ms = sum(p["ms"] for p in tc)
        step_ms_prof = sum(p["ms"] for p in prof)
        achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
        peak = peaks["bf16_sustained"]
        roofline = {"bound": "tensor", "kernel": "conv_tc_kernel", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": "%s bf16_tflops_sustained (burst %.1f)" % (peaks["source"], peaks["bf16_burst"]),
                    "kernel_share_of_step": tc_ms / step_ms_prof if step_ms_prof else None,
                    "launches_per_step": len(tc),
                    "whole_step_model_tflops": sum(m * g.total_flops() for g, (_, _, m, _) in zip(all_graphs, groups)) * batch
                                               / (ms / args.steps / 1e3) / 1e12 / (world if by_members else 1)}
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath)).get(args.workload)
            if tr and [tr.get("micro_batch")] == ens.micro_batch and tr.get("launches_captured") == len(tc):
                roofline["traffic"] = tr["dram_bytes_per_launch_avg"]
                roofline["traffic_unit"] = "B per conv_tc launch (ncu dram read+write, avg over the member's launches)"
                roofline["algorithmic_flops_per_launch_avg"] = tc_flops / (len(tc) * ens.M * (batch // ens.micro_batch[0]))
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"workload": args.workload, "batch": batch, "members": members, "micro_batch": ens.micro_batch,
                           "ops": prof}, f, indent=1)
        # HBM-bound kernels of the path (pre-processing, pooling): algorithmic bytes / CUDA-event time
        hbm = {}
        for p_ in prof:
            if p_.get("bytes", 0) > 0 and p_["ms"] > 0:
                k_ = hbm.setdefault(p_["kind"], {"bytes": 0.0, "ms": 0.0, "launch_groups": 0})
                k_["bytes"] += p_["bytes"]; k_["ms"] += p_["ms"]; k_["launch_groups"] += 1
        roofline["hbm_kernels"] = [{"kind": k_, "ops": v_["launch_groups"], "ms": round(v_["ms"], 3),
                                    "achieved_gbs": round(v_["bytes"] / (v_["ms"] / 1e3) / 1e9, 1),
                                    "peak_gbs": peaks["hbm_gbs"],
                                    "frac": round(v_["bytes"] / (v_["ms"] / 1e3) / 1e9 / peaks["hbm_gbs"], 3)}
                                   for k_, v_ in sorted(hbm.items())]
        top = sorted(prof, key=lambda p: -p["ms"])[:6]
        roofline["top_ops"] = [{"op": p["name"], "engine": p["engine"], "ms": round(p["ms"], 3),
                                "tflops": round(p["flops"] / (p["ms"] / 1e3) / 1e12, 1) if p["ms"] > 0 else 0}
                               for p in top]
        cpu = None
        if not args.no_cpu_baseline:
            sample = cpu_sample_size(groups, 6000.0)
            v, sec, threads = cpu_reference_clips_per_s(groups, sample, 2, 1)
            cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                   "sample": "%d clips x %d members per pass, 1 warm-up + 2 timed passes, torch CPU fp32 oracle "
                             "restatement (Keras 2.2.4/TF 1.15 not installable offline)" % (sample, members)}
        line = {
            "metric": "ensemble clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if by_members else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": dict(workload_config(args.workload, groups, batch, world, by_members),
                           l2_policy="input batch (%d MB uint8) and activations exceed the 126 MB L2" % (in_bytes >> 20)),
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": batch * 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else any library prints
    while the bench runs (NCCL's version banner, torchrun notices) is routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3d_ens", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--members", type=int, default=0)
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--shard", default="clips", choices=["clips", "members"],
                    help="N > 1: shard the clips (default, weak scaling) or the ensemble members (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
