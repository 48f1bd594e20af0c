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


def workload_config(name, groups, batch, world, shard="clips"):
    """The `config` object of the JSON line - identical in both arms (ours and --impl reference)."""
    par = {"members": "members sharded over %d GPU(s) by FLOPs, every GPU runs all %d clips, 1 all-gather of the fp32 member "
                      "probabilities per step, vote in member order" % (world, batch),
           "units": "ONE batch of %d clips: (member, clip-chunk) units spread over %d GPU(s) by FLOPs, 1 all-reduce of the "
                    "disjoint fp32 probability blocks per step, vote in member order" % (batch, world),
           "clips": "clips sharded over %d GPU(s), members replicated, 1 all-gather of int32 predictions per step" % world}[shard]
    chans = lambda mt, shape: 5 if mt == "TWOSTREAM_I3D" else shape[3]
    in_mb = sum(batch * shape[0] * shape[1] * shape[2] * chans(mt, shape) for mt, shape, _, _ in groups) >> 20
    l2 = ("input batch (%d MB uint8) and activations exceed the 126 MB L2" % in_mb) if in_mb > 126 else (
        "input batch %d MB uint8; every member streams its own weights and activations (> L2 per step)" % in_mb)
    return {"workload": name, "models": [g[0] for g in groups], "clips": [list(g[1]) for g in groups],
            "members": [g[2] for g in groups], "batch_per_gpu": batch, "vote": "SUM", "classes": 11,
            "parallelism": par, "l2_policy": l2,
            "cpu_sample_clips": cpu_sample_size(groups),
            "cpu_sample_note": "the CPU arm (cpu_baseline / --impl reference) runs a bounded sample of this workload: "
                               "cpu_sample_clips clips x all members per step"}


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
def measure(name, args, steps, warmup, headline):
    """Builds the ensemble of workload `name` on this rank's GPU, times `steps` resident steps and `steps` end-to-end
    steps, self-checks the predictions it timed, profiles every op launch (rank 0) and returns the result dict."""
    import torch
    import torch.distributed as dist
    from cse_b200 import graph as G, runtime as rt
    from cse_b200.ensemble_runtime import HeteroEnsemble
    from cse_b200.weights import synthetic_weights
    from oracle import vote as OV

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    groups, batch = WORKLOADS[name]
    if headline:
        if args.batch:
            batch = args.batch
        if args.members:
            groups = [(mt, shape, args.members, mb) for mt, shape, _, mb in groups]
        if args.micro_batch:
            groups = [(mt, shape, m, args.micro_batch) for mt, shape, m, _ in groups]
    shard = args.shard if headline else ("units" if name in STRONG_WORKLOADS else "clips")
    by_members = shard == "members" and world > 1
    by_units = shard == "units" and world > 1
    strong = by_members or by_units
    groups = [(mt, shape, m, min(mb, batch)) for mt, shape, m, mb in groups]
    lower_kw = json.loads(os.environ.get("CSE_LOWER_KW", "{}"))      # lowering experiments (e.g. {"fuse_pool": false})
    all_graphs = [G.build_model_graph(mt, shape, 11) for mt, shape, _, _ in groups]
    members = sum(m for _, _, m, _ in groups)
    costs = [g.total_flops() for g, (_, _, m, _) in zip(all_graphs, groups) for _ in range(m)]
    owned_all = mine = units = None
    if by_members:
        # member-sharded partition: every rank sees all clips and runs only the members it
        # owns (balanced by FLOPs per clip); probabilities are all-gathered into member order
        from cse_b200.ensemble import shard_members, MemberGather
        if world > members:
            raise SystemExit("--shard members needs at least one member per rank (%d members, %d ranks)" % (members, world))
        owned_all = shard_members(costs, world)
        mine = set(owned_all[rank])
    if by_units:
        # (member, clip-chunk) units spread over the ranks by FLOPs: strong scaling of ONE batch (latency mode)
        from cse_b200.ensemble import shard_units, UnitGather
        units = shard_units(costs, batch, world)
        longest = max(hi - lo for o in units for _, lo, hi in o)
        groups = [(mt, shape, m, min(mb, longest)) for mt, shape, m, mb in groups]
    built, graphs = [], []
    flat = 0
    for gi, ((mt, shape, m, mb), g) in enumerate(zip(groups, all_graphs)):
        ws = [synthetic_weights(g, seed=100 + 10 * gi + j) for j in range(m) if mine is None or flat + j in mine]
        flat += m
        if ws:
            graphs.append(g)
            built.append((g, ws, mb))
    ens = HeteroEnsemble(built, precision=args.precision, max_batch=batch, use_graphs=not args.no_graphs, **lower_kw)
    del built
    if by_members:
        ens.gather = MemberGather(owned_all, members, dist, world)
    if by_units:
        ens.set_units(units[rank], UnitGather(units, members, batch, 11, dist, world, ens.device))
    gen = torch.Generator(device="cpu").manual_seed(1234 + (0 if strong else rank))
    host = [[torch.randint(0, 256, (batch,) + tuple(g.shape(n)), dtype=torch.uint8, generator=gen).pin_memory()
             for n in g.inputs] for g in graphs]
    dev_in = [[h.cuda() for h in hs] for hs in host]
    in_bytes = ens.upload_bytes(host)          # unit-sharded steps upload only the clip ranges of this rank's units
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1, clips sharded over the ranks (members replicated): the one exchange of the path is the all-gather of
    # the per-clip predictions (evaluate_ensemble.py:1262-1268 writes them all)
    gathered = torch.empty((world * batch,), dtype=torch.int32, device="cuda") if world > 1 and not strong else None

    def step_resident():
        pred = ens.predict_device(dev_in, graph=not args.no_graphs)
        if strong:
            return pred                  # every rank voted on the merged [M, batch, C] block
        if world > 1:
            dist.all_gather_into_tensor(gathered, pred)
            return gathered
        return pred

    pinned_pred = torch.empty((world * batch,), dtype=torch.int32).pin_memory()

    def run_e2e(nsteps):
        """Host API: every step uploads its pinned uint8 batch (H2D on a copy stream, overlapped with the previous
        steps' compute), runs members + vote (+ exchange), and reads the predictions back (D2H); all of it inside
        the timed region."""
        for pred in ens.stream_host(host for _ in range(nsteps)):
            if strong or world == 1:
                pinned_pred[:batch].copy_(pred, non_blocking=True)
            else:
                dist.all_gather_into_tensor(gathered, pred)
                pinned_pred.copy_(gathered, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pinned_pred

    # ---- kernel-resident timing ----
    for _ in range(warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    def time_resident():
        """EXACTLY `steps` steps between two events; one more event per step only to tell a transient stall of the
        box (host descheduled, a neighbour's NCCL bootstrap: one 2-GPU run showed a single 0.7 s gap in 20 steps of
        65 ms) from the steady state: -> (last prediction, ms of the whole region, median ms of one step)."""
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record(stream)
        for i in range(steps):
            out = step_resident()
            evs[i + 1].record(stream)
        barrier()
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
        return out, evs[0].elapsed_time(evs[steps]), per[steps // 2]

    pred, ms, med = time_resident()
    remeasured = None
    stalled = torch.tensor([1.0 if ms > 1.2 * med * steps else 0.0, ms / steps, med], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stalled, op=dist.ReduceOp.MAX)          # any rank's stall counts; the record holds the maxima over ranks
    if float(stalled[0]) > 0:          # rejected and re-measured ONCE, like a throttled run; the second number stands
        remeasured = {"first_ms_per_step": float(stalled[1]), "median_step_ms": float(stalled[2]),
                      "why": "transient stall in the timed region (some rank: region > 1.2 x steps x its median step)"}
        barrier()
        pred, ms, med = time_resident()
    launches = ens.last_launches * steps
    clocks = sampler.stop() if sampler else None
    pred_resident = pred.cpu().numpy().copy()
    probs_gpu = (ens.gather.full if by_units else ens.probs)[:, :batch].cpu().numpy().astype(np.float64) \
        if not by_members else None

    # ---- end-to-end timing (host buffers, H2D + D2H inside) ----
    run_e2e(max(warmup, 7))           # every buffer set of the 3-deep pipeline is seen twice: its step graph is captured
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    pred_e2e = run_e2e(steps)
    e3.record(stream)
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)
    pred_e2e = pred_e2e.numpy().copy()
    ens_h2d = ens.h2d_copy_gbs()

    # ---- self-check of what was timed: resident == end-to-end == oracle vote on the GPU's own probabilities ----
    ncheck = len(pred_resident)
    if not np.array_equal(pred_resident, pred_e2e[:ncheck]):
        raise SystemExit("selfcheck failed (%s): resident and end-to-end predictions differ" % name)
    if probs_gpu is not None:
        mine_pred = pred_resident[rank * batch:(rank + 1) * batch] if (world > 1 and not strong) else pred_resident
        exp = OV.ensemble_predictions(probs_gpu, np.ones(probs_gpu.shape[0])).astype(np.int32)
        if not np.array_equal(mine_pred, exp):
            raise SystemExit("selfcheck failed (%s): GPU vote differs from the oracle vote on the same probabilities" % name)
        if not (np.isfinite(probs_gpu).all() and np.abs(probs_gpu.sum(-1) - 1.0).max() < 1e-4):
            raise SystemExit("selfcheck failed (%s): member probabilities are not normalised" % name)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    total_clips = batch * (1 if strong else world) * steps
    res = None
    if rank == 0:
        peaks = load_peaks()
        prof = ens.profile_ops(dev_in, iters=2) if not by_units else []
        tc = [p for p in prof if p["engine"] == "tcgen05"]
        tc_flops = sum(p["flops"] for p in tc)
        tc_ms = sum(p["ms"] for p in tc)
        step_ms_prof = sum(p["ms"] for p in prof)
        achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
        peak = peaks["bf16_sustained"]
        model_flops = sum(m * g.total_flops() for g, (_, _, m, _) in zip(all_graphs, groups)) * batch
        roofline = {"bound": "tensor", "kernel": "conv_tc_kernel", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": "%s bf16_tflops_sustained (burst %.1f)" % (peaks["source"], peaks["bf16_burst"]),
                    "kernel_share_of_step": tc_ms / step_ms_prof if step_ms_prof else None,
                    "launches_per_step": len(tc),
                    "whole_step_model_tflops": model_flops / (ms / steps / 1e3) / 1e12 / (world if strong else 1)}
        if by_units:
            roofline.update(achieved=roofline["whole_step_model_tflops"], frac=roofline["whole_step_model_tflops"] / peak,
                            note="unit-sharded step: whole-step model FLOPs / step time / GPU (no per-op profile)")
        tr = None
        for tname in ("r2_traffic.json", "r1_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if tr is None and os.path.exists(tpath):
                tr = json.load(open(tpath)).get(name)
                if tr and not ([tr.get("micro_batch")] == ens.micro_batch and tr.get("launches_captured") == len(tc)):
                    tr = None
                if tr:
                    roofline["traffic"] = tr["dram_bytes_per_launch_avg"]
                    roofline["traffic_unit"] = "B per conv_tc launch (ncu dram read+write, avg over the member's launches)"
                    roofline["traffic_source"] = "static: one ncu --set full capture of this configuration, profiles/" + tname
                    roofline["algorithmic_flops_per_launch_avg"] = tc_flops / (len(tc) * ens.M * (batch // ens.micro_batch[0]))
        if headline and args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"workload": name, "batch": batch, "members": members, "micro_batch": ens.micro_batch,
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
        h2d_gbs = ens_h2d
        res = {
            "metric": "ensemble clips/sec", "value": total_clips / (ms / 1e3), "unit": "clips/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(name, WORKLOADS[name][0] if not headline else groups, batch, world,
                                      shard if world > 1 else "clips"),
            "e2e": {"value": total_clips / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": batch * 4, "ms_per_step": ms_e2e / steps,
                    "h2d_copy_gbs": round(h2d_gbs, 1) if h2d_gbs else None, "pipeline_depth": 3},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "selfcheck": "ok", "remeasured": remeasured,
        }
    del ens, dev_in, host
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


# every BASELINE.json configuration is timed in the default run, after the headline (configs[1])
EXTRA_WORKLOADS = ["c3d_single_b8", "i3d64_ens", "twostream64_ens", "global_hetero", "i3d20_ens", "twostream20_ens",
                   "r3d34_ens"]
STRONG_WORKLOADS = {"global_hetero"}       # N > 1: ONE 256-clip batch split into (member, clip-chunk) units


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cse_b200 import runtime as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt.load_library()
    line = measure(args.workload, args, args.steps, args.warmup, True)
    extra = []
    if args.workload == "c3d_ens" and not args.no_workloads and args.shard == "clips" and not (args.batch or args.members or args.micro_batch):
        for name in EXTRA_WORKLOADS:
            r = measure(name, args, max(3, min(args.steps, 8)), 3, False)
            if r is not None:
                extra.append({k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "e2e", "selfcheck",
                                                "gpu_launches", "remeasured")}
                             | {"workload": name, "config": r["config"],
                                "roofline": {k: r["roofline"].get(k) for k in ("achieved", "peak", "unit", "frac", "kernel_share_of_step",
                                                                               "whole_step_model_tflops", "hbm_kernels", "top_ops", "note")}})
    if rank == 0:
        groups, _ = WORKLOADS[args.workload]
        cpu = None
        if not args.no_cpu_baseline:
            sample = cpu_sample_size(groups)
            members = sum(g[2] for g in groups)
            v, sec, threads = cpu_reference_clips_per_s(groups, sample, 2, 1)
            cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                   "sample": "%d clips x %d members per pass, 1 warm-up + 2 timed passes, torch CPU fp32 oracle "
                             "restatement (Keras 2.2.4/TF 1.15 not installable offline)" % (sample, members)}
        line["cpu_baseline"] = cpu
        if extra:
            line["workloads"] = extra
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
    ap.add_argument("--shard", default="clips", choices=["clips", "members", "units"],
                    help="N > 1: shard the clips (default, weak scaling), the ensemble members, or (member, clip-chunk) "
                         "units of ONE batch (strong scaling)")
    ap.add_argument("--no-workloads", action="store_true", help="time the headline workload only")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
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
