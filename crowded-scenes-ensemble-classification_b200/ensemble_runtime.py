"""Device-resident ensemble: M members of one architecture + the soft vote, on one GPU.

This is the batched replacement of the reference's member loop (evaluate_ensemble.py:1048-1060,
one Keras model per member, batch 1, every clip re-decoded per member) followed by
ensemble_predictions (evaluate_ensemble.py:343-370): each micro-batch of clips is uploaded once and
run through every member while it is still L2/HBM resident; member probabilities land directly in
one [M, N, C] device buffer that the vote kernel reduces.  Members share one workspace arena
(they run back to back on one stream), so HBM holds M weight arenas + one set of activations.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import runtime as rt
from .graph import Graph
from .model import Member


class DeviceEnsemble:
    def __init__(self, graph: Graph, weight_sets: Sequence[Dict[str, List[np.ndarray]]], precision: str = "bf16",
                 max_batch: int = 256, micro_batch: int = 0, device=None, vote_weights=None, vote_mode: str = "SUM",
                 probs_out=None, logits_out=None, **lower_kw):
        torch = rt.require_cuda()
        self.torch = torch
        self.graph = graph
        self.max_batch = int(max_batch)
        self.micro_batch = int(micro_batch) if micro_batch else min(self.max_batch, 128)
        self.members: List[Member] = []
        shared = None
        # all members share one workspace and one lowering, hence the offset of the pre-processed
        # clip tensor: it is written once per micro-batch by the first member and kept alive
        self.share_input = bool(lower_kw.pop("share_input", True)) and len(weight_sets) > 1
        # fuse_stems: members 2k and 2k+1 read the same pre-processed clips, so their 64-filter 7x7x7 stems run as ONE
        # tcgen05 GEMM with N = 128 (the N = 64 MMA is bound by the shared-memory operand port, not by the math):
        # member 2k is lowered as the "lead" (its stem op carries the peer's stem weights and stores the peer's
        # activations into a persistent buffer), member 2k+1 as the "follow" (no stem op; its first consumers read that
        # buffer).  Results are bit-identical to the stand-alone members (same K order, per-column scale / shift).
        from .lowering import Lowerer
        plain = all(lower_kw.get(k, True) is True for k in ("tc", "s2d_stem", "stem_halo")) and "stem_role" not in lower_kw
        self.fuse_stems = (bool(lower_kw.pop("fuse_stems", True)) and self.share_input and plain
                           and Lowerer.stem_fusable(graph, precision))
        if self.share_input:
            lower_kw = dict(lower_kw, persist_input=True)
        self._member_kw = dict(precision=precision, max_batch=self.micro_batch, device=device, **lower_kw)
        self._weight_sets = list(weight_sets)
        self._solo: Dict[int, Member] = {}
        self.roles: List[Optional[str]] = []
        for j, w in enumerate(weight_sets):
            role_kw = {}
            if self.fuse_stems and j % 2 == 0 and j + 1 < len(weight_sets):
                role_kw = dict(stem_role="lead", stem_peer=weight_sets[j + 1])
            elif self.fuse_stems and j % 2 == 1:
                role_kw = dict(stem_role="follow")
            self.roles.append(role_kw.get("stem_role"))
            m = Member(graph, w, workspace=shared, **self._member_kw, **role_kw)
            shared = m.workspace
            self.members.append(m)
        self.device = self.members[0].device
        self.nb_classes = self.members[0].nb_classes
        self.M = len(self.members)
        shape = (self.M, self.max_batch, self.nb_classes)
        # probs_out / logits_out: slices of a wider [sum M, N, C] buffer (heterogeneous ensembles)
        for t in (probs_out, logits_out):
            if t is not None and (tuple(t.shape) != shape or t.dtype != torch.float32 or not t.is_contiguous()):
                raise ValueError("external probability buffer must be contiguous float32 %r" % (shape,))
        self.probs = probs_out if probs_out is not None else torch.empty(shape, dtype=torch.float32, device=self.device)
        self.logits = logits_out if logits_out is not None else torch.empty(shape, dtype=torch.float32,
                                                                             device=self.device)
        self.vote_mode = vote_mode
        self.vote_weights = None
        if vote_weights is not None:
            self.vote_weights = torch.as_tensor(np.asarray(vote_weights, np.float64)).to(self.device)
        self.last_launches = 0
        self._stage = None

    # ---- device path --------------------------------------------------------- #
    def forward_members(self, inputs_u8):
        """inputs: list of uint8 CUDA tensors [n,...] -> fills self.probs[:, :n], self.logits[:, :n]."""
        n = inputs_u8[0].shape[0]
        if n > self.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (n, self.max_batch))
        launches = 0
        mb = self.micro_batch
        for i in range(0, n, mb):
            chunk = [x[i:i + mb] for x in inputs_u8]
            for j, m in enumerate(self.members):
                m.forward_device(chunk, self.logits[j, i:i + mb], self.probs[j, i:i + mb],
                                 skip_input_ops=self.share_input and j > 0)
                launches += m.launches              # kernels only (the two D2D copies of logits / probs are memcpys)
        self.last_launches = launches
        return n

    def forward_subset(self, inputs_u8, member_ids, lo: int = 0):
        """Runs only the members `member_ids` (indices into self.members) on the given clips and writes their rows
        probs[j, lo : lo + n] (unit-sharded steps: this rank owns clip chunk [lo, lo + n) of these members)."""
        n = inputs_u8[0].shape[0]
        if lo + n > self.max_batch:
            raise ValueError("clip range [%d, %d) exceeds max_batch %d" % (lo, lo + n, self.max_batch))
        launches = 0
        mb = self.micro_batch
        ids = list(member_ids)
        run = []          # a leader runs its fused plan only when its follower comes right after it; anyone else runs alone
        for k, j in enumerate(ids):
            paired = ((self.roles[j] == "lead" and k + 1 < len(ids) and ids[k + 1] == j + 1) or
                      (self.roles[j] == "follow" and k > 0 and ids[k - 1] == j - 1))
            run.append(self.members[j] if (paired or self.roles[j] is None) else self.solo_member(j))
        for i in range(0, n, mb):
            chunk = [x[i:i + mb] for x in inputs_u8]
            written = False              # the pre-processed clips of this chunk are in the shared workspace
            for j, m in zip(ids, run):
                shared = m.workspace is self.members[0].workspace
                m.forward_device(chunk, self.logits[j, lo + i:lo + i + mb], self.probs[j, lo + i:lo + i + mb],
                                 skip_input_ops=self.share_input and shared and written)
                written = written or shared
                launches += m.launches
        self.last_launches += launches
        return n

    def solo_member(self, j: int) -> Member:
        """Member j lowered stand-alone (its own stem op), for clip chunks its stem partner does not run on this GPU
        (unit-sharded steps).  Built on first use - i.e. during warm-up, before any CUDA graph is captured - on the shared
        workspace (the pre-processed clips sit at the same offset in every lowering of the ensemble)."""
        if j not in self._solo:
            try:
                self._solo[j] = Member(self.graph, self._weight_sets[j], workspace=self.members[0].workspace, **self._member_kw)
            except rt.CseError:          # a stand-alone plan that needs more room than the paired ones: its own arena,
                self._solo[j] = Member(self.graph, self._weight_sets[j], **self._member_kw)      # its own pre-processing
        return self._solo[j]

    def vote(self, n):
        probs = self.probs[:, :n].contiguous() if n != self.max_batch else self.probs
        pred = rt.vote(probs, self.vote_weights, self.vote_mode)
        self.last_launches += 1
        return pred

    def predict_device(self, inputs_u8):
        n = self.forward_members(inputs_u8)
        return self.vote(n)

    def predict_host(self, host_inputs):
        """host_inputs: list of pinned uint8 CPU tensors [n,...]; uploads, runs, returns numpy int32 [n]."""
        torch = self.torch
        dev = [h.to(self.device, non_blocking=True) for h in host_inputs]
        pred = self.predict_device(dev)
        return pred.cpu().numpy()

    # ---- profiling -------------------------------------------------------------- #
    def profile_ops(self, inputs_u8, iters: int = 2):
        """CUDA-event duration of every op launch (summed over members and micro-batches) for one
        step; events are recorded on the stream the kernels are launched on.  Every member runs the ops of its own plan
        (a stem leader's fused stem op counts the FLOPs of both members, its follower has no stem op); the rows are
        keyed by op name in the order of member 0's plan."""
        torch = self.torch
        stream = torch.cuda.current_stream()
        n = inputs_u8[0].shape[0]
        mb = self.micro_batch
        rows: Dict[str, dict] = {}
        eng = {0: "", 1: "direct", 2: "tcgen05"}
        for m in self.members:
            for op in m.plan.ops:
                key = op.name[:-5] if op.name.endswith("+peer") else op.name
                rows.setdefault(key, {"name": key, "kind": rt.OP_NAMES[op.kind], "engine": eng[op.engine], "ms": 0.0,
                                      "flops": 0.0, "bytes": 0.0})
        for it in range(iters + 1):
            evs = []
            for i in range(0, n, mb):
                chunk = [x[i:i + mb] for x in inputs_u8]
                for jm, m in enumerate(self.members):
                    for k, op in enumerate(m.plan.ops):
                        if self.share_input and jm > 0 and op.kind == rt.OP_PREPROCESS:
                            continue            # pre-processed clips are shared with member 0
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(stream)
                        m.run_ops(chunk, k, k + 1)
                        b.record(stream)
                        evs.append((op, a, b))
            torch.cuda.synchronize()
            if it == 0:
                continue            # warm-up pass
            for op, a, b in evs:
                key = op.name[:-5] if op.name.endswith("+peer") else op.name
                r = rows[key]
                r["ms"] += a.elapsed_time(b) / iters
        # algorithmic FLOPs / bytes of one step (n clips through every member), independent of the micro-batching
        for jm, m in enumerate(self.members):
            for op in m.plan.ops:
                if self.share_input and jm > 0 and op.kind == rt.OP_PREPROCESS:
                    continue
                key = op.name[:-5] if op.name.endswith("+peer") else op.name
                rows[key]["flops"] += op.flops * n
                rows[key]["bytes"] += _algorithmic_bytes(op) * n
        return list(rows.values())


def _algorithmic_bytes(op) -> float:
    """Algorithmic HBM bytes per clip of the HBM-bound ops: preprocess = uint8 clip read +
    layout bytes written; pool / affine / add = (input + output elements) x element size.  0 for convs
    (they are reported against the tensor roofline)."""
    def elems(r):
        return float(np.prod(r.dims)) * r.C
    if op.kind == rt.OP_PREPROCESS:
        t, h, w, c = op.src_dims
        o = op.out0
        return float(t * h * w * c) + float(o.dims[0] * o.dims[1] * (o.wpitch or o.dims[2]) * o.ld * o.esize)
    if op.kind in (rt.OP_MAXPOOL3D, rt.OP_AVGPOOL3D, rt.OP_AFFINE):
        return elems(op.in0) * op.in0.esize + elems(op.out0) * op.out0.esize
    if op.kind == rt.OP_ADD:
        return (elems(op.in0) + elems(op.in1)) * op.in0.esize + elems(op.out0) * op.out0.esize
    return 0.0


class HeteroEnsemble:
    """Global (heterogeneous) ensemble on one GPU: several architectures, each with its own clip
    geometry and its fold members, voting together (global_evaluate_ensembles,
    evaluate_ensemble.py:1329-1474: weights = ones(n_arch * (folds-1)), :1455).  Every group writes
    its members' probabilities into its slice of one [sum M, N, C] buffer; one vote kernel reduces
    it in member order (architecture order of the models list, then val-fold order)."""

    def __init__(self, groups, precision: str = "bf16", max_batch: int = 256, device=None, vote_weights=None,
                 vote_mode: str = "SUM", use_graphs: bool = True, **lower_kw):
        """groups: list of (graph, weight_sets, micro_batch)."""
        torch = rt.require_cuda()
        self.torch = torch
        self.max_batch = int(max_batch)
        classes = {g.shape(g.output)[-1] for g, _, _ in groups}
        if len(classes) != 1:
            raise ValueError("all architectures of a global ensemble must share nb_classes")
        self.nb_classes = classes.pop()
        self.M = sum(len(ws) for _, ws, _ in groups)
        dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.probs = torch.empty((self.M, self.max_batch, self.nb_classes), dtype=torch.float32, device=dev)
        self.logits = torch.empty_like(self.probs)
        self.groups: List[DeviceEnsemble] = []
        m0 = 0
        for g, ws, mb in groups:
            m1 = m0 + len(ws)
            self.groups.append(DeviceEnsemble(g, ws, precision=precision, max_batch=self.max_batch, micro_batch=mb,
                                              device=dev, probs_out=self.probs[m0:m1], logits_out=self.logits[m0:m1],
                                              **lower_kw))
            m0 = m1
        self.device = dev
        self.vote_mode = vote_mode
        self.vote_weights = None
        if vote_weights is not None:
            self.vote_weights = torch.as_tensor(np.asarray(vote_weights, np.float64)).to(dev)
        self.last_launches = 0
        # set by the caller when the members are sharded over the ranks (ensemble.gather_member_probs)
        self.gather = None
        self._unit_plan = None
        self.use_graphs = bool(use_graphs)
        self._graphs = {}
        self._seen_once = set()

    @property
    def micro_batch(self):
        return [g.micro_batch for g in self.groups]

    # ---- CUDA graphs --------------------------------------------------------------------------------------------
    # A step is a fixed sequence of kernel launches on fixed addresses (cse_plan_run allocates nothing, the
    # workspaces / probability buffers are owned by the ensemble), so it is captured once per (input buffers, n)
    # into a CUDA graph and replayed: the launch-bound workloads (single C3D at batch 8: 13 kernels in 0.7 ms;
    # R3D-34: 164 kernels per step) no longer pay one CPU launch per kernel.  Steps that end in a collective
    # (member- / unit-sharded gathers) capture the member forwards only.
    MAX_GRAPHS = 8

    def _graph_key(self, group_inputs):
        return tuple((x.data_ptr(), tuple(x.shape)) for inputs in group_inputs for x in inputs)

    def predict_device(self, group_inputs, graph: bool = False):
        """group_inputs[k] = list of uint8 CUDA tensors for architecture k (same n for all).  Runs every member of every
        group and the soft vote on the current stream -> int32 [n] predictions (device).  graph=True: the caller
        keeps these input buffers alive and refills them in place (stream_host's buffer sets, a resident batch): the
        step is captured into a CUDA graph the second time the buffers are seen and replayed from then on; the
        returned tensor is then overwritten by the next step on the same buffers."""
        if not (graph and self.use_graphs):
            return self._predict_eager(group_inputs)
        torch = self.torch
        key = self._graph_key(group_inputs)
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= self.MAX_GRAPHS:
                return self._predict_eager(group_inputs)
            if key not in self._seen_once:
                # first sight of these buffers: run eagerly (also warms up per-device kernel attributes); capture
                # when the same buffers come back
                self._seen_once.add(key)
                return self._predict_eager(group_inputs)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    out = self._predict_eager(group_inputs, collective=False)
            cur.wait_stream(side)
            entry = (g, out, self.last_launches, [x for inputs in group_inputs for x in inputs])
            self._graphs[key] = entry
        g, out, launches, _keepalive = entry
        g.replay()
        self.last_launches = launches
        if self.gather is not None:
            return self._finish_collective(out)
        return out

    def _finish_collective(self, probs):
        probs = self.gather(probs)
        pred = rt.vote(probs, self.vote_weights, self.vote_mode)
        self.last_launches += 1
        return pred

    def _predict_eager(self, group_inputs, collective: bool = True):
        if getattr(self, "_unit_plan", None) is not None:
            return self.predict_units(group_inputs, collective)
        n = group_inputs[0][0].shape[0]
        launches = 0
        if self.torch.cuda.current_device() != self.device.index:
            raise rt.CseError("HeteroEnsemble on %s called while cuda:%d is current; wrap the call in "
                              "torch.cuda.device(ens.device)" % (self.device, self.torch.cuda.current_device()))
        for ens, inputs in zip(self.groups, group_inputs):
            if inputs[0].shape[0] != n:
                raise ValueError("every architecture must see the same clips")
            ens.forward_members(inputs)
            launches += ens.last_launches
        probs = self.probs[:, :n].contiguous() if n != self.max_batch else self.probs
        self.last_launches = launches
        if self.gather is not None:          # member-sharded partition: [M_local, n, C] -> [M, n, C] in member order
            return self._finish_collective(probs) if collective else probs
        pred = rt.vote(probs, self.vote_weights, self.vote_mode)
        self.last_launches = launches + 1
        return pred

    def set_units(self, my_units, gather):
        """Unit-sharded step (ensemble.shard_units): this rank runs only `my_units` = [(member, lo, hi)] (global
        member index in group order) and `gather` (ensemble.UnitGather) merges the blocks of all ranks."""
        plan = []
        m0 = 0
        for ens in self.groups:
            by_range = {}
            for m, lo, hi in my_units:
                if m0 <= m < m0 + ens.M:
                    by_range.setdefault((lo, hi), []).append(m - m0)
            plan.append(sorted((r, sorted(ms)) for r, ms in by_range.items()))     # members ascending: stem pairs stay adjacent
            m0 += ens.M
        self._unit_plan, self.gather = plan, gather
        self._graphs.clear()
        self._seen_once.clear()

    def upload_ranges(self):
        """Unit-sharded step: per group, the merged clip ranges [(lo, hi)] this rank's units read (None = everything)."""
        if getattr(self, "_unit_plan", None) is None:
            return None
        out = []
        for ranges in self._unit_plan:
            merged = []
            for (lo, hi), _ in sorted(ranges):
                if merged and lo <= merged[-1][1]:
                    merged[-1] = (merged[-1][0], max(merged[-1][1], hi))
                else:
                    merged.append((lo, hi))
            out.append(merged)
        return out

    def upload_bytes(self, host_group_inputs) -> int:
        """Bytes one stream_host step copies host -> device (all inputs, or the owned clip ranges of a unit-sharded step)."""
        needed = self.upload_ranges()
        total = 0
        for gi, hs in enumerate(host_group_inputs):
            for h in hs:
                per_clip = h[0].numel() * h.element_size()
                total += per_clip * (h.shape[0] if needed is None else sum(hi - lo for lo, hi in needed[gi]))
        return total

    def predict_units(self, group_inputs, collective: bool = True):
        """group_inputs as in predict_device (all n clips; only the owned ranges are read)."""
        n = group_inputs[0][0].shape[0]
        self.probs.zero_()
        launches = 1
        for ens, inputs, ranges in zip(self.groups, group_inputs, self._unit_plan):
            ens.last_launches = 0
            for (lo, hi), member_ids in ranges:
                ens.forward_subset([x[lo:hi] for x in inputs], member_ids, lo)
            launches += ens.last_launches
        probs = self.probs[:, :n].contiguous() if n != self.max_batch else self.probs
        self.last_launches = launches
        return self._finish_collective(probs) if collective else probs

    def predict_host(self, host_group_inputs):
        dev = [[h.to(self.device, non_blocking=True) for h in inputs] for inputs in host_group_inputs]
        return self.predict_device(dev).cpu().numpy()

    def _pipeline(self, batch, depth: int):
        """Copy stream, `depth` device buffer sets and their ready / free events, created once and reused by every
        stream_host call (they live on the object, like the buffers they guard)."""
        torch = self.torch
        shapes = [[tuple(h.shape) for h in hs] for hs in batch]
        p = getattr(self, "_pipe", None)
        if p is None or p["shapes"] != shapes or p["depth"] != depth:
            with torch.cuda.device(self.device):
                p = {"shapes": shapes, "depth": depth, "copy": torch.cuda.Stream(device=self.device),
                     "bufs": [[[torch.empty(sh, dtype=torch.uint8, device=self.device) for sh in hs] for hs in shapes]
                              for _ in range(depth)],
                     "ready": [torch.cuda.Event(enable_timing=True) for _ in range(depth)],
                     "start": [torch.cuda.Event(enable_timing=True) for _ in range(depth)],
                     "free": [torch.cuda.Event() for _ in range(depth)],
                     "used": [False] * depth, "bytes": sum(int(np.prod(sh)) for hs in shapes for sh in hs)}
            self._pipe = p
            self._graphs.clear()           # graphs captured on the previous buffer sets keep those buffers alive
            self._seen_once.clear()
        return p

    def stream_host(self, batches, depth: int = 3):
        """Pipelined host path: `batches` yields pinned uint8 host inputs (group_inputs layout); the H2D copies of
        the next depth-1 batches run on a copy stream while batch i computes.  Yields the int32 prediction tensor
        (device) of every batch, in order; the caller reads it back.  The buffer sets and their events persist on
        the object: a later call (or one started after an abandoned generator) first waits, on the copy stream,
        for the compute that last read the buffer set it is about to overwrite."""
        torch = self.torch
        it = iter(batches)
        first = next(it, None)
        if first is None:
            return
        with torch.cuda.device(self.device):
            comp = torch.cuda.current_stream()
            p = self._pipeline(first, depth)
            copy = p["copy"]

            needed = self.upload_ranges()      # unit-sharded step: only the clip ranges this rank runs
            self._last_upload_bytes = self.upload_bytes(first)

            def upload(k, batch):
                with torch.cuda.stream(copy):
                    if p["used"][k]:
                        copy.wait_event(p["free"][k])     # the compute that read this buffer set has finished
                    p["start"][k].record(copy)
                    for gi, (ds, hs) in enumerate(zip(p["bufs"][k], batch)):
                        for d, h in zip(ds, hs):
                            if needed is None:
                                d.copy_(h, non_blocking=True)
                            else:
                                for lo, hi in needed[gi]:
                                    d[lo:hi].copy_(h[lo:hi], non_blocking=True)
                    p["ready"][k].record(copy)

            queue = []                     # buffer sets uploaded and not yet computed, in order
            nxt_slot = 0
            pending = first
            while pending is not None or queue:
                while pending is not None and len(queue) < depth:
                    upload(nxt_slot, pending)
                    queue.append(nxt_slot)
                    nxt_slot = (nxt_slot + 1) % depth
                    pending = next(it, None)
                k = queue.pop(0)
                comp.wait_event(p["ready"][k])
                pred = self.predict_device(p["bufs"][k], graph=True)
                p["free"][k].record(comp)
                p["used"][k] = True
                yield pred

    def h2d_copy_gbs(self):
        """Achieved host -> device bandwidth of the most recent uploads of stream_host (CUDA events on the copy
        stream around each batch's copies; call after a synchronize)."""
        p = getattr(self, "_pipe", None)
        if p is None or not any(p["used"]):
            return None
        ms = [p["start"][k].elapsed_time(p["ready"][k]) for k in range(p["depth"]) if p["used"][k]]
        nbytes = getattr(self, "_last_upload_bytes", None) or p["bytes"]
        return nbytes / (min(ms) / 1e3) / 1e9 if ms and min(ms) > 0 else None

    def profile_ops(self, group_inputs, iters: int = 2):
        out = []
        for k, (ens, inputs) in enumerate(zip(self.groups, group_inputs)):
            for p in ens.profile_ops(inputs, iters):
                p = dict(p)
                p["name"] = "%s/%s" % (ens.graph.name, p["name"])
                out.append(p)
        return out
