"""Ensemble member on one GPU: the object ``evaluate_load_model`` returns.

Mirrors the part of the Keras ``Model`` surface the reference's evaluation path touches
(evaluate_ensemble.py:1051-1056): ``compile(**kw)`` (no-op), ``predict_generator(generator,
workers=, use_multiprocessing=, verbose=)``, plus ``predict`` / ``predict_on_batch``.
Inputs are the generator's clips: NDHWC, BGR, 0..255, uint8 or float32 holding integer values
(train.py:257-291, 466-478); TwoStream members take ``[rgb, flow]`` (train.py:1009).

All arithmetic runs in libcse_b200 (no CPU fallback).  The reference hard-wires batch 1
(evaluate_ensemble.py:1032-1040); here clips are batched up to ``max_batch`` - results are
batch-size invariant.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np

from . import runtime as rt
from .graph import Graph, build_model_graph
from .lowering import Plan, lower


class Member:
    def __init__(self, graph: Graph, weights: Dict[str, List[np.ndarray]], precision: str = "bf16",
                 max_batch: int = 8, device=None, workspace=None, **lower_kw):
        torch = rt.require_cuda()
        self.lib = rt.load_library()
        self.torch = torch
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.graph = graph
        self.precision = precision
        self.max_batch = int(max_batch)
        self.plan: Plan = lower(graph, weights, precision, self.max_batch, **lower_kw)
        self.input_dtypes = tuple(lower_kw.get("input_dtypes") or ("u8",) * len(graph.inputs))
        self._weights, self._lower_kw, self._variants = weights, dict(lower_kw), {}
        self.nb_classes = self.plan.nb_classes
        self.input_shapes = [graph.shape(n) for n in graph.inputs]
        with torch.cuda.device(self.device):
            if workspace is not None:       # members that run back to back may share one activation arena
                if workspace.numel() < self.plan.workspace_bytes + 1024:
                    raise rt.CseError("shared workspace too small")
                self.workspace = workspace
            else:
                self.workspace = torch.empty(self.plan.workspace_bytes + 1024, dtype=torch.uint8,
                                             device=self.device)
            self.weights_dev = torch.from_numpy(self.plan.weight_arena).to(self.device)
            ws_ptr = (self.workspace.data_ptr() + 1023) // 1024 * 1024
            if self.weights_dev.data_ptr() % 256:
                raise rt.CseError("weight arena is not 256-byte aligned")
            handle = C.c_void_p()
            rt.check(self.lib.cse_plan_create(C.byref(handle), self.max_batch, self.nb_classes))
            self.handle = handle
            for s in self.plan.to_structs():
                rt.check(self.lib.cse_plan_add_op(self.handle, C.byref(s)))
            rt.check(self.lib.cse_plan_finalize(self.handle, ws_ptr, self.plan.workspace_bytes,
                                                self.weights_dev.data_ptr(), self.weights_dev.numel(),
                                                self.plan.logits.byte_off() if self.plan.logits else -1,
                                                self.plan.probs.byte_off() if self.plan.probs else -1))
            self._ws_ptr = ws_ptr
        self.launches = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.cse_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- Keras-surface ------------------------------------------------------ #
    def compile(self, *a, **kw):           # evaluate_ensemble.py:1052 - irrelevant for predict
        return None

    @property
    def num_ops(self) -> int:
        return len(self.plan.ops)

    def flops_per_clip(self) -> float:
        return self.graph.total_flops()

    # ---- device-level forward ------------------------------------------------ #
    def forward_device(self, inputs_u8, logits_out=None, probs_out=None, skip_input_ops: bool = False):
        """inputs_u8: list of uint8 CUDA tensors [n,T,H,W,C] (rgb[, flow]); runs on the current
        stream; returns (logits, probs) fp32 CUDA tensors [n, nb_classes]."""
        torch = self.torch
        if not isinstance(inputs_u8, (list, tuple)):
            inputs_u8 = [inputs_u8]
        if len(inputs_u8) != len(self.input_shapes):
            raise ValueError("model %s takes %d input(s), got %d" % (self.graph.name, len(self.input_shapes),
                                                                     len(inputs_u8)))
        n = inputs_u8[0].shape[0]
        if n > self.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (n, self.max_batch))
        for x, shp, dt in zip(inputs_u8, self.input_shapes, self.input_dtypes):
            want = torch.uint8 if dt == "u8" else torch.float32
            if not (x.is_cuda and x.dtype == want and x.is_contiguous()):
                raise ValueError("inputs must be contiguous %s CUDA tensors" % ("uint8" if dt == "u8" else "float32"))
            if tuple(x.shape[1:]) != tuple(shp) or x.shape[0] != n:
                raise ValueError("input shape %r does not match model input %r" % (tuple(x.shape), (n,) + tuple(shp)))
        if logits_out is None:
            logits_out = torch.empty((n, self.nb_classes), dtype=torch.float32, device=self.device)
        if probs_out is None:
            probs_out = torch.empty((n, self.nb_classes), dtype=torch.float32, device=self.device)
        flow_ptr = inputs_u8[1].data_ptr() if len(inputs_u8) > 1 else None
        # skip_input_ops: the pre-processed clips of this batch are already in the (shared) workspace,
        # written by another member of the same architecture (DeviceEnsemble)
        first = self.lib.cse_plan_num_input_ops(self.handle) if skip_input_ops else 0
        rt.check(self.lib.cse_plan_run_from(self.handle, inputs_u8[0].data_ptr(), flow_ptr, n, first,
                                            logits_out.data_ptr(), probs_out.data_ptr(), rt.current_stream_ptr()))
        self.launches = self.lib.cse_plan_last_launches(self.handle)
        return logits_out, probs_out

    def run_ops(self, inputs_u8, first: int, last: int):
        if not isinstance(inputs_u8, (list, tuple)):
            inputs_u8 = [inputs_u8]
        n = inputs_u8[0].shape[0]
        flow_ptr = inputs_u8[1].data_ptr() if len(inputs_u8) > 1 else None
        rt.check(self.lib.cse_plan_run_range(self.handle, inputs_u8[0].data_ptr(), flow_ptr, n, first, last,
                                             rt.current_stream_ptr()))

    def read_tensor(self, ref, n: int):
        """Copy a workspace view (lowering.TRef) back as a float32 numpy array [n,D,H,W,C]."""
        torch = self.torch
        d, h, w = ref.dims
        tdt = torch.float32 if ref.dtype == rt.F32 else torch.bfloat16
        es = 4 if ref.dtype == rt.F32 else 2
        start = (self._ws_ptr - self.workspace.data_ptr()) + ref.buf.offset
        wp = ref.wpitch or w
        nelem = n * d * h * wp * ref.ld
        flat = self.workspace[start:start + nelem * es].view(tdt)
        coff = ref.coff + (ref.C if getattr(ref, "unroll_w", 0) else 0)  # unrolled stem input: slot 1 = the pixel itself
        t = flat.view(n, d, h, wp, ref.ld)[:, :, :, ref.wpad:ref.wpad + w, coff:coff + ref.C]
        return t.float().cpu().numpy()

    # ---- host-level API -------------------------------------------------------- #
    @staticmethod
    def _as_u8(x):
        if hasattr(x, "is_cuda"):              # torch tensor: clips assembled on the GPU (clips.ClipSequence(device=...))
            if not (x.is_cuda and str(x.dtype) in ("torch.uint8", "torch.float32")):
                raise ValueError("tensor clips must be uint8 (frames) or float32 (on-the-fly flow) CUDA tensors")
            return x.contiguous()
        a = np.asarray(x)
        if a.dtype == np.uint8:
            return np.ascontiguousarray(a)
        r = np.rint(a)
        if not np.array_equal(r, a) or r.min() < 0 or r.max() > 255:
            if a.dtype in (np.float32, np.float64):
                return np.ascontiguousarray(a, dtype=np.float32)     # dense optical flow (train.py:294-332): stays float
            raise ValueError("clips must hold integer values 0..255 (decoded frames) or float32 flow; got %s" % a.dtype)
        return np.ascontiguousarray(r.astype(np.uint8))

    def _for_inputs(self, xs):
        """The member lowered for the dtypes of these inputs: decoded frames are uint8 (1 byte per value on the way to
        the GPU), the FarneBack_onTheFly flow is float32.  A member is lowered for uint8 inputs unless told otherwise
        (input_dtypes); the float variant shares graph and weights and is built on first use."""
        dts = tuple("f32" if ("float" in str(v.dtype)) else "u8" for v in xs)
        if dts == self.input_dtypes:
            return self
        if dts not in self._variants:
            kw = dict(self._lower_kw, input_dtypes=dts)
            self._variants[dts] = Member(self.graph, self._weights, precision=self.precision, max_batch=self.max_batch,
                                         device=self.device, **kw)
        return self._variants[dts]

    def predict(self, x, batch_size: Optional[int] = None, return_logits: bool = False):
        """x: NDHWC array (or [rgb, flow]) -> float32 [N, nb_classes] probabilities."""
        torch = self.torch
        xs = [self._as_u8(v) for v in (x if isinstance(x, (list, tuple)) else [x])]
        me = self._for_inputs(xs)
        n = xs[0].shape[0]
        bs = min(batch_size or self.max_batch, self.max_batch)
        probs = np.empty((n, self.nb_classes), np.float32)
        logits = np.empty((n, self.nb_classes), np.float32)
        with torch.cuda.device(self.device):
            for i in range(0, n, bs):
                dev = [v[i:i + bs].to(self.device).contiguous() if hasattr(v, "is_cuda")
                       else torch.from_numpy(v[i:i + bs]).to(self.device, non_blocking=False) for v in xs]
                lg, pr = me.forward_device(dev)
                probs[i:i + bs] = pr.cpu().numpy()
                logits[i:i + bs] = lg.cpu().numpy()
        return (probs, logits) if return_logits else probs

    def predict_on_batch(self, x):
        return self.predict(x)

    def predict_generator(self, generator, steps=None, workers=1, use_multiprocessing=False, verbose=0,
                          max_queue_size=10):
        """keras.utils.Sequence protocol (train.py:378-411): iterate len(generator) batches in order,
        take x of (x, y), concatenate outputs along axis 0."""
        steps = len(generator) if steps is None else steps
        pend: List[List[np.ndarray]] = []
        outs = []

        def flush():
            if not pend:
                return
            ninp = len(pend[0])
            cat = [self.torch.cat([p[j] for p in pend], dim=0) if hasattr(pend[0][j], "is_cuda")
                   else np.concatenate([p[j] for p in pend], axis=0) for j in range(ninp)]
            outs.append(self.predict(cat if ninp > 1 else cat[0]))
            pend.clear()

        count = 0
        for i in range(steps):
            item = generator[i]
            x = item[0] if isinstance(item, tuple) else item
            xs = [self._as_u8(v) for v in (x if isinstance(x, (list, tuple)) else [x])]
            pend.append(xs)
            count += xs[0].shape[0]
            if count >= self.max_batch:
                flush()
                count = 0
        flush()
        return np.concatenate(outs, axis=0) if outs else np.zeros((0, self.nb_classes), np.float32)


def build_member(model_type: str, weights, input_shape=None, nb_classes: int = 11, precision: str = "bf16",
                 max_batch: int = 8, device=None, **kw) -> Member:
    g = build_model_graph(model_type, input_shape, nb_classes)
    return Member(g, weights, precision=precision, max_batch=max_batch, device=device, **kw)
