"""Graph -> device plan: fusion, engine selection, weight packing, buffer planning.

The layer graph (graph.py, one node per Keras layer of the reference model) is lowered into
the flat list of fused ``cse_op`` records that libcse_b200 executes:

* Conv3D (+bias) [+BatchNormalization] [+ReLU]  -> one CONV3D op with a scale/shift/ReLU epilogue
  (conv3d_bn train.py:646-668; Conv3D(activation='relu') train.py:1230-1258; _conv_bn_relu3D :1283)
* add([shortcut, residual]) (train.py:1346)     -> folded into the residual conv's epilogue; the
  following BN-ReLU of the next pre-activation block (_bn_relu train.py:1278) becomes the op's
  second output, so each R3D block is two or three kernels
* concatenate (Inception Mixed blocks, train.py:1048-1193) -> no kernel: every branch writes its
  channel slice of the concat buffer
* ZeroPadding3D + MaxPooling3D (train.py:1259-1261) -> one pool with 0-valued padding
* Flatten -> view;  Dense -> CONV3D on [n,1,1,1,K];  softmax -> SOFTMAX op on fp32 logits
* TwoStream feature concat + Dense (train.py:1006-1007) -> two chained Dense ops
  (f_rgb @ W[:F] + b, then + f_flow @ W[F:]) so no concat copy exists

Weight folding is done in float32 exactly like TF's non-fused BN: inv = gamma*rsqrt(var+eps);
y = x*inv + (beta - mean*inv)  (SURVEY App. A.0).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import runtime as rt
from .graph import Graph, Node, BN_EPS
from .weights import check_weights

ALIGN = 1024


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


@dataclass
class Buf:
    name: str
    nbytes: int
    first: int = 1 << 30
    last: int = -1
    offset: int = -1
    pinned_to_end: bool = False


@dataclass
class TRef:
    """A [n,D,H,W,C] view: channel slice [coff, coff+C) of a buffer whose pixels have ld channels."""
    buf: Buf
    coff: int
    C: int
    ld: int
    dims: Tuple[int, int, int]
    dtype: int
    wpitch: int = 0          # row pitch in pixels (0 = dims[2]); > W for the W-padded packed-stem input
    wpad: int = 0            # zero columns on the left of each row
    unroll_w: int = 0        # packed stem: channels hold unroll_w neighbouring pixels x 8 channels
    s2d: int = 0             # stride-2 stem: 2x2 space-to-depth cells over (H, W); dims = (T, H/2, W/2)
    src_c: int = 0           # s2d: channels of the source clip

    @property
    def esize(self) -> int:
        return 4 if self.dtype == rt.F32 else 2

    def byte_off(self) -> int:
        return self.buf.offset + self.coff * self.esize


@dataclass
class DevOp:
    kind: int
    name: str
    in0: Optional[TRef] = None
    in1: Optional[TRef] = None
    out0: Optional[TRef] = None
    out1: Optional[TRef] = None
    engine: int = 0
    w_dtype: int = 0
    k: Tuple[int, int, int] = (1, 1, 1)
    s: Tuple[int, int, int] = (1, 1, 1)
    pad: Tuple[int, int, int] = (0, 0, 0)
    relu0: int = 0
    relu1: int = 0
    pad_is_zero: int = 0
    ext_input: int = 0
    crop: Tuple[int, int, int] = (0, 0, 0)
    src_dims: Tuple[int, int, int, int] = (0, 0, 0, 0)
    pre_mean: Tuple[float, ...] = (0.0, 0.0, 0.0, 0.0)
    pre_scale: Tuple[float, ...] = (1.0, 1.0, 1.0, 1.0)
    kc: int = 0
    bn: int = 0
    brick: Tuple[int, int, int, int] = (0, 0, 0, 0)
    halo: int = 0
    pair_pool: int = 0
    ksplit: int = 1               # TCGEN05 split-K factor
    part: Optional[Buf] = None    # fp32 partial tiles of a split-K conv (live during that op only)
    src_dtype: int = 2            # PREPROCESS: dtype of the external clip (rt.U8 frames / rt.F32 on-the-fly flow)
    out_split: int = 0            # fused sibling 1x1x1 convs: columns >= out_split go to out1, >= out_split2 to out2
    out_split2: int = 0
    out2: Optional[TRef] = None
    pool_k: Tuple[int, int, int] = (0, 0, 0)
    pool_zero: int = 0
    conv_out_dims: Optional[Tuple[int, int, int]] = None   # conv's own output dims when a pool is fused
    w_blob: int = -1
    scale0: int = -1
    shift0: int = -1
    scale1: int = -1
    shift1: int = -1
    flops: float = 0.0            # algorithmic 2*MACs per clip (un-padded reference shape)
    layers: Tuple[str, ...] = ()  # Keras layers folded into this op
    softmax_C: int = 0


@dataclass
class Plan:
    ops: List[DevOp]
    buffers: List[Buf]
    workspace_bytes: int
    weight_arena: np.ndarray          # uint8
    blob_offsets: List[int]
    logits: TRef
    probs: TRef
    max_batch: int
    nb_classes: int
    precision: str
    n_inputs: int
    tensors: Dict[str, object] = field(default_factory=dict)   # Keras layer name -> TRef

    def to_structs(self) -> List[rt.CseOp]:
        out = []
        for op in self.ops:
            s = rt.CseOp()
            s.kind, s.engine, s.w_dtype = op.kind, op.engine, op.w_dtype
            i0, o0 = op.in0, op.out0
            if op.kind == rt.OP_SOFTMAX:
                s.in_dims[:] = (1, 1, 1, op.softmax_C)
                s.out_dims[:] = (1, 1, 1, op.softmax_C)
                s.in_dtype = s.out_dtype = rt.F32
                s.in0_off, s.out0_off = i0.byte_off(), o0.byte_off()
                s.in1_off = s.out1_off = s.w_off = s.part_off = -1
                s.scale0_off = s.shift0_off = s.scale1_off = s.shift1_off = -1
                out.append(s)
                continue
            if i0 is not None:
                s.in_dims[:] = tuple(i0.dims) + (i0.C,)
                s.in_ld, s.in_dtype, s.in0_off = i0.ld, i0.dtype, i0.byte_off()
                s.in_wpitch = i0.wpitch
            else:
                s.in0_off = -1
            s.out_dims[:] = tuple(op.conv_out_dims or o0.dims) + (o0.C * (2 if op.pair_pool else 1),)
            s.tc_pair_pool = op.pair_pool
            s.out_split, s.out_split2 = op.out_split, op.out_split2
            s.out_ld, s.out_dtype, s.out0_off = o0.ld, o0.dtype, o0.byte_off()
            if op.pool_k[0] > 0:
                s.pool_k[:] = op.pool_k
                s.pool_dims[:] = tuple(o0.dims)
                s.pool_zero = op.pool_zero
            if op.kind == rt.OP_PREPROCESS:
                s.out_wpitch, s.out_wpad, s.pre_unroll_w, s.pre_s2d = o0.wpitch, o0.wpad, o0.unroll_w, o0.s2d
                s.in_dtype = op.src_dtype
            if op.in1 is not None:
                s.in1_ld, s.in1_off = op.in1.ld, op.in1.byte_off()
            else:
                s.in1_off = -1
            if op.out1 is not None:
                s.out1_ld, s.out1_off = op.out1.ld, op.out1.byte_off()
            else:
                s.out1_off = -1
            if op.out2 is not None:
                s.out2_ld, s.out2_off = op.out2.ld, op.out2.byte_off()
            else:
                s.out2_off = -1
            s.k[:], s.s[:], s.pad[:] = op.k, op.s, op.pad
            s.relu0, s.relu1, s.pad_is_zero, s.ext_input = op.relu0, op.relu1, op.pad_is_zero, op.ext_input
            s.crop[:] = op.crop
            s.src_dims[:] = op.src_dims
            s.pre_mean[:] = tuple(op.pre_mean)
            s.pre_scale[:] = tuple(op.pre_scale)
            s.kc, s.bn, s.tc_halo = op.kc, op.bn, op.halo
            s.brick[:] = op.brick
            bo = self.blob_offsets
            s.w_off = bo[op.w_blob] if op.w_blob >= 0 else -1
            s.scale0_off = bo[op.scale0] if op.scale0 >= 0 else -1
            s.shift0_off = bo[op.shift0] if op.shift0 >= 0 else -1
            s.scale1_off = bo[op.scale1] if op.scale1 >= 0 else -1
            s.shift1_off = bo[op.shift1] if op.shift1 >= 0 else -1
            s.ksplit, s.part_off, s.part_bytes = op.ksplit, (op.part.offset if op.part is not None else -1), \
                (op.part.nbytes if op.part is not None else 0)
            out.append(s)
        return out

    def describe(self) -> str:
        lines = []
        for i, op in enumerate(self.ops):
            eng = {0: "", 1: "direct", 2: "tcgen05"}[op.engine]
            lines.append("%3d %-10s %-8s %-40s out=%s%s" % (
                i, rt.OP_NAMES[op.kind], eng, op.name,
                (tuple(op.out0.dims) + (op.out0.C,)) if op.out0 else "",
                " kc=%d bn=%d brick=%s" % (op.kc, op.bn, op.brick) if op.engine == 2 else ""))
        return "\n".join(lines)


# --------------------------------------------------------------------------- #
def to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 (round to nearest even), returned as uint16 bit patterns."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16)


def fold_bn(bias, bn_weights, has_gamma, co):
    """(conv bias, BN tensors) -> float32 (scale, shift) with y = acc*scale + shift."""
    one = np.float32(1.0)
    if bn_weights is None:
        return None, (None if bias is None else bias.astype(np.float32))
    if has_gamma:
        gamma, beta, mean, var = bn_weights
    else:
        beta, mean, var = bn_weights
        gamma = None
    inv = (one / np.sqrt(var.astype(np.float32) + np.float32(BN_EPS))).astype(np.float32)
    if gamma is not None:
        inv = (inv * gamma.astype(np.float32)).astype(np.float32)
    shift = (beta.astype(np.float32) - mean.astype(np.float32) * inv).astype(np.float32)
    if bias is not None:
        shift = (bias.astype(np.float32) * inv + shift).astype(np.float32)
    return inv, shift


def choose_kc(ci: int) -> int:
    """Channels per pipeline stage (64/32/16 -> 128B/64B/32B swizzle).  Every stage costs a barrier
    round trip and one TMA issue on top of its kc/16 MMAs, so a few zero-padded channels (TMA fills
    the tail beyond Cin with zeros, the packed weights hold zeros there) beat many thin stages:
    Cin = 144 runs as 3 x 64 rather than 9 x 16."""
    best, best_cost = 64, None
    for kc in (64, 32, 16):
        chunks = -(-ci // kc)
        cost = chunks * (24 + kc)
        if best_cost is None or cost < best_cost:
            best, best_cost = kc, cost
    return best


SM_COUNT = 148      # B200: persistent grid = one CTA per SM


def choose_bn(co: int, m_tiles: int = 0) -> Tuple[int, int]:
    """N tile (multiple of 16, <= 256) and number of N tiles.  With m_tiles given, the split that
    minimises waves x per-tile cost on 148 SMs: a 128 x bn x 64 k-step costs max(MMA issue time,
    operand bytes / L2->SM bandwidth), so narrower N tiles are cheaper per tile but re-read A, and
    only pay off when the wide tiling would leave SMs idle (small-M layers: R3D stages 3-4, I3D 5x)."""
    best, best_key = None, None
    min_t = -(-co // 256)
    for n_tiles in range(min_t, max(min_t + 1, 9, co // 64 + 1)):      # up to 64-wide tiles (dense heads: M is 1-2 tiles)
        bn = _round_up(-(-co // n_tiles), 16)
        if n_tiles > min_t and bn < 64:
            break
        if m_tiles <= 0:
            return bn, n_tiles
        mma = 4 * max(bn / 2.0, 48.0)                 # cycles per k-step: 4 x (128 x bn x 16)
        l2 = (128 * 128 + bn * 128) / 85.0            # A + B bytes per k-step at ~85 B/clk/SM
        waves = -(-(m_tiles * n_tiles) // SM_COUNT)
        key = (waves * max(mma, l2), n_tiles)
        if best_key is None or key < best_key:
            best, best_key = (bn, n_tiles), key
    return best


def choose_brick(nb: int, do: int, ho: int, wo: int, mult=(1, 1, 1)) -> Tuple[int, int, int, int]:
    """Output-pixel brick (n,d,h,w), product <= 128, minimising the number of M tiles.  `mult`:
    the (d,h,w) brick dims must be multiples of a fused pool window."""
    best, best_key = None, None
    md, mh, mw = mult
    for bw in range(mw, min(_round_up(wo, mw), 128) + 1, mw):
        tw = -(-wo // bw)
        for bh in range(mh, min(_round_up(ho, mh), 128 // bw) + 1, mh):
            th = -(-ho // bh)
            for bd in range(md, min(_round_up(do, md), 128 // (bw * bh)) + 1, md):
                td = -(-do // bd)
                bn_ = max(1, min(nb, 128 // (bw * bh * bd)))
                tn = -(-nb // bn_)
                tiles = tn * td * th * tw
                rows = bn_ * bd * bh * bw
                key = (tiles, bn_, -bw, -bh)
                if best_key is None or key < best_key:
                    best, best_key = (bn_, bd, bh, bw), key
    return best


def pack_tc_weights(kernel: np.ndarray, kc: int, bn: int, n_tiles: int) -> np.ndarray:
    """Keras [kd,kh,kw,Ci,Co] -> [Co_pad][taps*kchunks*kc] bf16 bits (K-major rows)."""
    kd, kh, kw, ci, co = kernel.shape
    taps = kd * kh * kw
    cip = _round_up(ci, kc)
    w = np.zeros((taps, cip, n_tiles * bn), np.float32)
    w[:, :ci, :co] = kernel.reshape(taps, ci, co)
    w = np.ascontiguousarray(w.reshape(taps * cip, n_tiles * bn).T)
    return to_bf16_bits(w)


def pack_tc_weights_halo(kernel: np.ndarray, kc: int, bn: int, n_tiles: int) -> np.ndarray:
    """Keras [kd,kh,1,Ci,Co] -> [n_tile][tap][bn][kc] bf16 bits (halo mode: all taps of an N tile
    are consecutive row blocks of one 2-D tensor with kc columns)."""
    kd, kh, kw, ci, co = kernel.shape
    assert kw == 1 and ci <= kc
    taps = kd * kh
    w = np.zeros((n_tiles * bn, taps, kc), np.float32)
    w[:co, :, :ci] = kernel.reshape(taps, ci, co).transpose(2, 0, 1)
    w = w.reshape(n_tiles, bn, taps, kc).transpose(0, 2, 1, 3)
    return to_bf16_bits(np.ascontiguousarray(w).reshape(n_tiles * taps * bn, kc))


def pack_tc_weights_hhalo(kernel: np.ndarray, kc: int, bn: int, n_tiles: int) -> np.ndarray:
    """Keras [kd,kh,kw,Ci,Co] -> [n_tile][fd][fw][chunk][fh][bn][kc] bf16 bits (h-halo modes: the kh taps of one
    (fd, fw, channel chunk) are consecutive row blocks of one 2-D tensor with kc columns; kw = 1 for tc_halo = 2)."""
    kd, kh, kw, ci, co = kernel.shape
    nch = -(-ci // kc)
    w = np.zeros((n_tiles * bn, kd, kh, kw, nch * kc), np.float32)
    w[:co, :, :, :, :ci] = kernel.transpose(4, 0, 1, 2, 3)
    w = w.reshape(n_tiles, bn, kd, kh, kw, nch, kc).transpose(0, 2, 4, 5, 3, 1, 6)
    return to_bf16_bits(np.ascontiguousarray(w).reshape(-1, kc))


def choose_brick_pair_halo(ho: int, wo: int, kh: int):
    """Brick (1,1,h,w), w % 8 == 0, for the CTA-pair h-halo conv: fewest 128-row tiles per plane first (the layer is
    MMA-bound there), then the fewest rows loaded.  -> (brick, fraction of the MMA rows that are real outputs)."""
    best, best_key = None, None
    for bw in range(8, min(_round_up(wo, 8), 128) + 1, 8):
        for bh in range(1, 128 // bw + 1):
            tiles = -(-ho // bh) * -(-wo // bw)
            key = (tiles, tiles * (bh + kh - 1) * bw)
            if best_key is None or key < best_key:
                best, best_key = (1, 1, bh, bw), key
    return best, ho * wo / float(best_key[0] * 128)


def choose_brick_hhalo(ho: int, wo: int, kh: int) -> Tuple[int, int, int, int]:
    """Brick (1,1,h,w), w % 8 == 0, minimising the rows loaded per output plane (tiles x haloed box)."""
    best, best_key = None, None
    for bw in range(8, min(_round_up(wo, 8), 128) + 1, 8):
        for bh in range(1, 128 // bw + 1):
            tiles = -(-ho // bh) * -(-wo // bw)
            key = (tiles * (bh + kh - 1) * bw, tiles, -bw)
            if best_key is None or key < best_key:
                best, best_key = (1, 1, bh, bw), key
    return best


def choose_brick_hw(ho: int, wo: int, mult=(1, 1, 1)) -> Optional[Tuple[int, int, int, int]]:
    """Brick (1,1,h,w) with w % 8 == 0 (halo mode: a one-row shift must be a whole swizzle atom)."""
    best, best_key = None, None
    if mult[0] != 1:
        return None
    step_w = 8 * mult[2] // math.gcd(8, mult[2])
    for bw in range(step_w, min(_round_up(wo, step_w), 128) + 1, step_w):
        for bh in range(mult[1], 128 // bw + 1, mult[1]):
            tiles = -(-ho // bh) * -(-wo // bw)
            key = (tiles, -bw)
            if best_key is None or key < best_key:
                best, best_key = (1, 1, bh, bw), key
    return best


class Lowerer:
    def __init__(self, g: Graph, weights: Dict[str, List[np.ndarray]], precision: str = "bf16",
                 max_batch: int = 8, tc: bool = True, tc_strided: bool = True,
                 crop=None, mean=None, scale=None, keep_all: bool = False, packed_stem: bool = True,
                 stem_halo: bool = True, stem_unroll: bool = True, fuse_pool: bool = True, s2d_stem: bool = True,
                 balance_n: bool = True, pair_pool: bool = True, persist_input: bool = False,
                 fuse_siblings: bool = True, s2d_depth: bool = True, input_dtypes=None, split_k: bool = True,
                 pair_halo: bool = True, stem_role: Optional[str] = None, stem_peer=None):
        # Two members of one ensemble read the same clips: their 7x7x7 stems can run as ONE GEMM with N = 2 x 64
        # (DeviceEnsemble pairs members 2k / 2k+1).  "lead": this member's stem op carries the peer's stem weights
        # (`stem_peer`, a weights dict) as columns 64..127 and stores them into a persistent peer buffer; "follow":
        # this member has no stem op, its first consumers read that peer buffer (written by the leader just before).
        if stem_role not in (None, "lead", "follow") or (stem_role == "lead") != (stem_peer is not None):
            raise ValueError("stem_role is None, 'lead' (with stem_peer weights) or 'follow'")
        self.stem_role, self.stem_peer = stem_role, stem_peer
        self.persistent: List[Buf] = []
        self.s2d_depth = s2d_depth
        self.pair_halo = pair_halo
        self.split_k = split_k
        # dtype of every external input: "u8" (decoded frames - the default) or "f32" (the dense flow of the
        # FarneBack_onTheFly TwoStream variant, train.py:294-332, evaluate_ensemble.py:1365-1386)
        self.input_dtypes = tuple(input_dtypes) if input_dtypes else tuple("u8" for _ in g.inputs)
        if len(self.input_dtypes) != len(g.inputs) or any(d not in ("u8", "f32") for d in self.input_dtypes):
            raise ValueError("input_dtypes must hold 'u8' / 'f32' for each of the %d inputs" % len(g.inputs))
        self.keep_all = keep_all
        self.fuse_siblings = fuse_siblings
        self.persist_input = persist_input
        self.pair_pool = pair_pool
        self.balance_n = balance_n
        self.s2d_stem = s2d_stem
        self.fuse_pool = fuse_pool and precision == "bf16" and tc
        self.stem_unroll = stem_unroll
        self.stem_halo = stem_halo
        self.packed_stem = packed_stem
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        check_weights(g, weights)
        self.g, self.w = g, weights
        self.precision = precision
        self.act = rt.BF16 if precision == "bf16" else rt.F32
        self.nb = max_batch
        self.use_tc = tc and precision == "bf16"
        self.tc_strided = tc_strided
        self.crop, self.mean, self.scale = crop, mean, scale
        self.ops: List[DevOp] = []
        self.bufs: List[Buf] = []
        self.blobs: List[np.ndarray] = []
        self.val: Dict[str, object] = {}
        self.consumers: Dict[str, List[str]] = {n: [] for n in g.nodes}
        for n in g.nodes.values():
            for i in n.inputs:
                self.consumers[i].append(n.name)
        self.done = set()
        self.place: Dict[str, Tuple[str, int]] = {}
        self.concat_ref: Dict[str, TRef] = {}
        for n in g.nodes.values():
            if n.op == "concat" and len(n.out_shape) == 4:
                off = 0
                for i in n.inputs:
                    if len(self.consumers[i]) == 1:
                        self.place[i] = (n.name, off)
                    off += g.shape(i)[-1]

    # ---- allocation ------------------------------------------------------- #
    def esize(self, dt):
        return 4 if dt == rt.F32 else 2

    def new_buf(self, name, dims, ld, dt) -> Buf:
        nbytes = self.nb * dims[0] * dims[1] * dims[2] * ld * self.esize(dt)
        b = Buf(name, _round_up(max(nbytes, 16), ALIGN))
        self.bufs.append(b)
        return b

    def out_ref(self, final_name: str, dims, C, dt) -> TRef:
        """Output view for the chain ending in `final_name`: a concat slice if one was planned."""
        if final_name in self.place and dt == self.act:
            cname, coff = self.place[final_name]
            if cname not in self.concat_ref:
                ctot = self.g.shape(cname)[-1]
                b = self.new_buf(cname, dims, ctot, dt)
                self.concat_ref[cname] = TRef(b, 0, ctot, ctot, tuple(dims), dt)
            base = self.concat_ref[cname]
            return TRef(base.buf, coff, C, base.ld, tuple(dims), dt)
        b = self.new_buf(final_name, dims, C, dt)
        return TRef(b, 0, C, C, tuple(dims), dt)

    def blob(self, arr: np.ndarray) -> int:
        self.blobs.append(np.ascontiguousarray(arr))
        return len(self.blobs) - 1

    def fblob(self, arr) -> int:
        return -1 if arr is None else self.blob(np.asarray(arr, np.float32))

    def emit(self, op: DevOp):
        self.ops.append(op)

    def sole_consumer(self, name: str, kind: str) -> Optional[Node]:
        c = self.consumers[name]
        if len(c) == 1 and self.g.nodes[c[0]].op == kind:
            return self.g.nodes[c[0]]
        return None

    # ---- node handlers ------------------------------------------------------ #
    def lower(self) -> Plan:
        g = self.g
        for node in g.nodes.values():
            if node.name in self.done:
                continue
            getattr(self, "_" + node.op)(node)
        logits, probs = self.val.get("__logits__"), self.val.get("__probs__")
        # liveness
        for idx, op in enumerate(self.ops):
            for r in (op.in0, op.in1, op.out0, op.out1, op.out2):
                if r is not None:
                    r.buf.first = min(r.buf.first, idx)
                    r.buf.last = max(r.buf.last, idx)
            if op.part is not None:
                op.part.first = op.part.last = idx
        for r in (logits, probs):
            if r is not None:
                r.buf.last = len(self.ops) + 1
        if self.persist_input:      # pre-processed clips stay valid for the next member of the ensemble
            for op in self.ops:
                if op.kind == rt.OP_PREPROCESS:
                    op.out0.buf.last = len(self.ops) + 1
                    op.out0.buf.first = -2 if self.stem_role else 0
        # peer stem buffers: placed right after the pre-processed clips (which every role of one ensemble places first,
        # in the same order), before any plan-local buffer - the same offset in the leader's and the follower's plan
        for b in self.persistent:
            b.first, b.last = -1, len(self.ops) + 1
        if self.keep_all:           # tests: every intermediate stays readable after the run
            for b in self.bufs:
                if b.last >= 0:
                    b.last = len(self.ops) + 1
        ws = self._assign_offsets()
        # weight arena
        offs, cur = [], 0
        for b in self.blobs:
            cur = _round_up(cur, 256)
            offs.append(cur)
            cur += b.nbytes
        arena = np.zeros(_round_up(max(cur, 256), 256), np.uint8)
        for b, o in zip(self.blobs, offs):
            arena[o:o + b.nbytes] = b.view(np.uint8).reshape(-1)
        tensors = {k: v for k, v in self.val.items() if isinstance(v, TRef) and not k.startswith("__")}
        return Plan(self.ops, self.bufs, ws, arena, offs, logits, probs, self.nb,
                    (g.shape(g.output)[-1] if g.output else 1), self.precision, len(g.inputs), tensors)

    def _assign_offsets(self) -> int:
        live = [b for b in self.bufs if b.last >= 0]
        live.sort(key=lambda b: (b.first, -b.nbytes))
        placed: List[Buf] = []
        top = 0
        for b in live:
            busy = sorted(((p.offset, p.offset + p.nbytes) for p in placed
                           if not (p.last < b.first or p.first > b.last)))
            off = 0
            for lo, hi in busy:
                if off + b.nbytes <= lo:
                    break
                off = max(off, hi)
            b.offset = off
            placed.append(b)
            top = max(top, off + b.nbytes)
        return _round_up(top, ALIGN)

    def _input(self, node: Node):
        t, h, w, c = node.out_shape
        idx = self.g.inputs.index(node.name)
        src_dt = rt.F32 if self.input_dtypes[idx] == "f32" else rt.U8
        crop = self.crop
        if crop is not None:
            t0, h0, w0, to, ho, wo = crop
            raise NotImplementedError("crop changes the graph's input shape; build the graph for the cropped shape")
        ld = c if self.act == rt.F32 else 8
        wpitch = wpad = 0
        cons = self.consumers[node.name]
        first = self.g.nodes[cons[0]] if len(cons) == 1 else None
        if (self.use_tc and self.packed_stem and src_dt == rt.U8 and first is not None and first.op == "conv3d" and c <= 8
                and first.attrs["k"] == (3, 3, 3) and first.attrs["s"] == (1, 1, 1)
                and first.attrs["padding"] == "same" and first.attrs["filters"] % 8 == 0 and c <= 4):
            # packed stem: the 3 kw taps of a pixel become one contiguous 32-wide K chunk
            # Either materialised by the pre-processing kernel ([.., W, 16]: 3 pixels x C channels
            # tightly packed, K = 16 per (fd,fh) tap) or read through an overlapping-stride TMA view of
            # the W-padded [.., W+4, 8] tensor (3 + 1 zero-weighted 8-channel pixels, K = 32 per tap).
            if self.stem_unroll and self._pair_pool_ok(first, (t, h, w, c)):
                # pair-packed stem: one element per pixel PAIR (2p, 2p+1) carrying pixels 2p-1..2p+2
                b = self.new_buf(node.name, (t, h, w // 2), 16, self.act)
                out = TRef(b, 0, c, 16, (t, h, w // 2), self.act, 0, 0, 4)
                mean = tuple(self.mean) + (0.0,) * (4 - len(self.mean)) if self.mean is not None else (0.0,) * 4
                scale = tuple(self.scale) + (1.0,) * (4 - len(self.scale)) if self.scale is not None else (1.0,) * 4
                self.emit(DevOp(rt.OP_PREPROCESS, node.name, None, None, out, ext_input=idx, src_dtype=src_dt,
                                src_dims=(t, h, w, c), pre_mean=mean, pre_scale=scale, layers=(node.name,)))
                self.val[node.name] = out
                return
            if self.stem_unroll:
                b = self.new_buf(node.name, (t, h, w), 16, self.act)
                out = TRef(b, 0, c, 16, (t, h, w), self.act, 0, 0, 3)
                mean = tuple(self.mean) + (0.0,) * (4 - len(self.mean)) if self.mean is not None else (0.0,) * 4
                scale = tuple(self.scale) + (1.0,) * (4 - len(self.scale)) if self.scale is not None else (1.0,) * 4
                self.emit(DevOp(rt.OP_PREPROCESS, node.name, None, None, out, ext_input=idx, src_dtype=src_dt,
                                src_dims=(t, h, w, c), pre_mean=mean, pre_scale=scale, layers=(node.name,)))
                self.val[node.name] = out
                return
            wpad, wpitch = 1, w + 4
        if (self.use_tc and self.s2d_stem and first is not None and first.op == "conv3d" and c <= 4
                and first.attrs["k"] == (7, 7, 7) and first.attrs["s"] == (2, 2, 2)
                and first.attrs["padding"] == "same" and first.attrs["filters"] % 8 == 0
                and first.attrs["filters"] >= 16 and (src_dt == rt.U8 or c <= 2)):
            # stride-2 7x7x7 stem (I3D Conv3d_1a_7x7 train.py:1026, R3D stem :1481): the pre-processing
            # kernel writes 2x2 space-to-depth cells over (H, W); rows are padded on the left by the
            # number of cells the 'same' padding reaches into and on the right so that the 4-cell
            # window of the last output pixel stays inside the row.
            h2, w2 = (h + 1) // 2, (w + 1) // 2
            cell = _round_up(4 * c, 8)
            pb_w = first.attrs["pads_before"][2]
            wpad = (pb_w + 1) // 2
            wpitch = w2 + 3
            mean = tuple(self.mean) + (0.0,) * (4 - len(self.mean)) if self.mean is not None else (0.0,) * 4
            scale = tuple(self.scale) + (1.0,) * (4 - len(self.scale)) if self.scale is not None else (1.0,) * 4
            # 2x2x2 cells (depth too) carry 8*C channels without padding: the stem becomes a stride-1 4x4x4-cell conv with
            # K = 16 x (4 cells x 8C) instead of 28 x (4 cells x round_up(4C, 8)).  Measured (tools/stem_bench.py,
            # profiles/r2_stem_variants.txt): the K chunk decides - 64-channel chunks (128-byte swizzle) run ~1.5x the
            # executed FLOP rate of 32-channel ones.  C = 2 (flow): window 4 x 16 = 64 channels, K 1024 at kc = 64 beats
            # K 896 at kc = 32 (1.52 vs 1.82 ms); C = 1: 512 vs 896, both kc = 32; C = 3: the 96-channel window needs
            # kc = 32 and loses to the 2-D layout's K 1792 at kc = 64 (3.0 vs 2.5 ms) - kept 2-D.
            want3d = c in (1, 2) if self.s2d_depth is True else self.s2d_depth == "always"
            if self.stem_halo and want3d:
                t2 = (t + 1) // 2
                cell3 = 8 * c
                b = self.new_buf(node.name, (t2, h2, wpitch), cell3, self.act)
                out = TRef(b, 0, cell3, cell3, (t2, h2, w2), self.act, wpitch, wpad, 0, 2, c)
            else:
                b = self.new_buf(node.name, (t, h2, wpitch), cell, self.act)
                out = TRef(b, 0, cell, cell, (t, h2, w2), self.act, wpitch, wpad, 0, 1, c)
            self.emit(DevOp(rt.OP_PREPROCESS, node.name, None, None, out, ext_input=idx, src_dtype=src_dt,
                            src_dims=(t, h, w, c), pre_mean=mean, pre_scale=scale, layers=(node.name,)))
            self.val[node.name] = out
            return
        b = self.new_buf(node.name, (t, h, wpitch or w), ld, self.act)
        out = TRef(b, 0, c, ld, (t, h, w), self.act, wpitch, wpad)
        mean = tuple(self.mean) + (0.0,) * (4 - len(self.mean)) if self.mean is not None else (0.0,) * 4
        scale = tuple(self.scale) + (1.0,) * (4 - len(self.scale)) if self.scale is not None else (1.0,) * 4
        self.emit(DevOp(rt.OP_PREPROCESS, node.name, None, None, out, ext_input=idx, src_dtype=src_dt,
                        src_dims=(t, h, w, c), pre_mean=mean, pre_scale=scale, layers=(node.name,)))
        self.val[node.name] = out

    def _conv_like(self, name, x: TRef, kernel, bias, k, s, pads, out_dims, chain_bn, relu, final_name,
                   layers, out_dtype=None, residual: Optional[TRef] = None, flops=0.0,
                   second: Optional[Tuple[np.ndarray, np.ndarray, int, str]] = None, halo: bool = False,
                   pool=None):
        """Emit one CONV3D op.  chain_bn = (bn_weights, has_gamma) or None.
        pool = (window, pooled_dims, pad_is_zero): MaxPooling3D fused into the tcgen05 epilogue."""
        co = kernel.shape[-1]
        out_dtype = self.act if out_dtype is None else out_dtype
        scale, shift = fold_bn(bias, chain_bn[0] if chain_bn else None, chain_bn[1] if chain_bn else False, co)
        out0 = self.out_ref(final_name, pool[1] if pool else out_dims, co, out_dtype)
        op = DevOp(rt.OP_CONV3D, name, x, residual, out0, k=tuple(k), s=tuple(s), pad=tuple(pads),
                   relu0=int(relu), layers=tuple(layers), flops=flops)
        mult = (1, 1, 1)
        if pool:
            op.pool_k, op.pool_zero, op.conv_out_dims = tuple(pool[0]), int(pool[2]), tuple(out_dims)
            mult = tuple(pool[0])
        op.scale0, op.shift0 = self.fblob(scale), self.fblob(shift)
        ci = x.C
        tc_ok = self._tc_ok(x, co, s, out_dtype, residual) and out0.ld % 8 == 0 and out0.coff % 8 == 0
        if pool and not tc_ok:
            raise RuntimeError("fused pooling was planned for a conv that cannot use the tcgen05 engine")
        if tc_ok:
            kc = choose_kc(ci)
            gen_brick = choose_brick(self.nb, *out_dims, mult=mult)
            m_tiles = (-(-self.nb // gen_brick[0]) * -(-out_dims[0] // gen_brick[1]) * -(-out_dims[1] // gen_brick[2])
                       * -(-out_dims[2] // gen_brick[3]))
            bn, n_tiles = choose_bn(co, m_tiles if self.balance_n else 0)
            op.engine, op.w_dtype, op.kc, op.bn = rt.ENGINE_TCGEN05, rt.BF16, kc, bn
            hw_brick = choose_brick_hw(out_dims[1], out_dims[2], mult) if halo is True or halo == 1 else None
            ph_brick, ph_eff = choose_brick_pair_halo(out_dims[1], out_dims[2], kernel.shape[1])
            if (self.pair_halo and not halo and not pool and kernel.shape[2] > 1 and tuple(s[1:]) == (1, 1) and kc == 64
                    and n_tiles == 1 and bn <= 128 and bn % 16 == 0 and (ph_eff >= 0.8 or self.pair_halo == "always")
                    and (self.pair_halo == "always" or
                         self.nb * out_dims[0] * -(-out_dims[1] // ph_brick[2]) * -(-out_dims[2] // ph_brick[3]) >= 2 * SM_COUNT)):
                # 3x3x3 / stride-1 conv with a small single N tile and 64-channel chunks (R3D stage 1, C3D conv2): in the
                # generic mode every tap re-loads its 16 KB A tile, and with N <= 128 that L2 -> shared-memory stream,
                # not the MMA, sets the pace (R3D stage 1: 750 TFLOP/s).  CTA-pair h-halo mode: one haloed box per
                # (fd, fw) feeds the kh taps, each CTA loads half of the weight rows, MMAs are 256 x N x 16.
                op.halo = 3
                op.brick = ph_brick
                op.w_blob = self.blob(pack_tc_weights_hhalo(kernel, kc, bn, n_tiles))
            elif (halo == 2 and kernel.shape[2] == 1 and tuple(s[1:]) == (1, 1) and kernel.shape[1] * bn <= 256
                    and (bn * kc * 2) % 1024 == 0 and not pool):
                op.halo = 2
                op.brick = choose_brick_hhalo(out_dims[1], out_dims[2], kernel.shape[1])
                op.w_blob = self.blob(pack_tc_weights_hhalo(kernel, kc, bn, n_tiles))
            elif (hw_brick is not None and kernel.shape[2] == 1 and ci <= kc and kernel.shape[1] * bn <= 256
                    and (bn * kc * 2) % 1024 == 0):
                op.halo = 1
                op.brick = hw_brick
                op.w_blob = self.blob(pack_tc_weights_halo(kernel, kc, bn, n_tiles))
            else:
                op.brick = gen_brick
                op.w_blob = self.blob(pack_tc_weights(kernel, kc, bn, n_tiles))
                # split-K: a layer with far fewer tiles than SMs (small batches: C3D conv5a/5b at batch 8; the Dense
                # layers, whose M is the batch) spreads the K loop of every tile over several CTAs; the fp32 partial
                # tiles are summed in a fixed order by a second kernel.  Decided here, from max_batch - never from the
                # batch actually run - so one member gives bit-identical results for every batch size.
                ks = 1 if pool else self._split_k_factor(ci, kernel, out_dims)
                if ks >= 2:
                    op.ksplit = ks
                    nbytes = ks * m_tiles * 128 * n_tiles * bn * 4
                    op.part = Buf(name + ":splitk", _round_up(nbytes, ALIGN))
                    self.bufs.append(op.part)
        else:
            op.engine = rt.ENGINE_DIRECT
            kflat = kernel.reshape(-1, co)
            if x.dtype == rt.BF16:
                op.w_dtype = rt.BF16
                op.w_blob = self.blob(to_bf16_bits(kflat))
            else:
                op.w_dtype = rt.F32
                op.w_blob = self.blob(kflat.astype(np.float32))
        if second is not None:
            sc1, sh1, relu1, second_name = second
            op.out1 = self.out_ref(second_name, out_dims, co, out_dtype)
            op.scale1, op.shift1, op.relu1 = self.fblob(sc1), self.fblob(sh1), int(relu1)
        self.emit(op)
        return op

    def _split_k_factor(self, ci: int, kernel, out_dims, mult=(1, 1, 1)) -> int:
        """Split-K factor the generic tcgen05 path would use for this conv (1 = no split); see _conv_like."""
        if not (self.split_k and self.use_tc):
            return 1
        kc = choose_kc(ci)
        brick = choose_brick(self.nb, *out_dims, mult=mult)
        m_tiles = (-(-self.nb // brick[0]) * -(-out_dims[0] // brick[1]) * -(-out_dims[1] // brick[2])
                   * -(-out_dims[2] // brick[3]))
        bn, n_tiles = choose_bn(kernel.shape[-1], m_tiles if self.balance_n else 0)
        ksteps = kernel.shape[0] * kernel.shape[1] * kernel.shape[2] * -(-ci // kc)
        tiles = m_tiles * n_tiles
        if tiles * 2 > SM_COUNT or ksteps < 16:
            return 1
        ks = min(SM_COUNT // tiles, ksteps // 8, 8)
        return ks if ks >= 2 else 1

    def _tc_ok(self, x: TRef, co: int, s, out_dtype: int, residual: Optional[TRef]) -> bool:
        return (self.use_tc and x.dtype == rt.BF16 and out_dtype == rt.BF16 and x.C % 8 == 0 and x.ld % 8 == 0
                and x.coff % 8 == 0 and co % 8 == 0 and co >= 16
                and (self.tc_strided or tuple(s) == (1, 1, 1))
                and (residual is None or (residual.ld % 8 == 0 and residual.coff % 8 == 0)))

    def _fusable_pool(self, final: str, out_dims):
        """MaxPooling3D (window == stride, 'valid', optionally behind a ZeroPadding3D that only pads at
        the end) that is the sole consumer of `final` -> (window, pooled dims, pad_is_zero, layer names)."""
        if not self.fuse_pool:
            return None
        names, zero, cur, dims = [], False, final, tuple(out_dims)
        zp = self.sole_consumer(cur, "zeropad")
        if zp is not None:
            pads = zp.attrs["pads"]
            if any(p[0] != 0 for p in pads):
                return None
            names.append(zp.name)
            zero, cur = True, zp.name
            dims = zp.out_shape[:3]
        mp = self.sole_consumer(cur, "maxpool")
        if mp is None or mp.attrs["padding"] != "valid" or tuple(mp.attrs["k"]) != tuple(mp.attrs["s"]):
            return None
        k = tuple(mp.attrs["k"])
        if any(128 % kk for kk in k) or k[0] * k[1] * k[2] > 16:
            return None
        if mp.name in self.place or (zp is not None and zp.name in self.place):
            return None
        return k, tuple(mp.out_shape[:3]), zero, names + [mp.name]

    def _conv_chain(self, node: Node):
        """Conv3D (+ BatchNormalization) (+ ReLU) chain folded into one op -> (chain_bn, relu, final, layers)."""
        layers = [node.name]
        final = node.name
        relu = node.attrs["activation"] == "relu"
        chain_bn = None
        nxt = self.sole_consumer(final, "bn")
        if nxt is not None and not relu:
            chain_bn = (self.w[nxt.name], nxt.attrs["scale"])
            layers.append(nxt.name)
            final = nxt.name
        nxt = self.sole_consumer(final, "relu")
        if nxt is not None and not relu:
            relu = True
            layers.append(nxt.name)
            final = nxt.name
        return chain_bn, relu, final, layers

    def _fused_siblings(self, node: Node) -> bool:
        """Horizontal fusion of the 1x1x1 convs that read the same tensor (branch 0, 1a, 2a of an Inception
        block, train.py:1048-1193): one GEMM with N = sum of their filters, so the block input is read
        once instead of three times.  The epilogue stores each column range through its own TMA map:
        branch 0 into its slice of the concat buffer, the 1a / 2a activations into dense buffers of
        their own (the 3x3x3 convs that follow re-read them 27 times, so they must stay contiguous)."""
        if not (self.use_tc and self.fuse_siblings):
            return False

        def is_pw(n):
            return (n.op == "conv3d" and tuple(n.attrs["k"]) == (1, 1, 1) and tuple(n.attrs["s"]) == (1, 1, 1)
                    and n.name not in self.done)
        if not is_pw(node):
            return False
        src = node.inputs[0]
        x = self.val[src]
        sibs = [self.g.nodes[c] for c in self.consumers[src] if is_pw(self.g.nodes[c])]
        if len(sibs) < 2 or node.name not in [n.name for n in sibs]:
            return False
        chains = [self._conv_chain(n) for n in sibs]
        placed = [i for i, ch in enumerate(chains) if ch[2] in self.place]
        if len(placed) != 1 or len(sibs) > 3 or len({ch[1] for ch in chains}) != 1:
            return False
        order = placed + [i for i in range(len(sibs)) if i != placed[0]]
        sibs, chains = [sibs[i] for i in order], [chains[i] for i in order]
        cos = [n.attrs["filters"] for n in sibs]
        if cos[0] % 16 or (len(cos) == 3 and cos[1] % 16) or any(c % 8 for c in cos):
            return False
        if any(ch[2] in self.place for ch in chains[1:]):
            return False
        co = sum(cos)
        if not self._tc_ok(x, co, (1, 1, 1), self.act, None):
            return False
        out_dims = tuple(node.out_shape[:3])
        outs = [self.out_ref(ch[2], out_dims, c, self.act) for ch, c in zip(chains, cos)]
        if any(o.ld % 8 or o.coff % 8 for o in outs):
            return False
        kernel = np.concatenate([self.w[n.name][0] for n in sibs], axis=-1)
        scales, shifts = [], []
        for n, ch, c in zip(sibs, chains, cos):
            bias = self.w[n.name][1] if n.attrs["use_bias"] else None
            sc, sh = fold_bn(bias, ch[0][0] if ch[0] else None, ch[0][1] if ch[0] else False, c)
            scales.append(np.ones(c, np.float32) if sc is None else sc)
            shifts.append(np.zeros(c, np.float32) if sh is None else sh)
        fl = self.g.conv_dense_flops()
        layers = tuple(l for ch in chains for l in ch[3])
        o0 = outs[0]
        op = DevOp(rt.OP_CONV3D, "+".join(n.name for n in sibs), x, None,
                   TRef(o0.buf, o0.coff, co, o0.ld, out_dims, self.act), outs[1], k=(1, 1, 1), s=(1, 1, 1), pad=(0, 0, 0),
                   relu0=int(chains[0][1]), layers=layers, flops=sum(fl[n.name] for n in sibs))
        op.scale0, op.shift0 = self.fblob(np.concatenate(scales)), self.fblob(np.concatenate(shifts))
        op.out_split = cos[0]
        if len(cos) == 3:
            op.out_split2, op.out2 = cos[0] + cos[1], outs[2]
        kc = choose_kc(x.C)
        op.brick = choose_brick(self.nb, *out_dims)
        m_tiles = (-(-self.nb // op.brick[0]) * -(-out_dims[0] // op.brick[1]) * -(-out_dims[1] // op.brick[2])
                   * -(-out_dims[2] // op.brick[3]))
        bn, n_tiles = choose_bn(co, m_tiles if self.balance_n else 0)
        # a TMA store chunk (64 / 32 / 16 channels) must divide the N tile and both splits: widen the
        # tile to the next multiple of 64 or 32 when that costs < 15 % padded columns (wider stores win)
        for gran in (64, 32):
            wide = _round_up(bn, gran)
            if (op.out_split % gran == 0 and op.out_split2 % gran == 0 and wide <= 256
                    and wide * n_tiles <= 1.15 * co):
                bn = wide
                break
        op.engine, op.w_dtype, op.kc, op.bn = rt.ENGINE_TCGEN05, rt.BF16, kc, bn
        op.w_blob = self.blob(pack_tc_weights(kernel, kc, bn, n_tiles))
        self.emit(op)
        for ch, ref in zip(chains, outs):
            for l in ch[3]:
                self.val[l] = ref
                self.done.add(l)
        return True

    def _conv3d(self, node: Node):
        g = self.g
        if self._fused_siblings(node):
            return
        x = self.val[node.inputs[0]]
        ws = self.w[node.name]
        kernel = ws[0]
        bias = ws[1] if node.attrs["use_bias"] else None
        chain_bn, relu, final, layers = self._conv_chain(node)
        out_dims = node.out_shape[:3]
        flops = g.conv_dense_flops()[node.name]
        pool = None
        if x.s2d:
            self._s2d_stem_conv(node, x, kernel, bias, chain_bn, relu, final, layers, out_dims, flops)
            return
        if x.unroll_w == 4:
            self._pair_pool_stem(node, x, kernel, bias, relu, final, layers, out_dims, flops)
            return
        if x.wpitch or x.unroll_w or self._tc_ok(x, kernel.shape[-1], node.attrs["s"], self.act, None):
            fp = self._fusable_pool(final, out_dims)
            if fp is not None and not (x.wpitch or x.unroll_w) and self._split_k_factor(x.C, kernel, out_dims) > 1:
                fp = None       # a layer this small gains more from split-K than from the fused pool (tiny stand-alone pool)
            if fp is not None and final not in self.place:
                pool = fp[:3]
                layers = layers + fp[3]
                final = fp[3][-1]
        if x.wpitch or x.unroll_w:
            self._packed_stem_conv(node, x, kernel, bias, chain_bn, relu, final, layers, out_dims, flops, pool)
            return
        # residual fusion: conv (no bn/relu tail) whose only consumer is add([shortcut, this])
        add = self.sole_consumer(final, "add") if (chain_bn is None and not relu) else None
        if add is not None and len(add.inputs) == 2 and add.inputs[1] == final and add.inputs[0] not in self.val:
            # projection shortcut (_shortcut3d, train.py:1338): created after the residual convs
            # in the graph but only depends on the block input -> lower it first
            sc_node = g.nodes[add.inputs[0]]
            if sc_node.op == "conv3d" and sc_node.inputs[0] in self.val:
                self._conv3d(sc_node)
        if add is not None and len(add.inputs) == 2 and add.inputs[1] == final and add.inputs[0] in self.val:
            self._fused_residual(node, add, x, kernel, bias, out_dims, flops)
            return
        self._conv_like(node.name, x, kernel, bias, node.attrs["k"], node.attrs["s"], node.attrs["pads_before"],
                        out_dims, chain_bn, relu, final, layers, flops=flops, pool=pool)
        ref = self.ops[-1].out0
        for l in layers:
            self.val[l] = ref
            self.done.add(l)

    def _packed_stem_conv(self, node, x, kernel, bias, chain_bn, relu, final, layers, out_dims, flops, pool=None):
        """3x3x3 'same' stem on a C<=8 input (C3D conv1, train.py:1230): the kw taps are folded into
        the channel axis.  The input rows carry one zero column on the left (x.wpad) and >= 2 on the
        right, channels padded to 8, so the window (w-1 .. w+2) x 8 channels of output pixel w is
        32 contiguous bf16 values starting at padded column w: an overlapping-stride TMA view
        [.., W, 32] with pixel stride 8.  The conv becomes k=(3,3,1), Cin=32 with zero weights for
        the 4th pixel and the padded channels."""
        kd, kh, kw, ci, co = kernel.shape
        assert (kd, kh, kw) == (3, 3, 3)
        if x.unroll_w:
            assert 3 * ci <= 16
            k2 = np.zeros((3, 3, 1, 16, co), np.float32)
            for j in range(3):
                k2[:, :, 0, j * ci:(j + 1) * ci, :] = kernel[:, :, j, :, :]
            view = TRef(x.buf, 0, 16, 16, x.dims, x.dtype)
        else:
            k2 = np.zeros((3, 3, 1, 32, co), np.float32)
            for j in range(3):
                k2[:, :, 0, j * 8:j * 8 + ci, :] = kernel[:, :, j, :, :]
            assert x.wpad == 1 and x.ld == 8
            view = TRef(x.buf, 0, 32, 8, x.dims, x.dtype, x.wpitch, x.wpad)
        op = self._conv_like(node.name, view, k2, bias, (3, 3, 1), (1, 1, 1), (1, 1, 0), out_dims, chain_bn, relu,
                             final, layers, flops=flops, halo=self.stem_halo, pool=pool)
        if op.engine != rt.ENGINE_TCGEN05:
            raise RuntimeError("packed stem must lower to the tcgen05 engine")
        for l in layers:
            self.val[l] = op.out0
            self.done.add(l)

    def _pair_pool_ok(self, conv: Node, in_shape) -> bool:
        """C3D-style stem: Conv3D 3x3x3 'same' + bias (+ReLU) on a C<=4 clip whose only consumer is a
        MaxPooling3D (1,2,2)/(1,2,2) 'valid' (train.py:1230-1233)."""
        t, h, w, c = in_shape
        if not (self.pair_pool and self.fuse_pool and self.stem_halo and 4 * c <= 16 and w % 2 == 0 and h % 2 == 0):
            return False
        if conv.attrs["filters"] != 64 or not conv.attrs["use_bias"] or conv.name in self.place:
            return False
        final = conv.name
        if conv.attrs["activation"] not in (None, "relu", "linear"):
            return False
        if conv.attrs["activation"] != "relu":
            nxt = self.sole_consumer(final, "relu")
            if nxt is not None:
                final = nxt.name
        if final in self.place:
            return False
        mp = self.sole_consumer(final, "maxpool")
        return (mp is not None and mp.attrs["padding"] == "valid" and tuple(mp.attrs["k"]) == (1, 2, 2)
                and tuple(mp.attrs["s"]) == (1, 2, 2) and mp.name not in self.place)

    def _pair_pool_stem(self, node, x, kernel, bias, relu, final, layers, out_dims, flops):
        """C3D conv1 + pool1 (train.py:1230-1233) as ONE tcgen05 op on the pair-unrolled clip: GEMM row =
        the pixel pair (2p, 2p+1), N = 2 x 64 output channels, K = 9 (kd,kh) taps x 16 (4 neighbouring
        pixels x C channels); the 3 kw taps of the left pixel read neighbours 0..2, those of the right
        pixel neighbours 1..3.  The (1,2,2) max-pool is done in registers by the epilogue."""
        kd, kh, kw, ci, co = kernel.shape
        assert (kd, kh, kw) == (3, 3, 3) and co == 64 and 4 * ci <= 16
        mp = self.sole_consumer(final, "maxpool")
        t, h, wp = x.dims
        k2 = np.zeros((3, 3, 1, 16, 2 * co), np.float32)
        for px in range(2):
            for fw in range(3):
                k2[:, :, 0, (px + fw) * ci:(px + fw + 1) * ci, px * co:(px + 1) * co] = kernel[:, :, fw, :, :]
        pooled_dims = tuple(mp.out_shape[:3])
        assert pooled_dims == (t, h // 2, wp)
        out0 = self.out_ref(mp.name, pooled_dims, co, self.act)
        if out0.ld % 8 or out0.coff % 8:
            raise RuntimeError("pair-pool stem output must be 16-byte aligned")
        view = TRef(x.buf, 0, 16, 16, x.dims, x.dtype)
        op = DevOp(rt.OP_CONV3D, node.name, view, None, out0, k=(3, 3, 1), s=(1, 1, 1), pad=(1, 1, 0), relu0=int(relu),
                   layers=tuple(layers) + (mp.name,), flops=flops)
        op.engine, op.w_dtype, op.kc, op.bn, op.halo, op.pair_pool = rt.ENGINE_TCGEN05, rt.BF16, 16, 2 * co, 1, 1
        op.pool_k, op.pool_zero, op.conv_out_dims = (1, 2, 1), 0, (t, h, wp)
        bricks = [(1, 1, 128 // bw, bw) for bw in (8, 16)]
        op.brick = min(bricks, key=lambda b: (-(-h // b[2]) * -(-wp // b[3]), b[3]))
        op.w_blob = self.blob(pack_tc_weights_halo(k2, 16, 2 * co, 1))
        op.shift0 = self.fblob(bias.astype(np.float32))
        self.emit(op)
        for l in op.layers:
            self.val[l] = out0
            self.done.add(l)

    def _s2d_stem_conv(self, node, x, kernel, bias, chain_bn, relu, final, layers, out_dims, flops):
        """7x7x7 / stride 2 'same' stem on a C<=4 clip (I3D Conv3d_1a_7x7 train.py:1026; R3D stem
        :1481) as a tcgen05 implicit GEMM.  The input was written as 2x2 space-to-depth cells over
        (H, W) (see _input); output pixel (h, w) needs input rows 2h-pb .. 2h-pb+6, i.e. 4 cell rows
        starting at cell h - ceil(pb/2) (one of the 8 covered rows gets a zero weight), and likewise 4
        cells along W, which are contiguous in memory: an overlapping-stride TMA view [.., W/2, 4*cell]
        with pixel stride `cell`.  The conv becomes k=(7,4,1), stride (2,1,1), Cin = 4*cell; the
        regrouped kernel has exact zeros at the padded taps / channels, so the result is the same sum
        of products."""
        kd, kh, kw, ci, co = kernel.shape
        assert (kd, kh, kw) == (7, 7, 7) and ci == x.src_c
        cell = x.ld
        pb = node.attrs["pads_before"]
        off_d, off_h, off_w = (p - 2 * ((p + 1) // 2) for p in pb)                 # 0 (pad even) or -1 (odd)
        depth = x.s2d == 2          # 2x2x2 cells: depth is regrouped like H and W (4 cell taps, stride 1)

        def regroup(kern):
            k2 = np.zeros((4 if depth else 7, 4, 1, 4 * cell, co), np.float32)
            for fd in range(k2.shape[0]):
                for pd in range(2 if depth else 1):
                    td = 2 * fd + pd + off_d if depth else fd
                    if not 0 <= td < 7:
                        continue
                    for fh in range(4):
                        for ph in range(2):
                            th = 2 * fh + ph + off_h
                            if not 0 <= th < 7:
                                continue
                            for fw in range(4):
                                for pw in range(2):
                                    tw = 2 * fw + pw + off_w
                                    if not 0 <= tw < 7:
                                        continue
                                    c0 = fw * cell + ((pd * 2 + ph) * 2 + pw) * ci
                                    k2[fd, fh, 0, c0:c0 + ci, :] = kern[td, th, tw, :, :]
            return k2

        view = TRef(x.buf, 0, 4 * cell, cell, x.dims, x.dtype, x.wpitch, x.wpad)
        if depth:
            k_, s_, pad_ = (4, 4, 1), (1, 1, 1), ((pb[0] + 1) // 2, (pb[1] + 1) // 2, 0)
        else:
            k_, s_, pad_ = (7, 4, 1), (2, 1, 1), (pb[0], (pb[1] + 1) // 2, 0)
        if self.stem_role is not None:
            if not self.stem_fusable(self.g, self.precision) or not self.use_tc or not self.stem_halo:
                raise RuntimeError("stem_role needs the bf16 tcgen05 path and a 64-filter 7x7x7 stem")
            # the peer buffer exists in both roles and is placed before every plan-local buffer (lower()), so that the
            # leader's store and the follower's loads agree on its address inside the shared workspace
            pbuf = self.new_buf(final + ":peer", out_dims, co, self.act)
            self.persistent.append(pbuf)
            peer = TRef(pbuf, 0, co, co, tuple(out_dims), self.act)
            if self.stem_role == "follow":
                for l in layers:
                    self.val[l] = peer
                    self.done.add(l)
                return
            pw_ = self.stem_peer
            pkernel = pw_[node.name][0]
            pbias = pw_[node.name][1] if node.attrs["use_bias"] else None
            bn_name = layers[1] if chain_bn is not None else None
            folds = [fold_bn(bias, chain_bn[0] if chain_bn else None, chain_bn[1] if chain_bn else False, co),
                     fold_bn(pbias, pw_[bn_name] if bn_name else None, chain_bn[1] if chain_bn else False, co)]
            scale = np.concatenate([np.ones(co, np.float32) if f[0] is None else f[0] for f in folds])
            shift = np.concatenate([np.zeros(co, np.float32) if f[1] is None else f[1] for f in folds])
            kcat = np.concatenate([regroup(kernel), regroup(pkernel)], axis=-1)
            o0 = self.out_ref(final, out_dims, co, self.act)
            op = DevOp(rt.OP_CONV3D, node.name + "+peer", view, None, TRef(o0.buf, o0.coff, 2 * co, o0.ld, tuple(out_dims), self.act),
                       peer, k=k_, s=s_, pad=pad_, relu0=int(relu), layers=tuple(layers), flops=2 * flops)
            op.scale0, op.shift0 = self.fblob(scale), self.fblob(shift)
            op.out_split = co
            op.engine, op.w_dtype, op.kc, op.bn, op.halo = rt.ENGINE_TCGEN05, rt.BF16, choose_kc(view.C), 2 * co, 3
            op.brick = choose_brick_hhalo(out_dims[1], out_dims[2], 4)
            op.w_blob = self.blob(pack_tc_weights_hhalo(kcat, op.kc, op.bn, 1))
            self.emit(op)
            for l in layers:
                self.val[l] = o0
                self.done.add(l)
            return
        op = self._conv_like(node.name, view, regroup(kernel), bias, k_, s_, pad_, out_dims,
                             chain_bn, relu, final, layers, flops=flops, halo=2 if self.stem_halo else 0)
        if op.engine != rt.ENGINE_TCGEN05:
            raise RuntimeError("s2d stem must lower to the tcgen05 engine")
        for l in layers:
            self.val[l] = op.out0
            self.done.add(l)

    @staticmethod
    def stem_fusable(g: Graph, precision: str = "bf16") -> bool:
        """True when every clip input of the graph feeds a 7x7x7 / stride-2 conv with 64 filters (I3D, TwoStream-I3D,
        R3D): the stems DeviceEnsemble can run for two members at once."""
        if precision != "bf16" or not g.inputs:
            return False
        for name in g.inputs:
            cons = [n for n in g.nodes.values() if name in n.inputs]
            if len(cons) != 1 or cons[0].op != "conv3d":
                return False
            a = cons[0].attrs
            if tuple(a["k"]) != (7, 7, 7) or tuple(a["s"]) != (2, 2, 2) or a["filters"] != 64:
                return False
        return True

    def _fused_residual(self, node, add, x, kernel, bias, out_dims, flops):
        """residual conv + add (+ the next block's BN-ReLU as a second output)."""
        sc = self.val[add.inputs[0]]
        layers = [node.name, add.name]
        second = None
        cons = self.consumers[add.name]
        bn_nodes = [self.g.nodes[c] for c in cons if self.g.nodes[c].op == "bn"]
        second_layers = []
        if len(bn_nodes) == 1:
            bn = bn_nodes[0]
            r = self.sole_consumer(bn.name, "relu")
            sc1, sh1 = fold_bn(None, self.w[bn.name], bn.attrs["scale"], kernel.shape[-1])
            second_name = r.name if r is not None else bn.name
            second = (sc1, sh1, r is not None, second_name)
            second_layers = [bn.name] + ([r.name] if r is not None else [])
        op = self._conv_like(node.name, x, kernel, bias, node.attrs["k"], node.attrs["s"],
                             node.attrs["pads_before"], out_dims, None, False, add.name,
                             layers + second_layers, residual=sc, flops=flops, second=second)
        for l in layers:
            self.val[l] = op.out0
            self.done.add(l)
        for l in second_layers:
            self.val[l] = op.out1
            self.done.add(l)

    def _bn(self, node: Node):
        x = self.val[node.inputs[0]]
        scale, shift = fold_bn(None, self.w[node.name], node.attrs["scale"], x.C)
        layers, final, relu = [node.name], node.name, False
        nxt = self.sole_consumer(final, "relu")
        if nxt is not None:
            relu, final = True, nxt.name
            layers.append(final)
        out = self.out_ref(final, x.dims, x.C, x.dtype)
        op = DevOp(rt.OP_AFFINE, node.name, x, None, out, relu0=int(relu), layers=tuple(layers))
        op.scale0, op.shift0 = self.fblob(scale), self.fblob(shift)
        self.emit(op)
        for l in layers:
            self.val[l] = out
            self.done.add(l)

    def _relu(self, node: Node):
        x = self.val[node.inputs[0]]
        out = self.out_ref(node.name, x.dims, x.C, x.dtype)
        self.emit(DevOp(rt.OP_AFFINE, node.name, x, None, out, relu0=1, layers=(node.name,)))
        self.val[node.name] = out

    def _dropout(self, node: Node):
        self.val[node.name] = self.val[node.inputs[0]]     # identity at inference

    def _add(self, node: Node):
        a, b = (self.val[i] for i in node.inputs)
        out = self.out_ref(node.name, a.dims, a.C, a.dtype)
        self.emit(DevOp(rt.OP_ADD, node.name, a, b, out, layers=(node.name,)))
        self.val[node.name] = out

    def _concat(self, node: Node):
        if len(node.out_shape) == 1:
            self.val[node.name] = [self.val[i] for i in node.inputs]     # consumed by _dense
            return
        if node.name not in self.concat_ref or any(i not in self.place for i in node.inputs):
            raise NotImplementedError("concat %s: an input was not written in place" % node.name)
        self.val[node.name] = self.concat_ref[node.name]

    def _pool(self, node: Node, kind: int, x: TRef, pads, pad_is_zero, out_dims, layers):
        out = self.out_ref(node.name, out_dims, x.C, x.dtype)
        self.emit(DevOp(kind, node.name, x, None, out, k=node.attrs["k"], s=node.attrs["s"], pad=tuple(pads),
                        pad_is_zero=int(pad_is_zero), layers=tuple(layers)))
        for l in layers:
            self.val[l] = out
            self.done.add(l)

    def _maxpool(self, node: Node):
        self._pool(node, rt.OP_MAXPOOL3D, self.val[node.inputs[0]], node.attrs["pads_before"], False,
                   node.out_shape[:3], [node.name])

    def _avgpool(self, node: Node):
        self._pool(node, rt.OP_AVGPOOL3D, self.val[node.inputs[0]], (0, 0, 0), False, node.out_shape[:3],
                   [node.name])

    def _zeropad(self, node: Node):
        nxt = self.sole_consumer(node.name, "maxpool")
        if nxt is None or nxt.attrs["padding"] != "valid":
            raise NotImplementedError("ZeroPadding3D is only supported in front of a 'valid' MaxPooling3D")
        pads = node.attrs["pads"]
        self._pool(nxt, rt.OP_MAXPOOL3D, self.val[node.inputs[0]], tuple(p[0] for p in pads), True,
                   nxt.out_shape[:3], [node.name, nxt.name])

    def _flatten(self, node: Node):
        x = self.val[node.inputs[0]]
        if x.ld != x.C or x.coff != 0:
            raise NotImplementedError("flatten of a channel slice")
        n = x.dims[0] * x.dims[1] * x.dims[2] * x.C
        self.val[node.name] = TRef(x.buf, 0, n, n, (1, 1, 1), x.dtype)

    def _dense(self, node: Node):
        g = self.g
        kernel, bias = self.w[node.name]
        act = node.attrs["activation"]
        is_final = node.name == g.output
        units = node.attrs["units"]
        xin = self.val[node.inputs[0]]
        parts = xin if isinstance(xin, list) else [xin]
        out_dtype = rt.F32 if is_final else self.act
        flops_total = g.conv_dense_flops()[node.name]
        row = 0
        prev = None
        for pi, x in enumerate(parts):
            kpart = kernel[row:row + x.C].reshape(1, 1, 1, x.C, units)
            row += x.C
            last = pi == len(parts) - 1
            name = node.name if len(parts) == 1 else "%s#%d" % (node.name, pi)
            final_name = node.name if last else name
            self._conv_like(name, x, kpart, bias if pi == 0 else None, (1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1),
                            None, act == "relu" and last, final_name, [node.name], out_dtype=out_dtype,
                            residual=prev, flops=flops_total * x.C / kernel.shape[0])
            prev = self.ops[-1].out0
        assert row == kernel.shape[0]
        self.val[node.name] = prev
        if is_final:
            if act != "softmax":
                raise NotImplementedError("final activation %r" % act)
            pb = self.new_buf(node.name + ":probs", (1, 1, 1), units, rt.F32)
            probs = TRef(pb, 0, units, units, (1, 1, 1), rt.F32)
            self.emit(DevOp(rt.OP_SOFTMAX, node.name + ":softmax", prev, None, probs, softmax_C=units,
                            layers=(node.name,)))
            self.val["__logits__"], self.val["__probs__"] = prev, probs


def lower(g: Graph, weights, precision="bf16", max_batch=8, **kw) -> Plan:
    return Lowerer(g, weights, precision, max_batch, **kw).lower()
