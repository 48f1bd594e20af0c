"""ctypes binding of libcse_b200.so (include/cse.h).  PyTorch is used only to own device
memory and streams; every pointer handed to the library is a raw device address.

There is NO CPU fallback: if the shared library has not been built, or a compute entry point
is called without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcse_b200.so")

# enums (include/cse.h)
ABI_VERSION = 2
F32, BF16, U8 = 0, 1, 2
ENGINE_AUTO, ENGINE_DIRECT, ENGINE_TCGEN05 = 0, 1, 2
OP_PREPROCESS, OP_CONV3D, OP_MAXPOOL3D, OP_AVGPOOL3D, OP_AFFINE, OP_ADD, OP_SOFTMAX = 1, 2, 3, 4, 5, 6, 7
OP_NAMES = {1: "preprocess", 2: "conv3d", 3: "maxpool3d", 4: "avgpool3d", 5: "affine", 6: "add", 7: "softmax"}

EXPORTS = ["cse_abi_version", "cse_last_error", "cse_device_info", "cse_tune",
           "cse_malloc", "cse_free", "cse_memcpy_h2d", "cse_memcpy_d2h", "cse_stream_synchronize", "cse_plan_create", "cse_plan_add_op",
           "cse_plan_finalize", "cse_plan_run", "cse_plan_run_from", "cse_plan_num_input_ops", "cse_plan_run_range",
           "cse_plan_num_ops",
           "cse_plan_last_launches", "cse_plan_destroy", "cse_preprocess", "cse_vote", "cse_vote_search",
           "cse_assemble_clip", "cse_resize_u8", "cse_bgr2gray", "cse_farneback_workspace_bytes", "cse_farneback",
           "cse_resize_linear_f32",
           "cse_model_create", "cse_model_set_option", "cse_model_num_layers", "cse_model_layer_info", "cse_model_tensor_info",
           "cse_model_set_weight", "cse_model_pair_stems", "cse_model_lower", "cse_model_num_ops", "cse_model_get_op", "cse_model_workspace_bytes",
           "cse_model_weight_bytes", "cse_model_copy_weight_arena", "cse_model_logits_offset", "cse_model_probs_offset",
           "cse_model_finalize", "cse_model_forward", "cse_model_forward_shared_input", "cse_model_destroy"]


class CseOp(C.Structure):
    """Mirror of ``struct cse_op``."""
    _fields_ = [
        ("kind", C.c_int32), ("engine", C.c_int32), ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("w_dtype", C.c_int32),
        ("in_dims", C.c_int32 * 4), ("out_dims", C.c_int32 * 4),
        ("in_ld", C.c_int32), ("in1_ld", C.c_int32), ("out_ld", C.c_int32), ("out1_ld", C.c_int32),
        ("k", C.c_int32 * 3), ("s", C.c_int32 * 3), ("pad", C.c_int32 * 3),
        ("relu0", C.c_int32), ("relu1", C.c_int32), ("pad_is_zero", C.c_int32), ("ext_input", C.c_int32),
        ("crop", C.c_int32 * 3), ("src_dims", C.c_int32 * 4),
        ("kc", C.c_int32), ("bn", C.c_int32), ("brick", C.c_int32 * 4),
        ("pre_mean", C.c_float * 4), ("pre_scale", C.c_float * 4),
        ("in_wpitch", C.c_int32), ("out_wpitch", C.c_int32), ("out_wpad", C.c_int32), ("pre_unroll_w", C.c_int32), ("pool_k", C.c_int32 * 3), ("pool_dims", C.c_int32 * 3), ("pool_zero", C.c_int32),
        ("tc_halo", C.c_int32), ("tc_pair_pool", C.c_int32), ("out_split", C.c_int32), ("out_split2", C.c_int32), ("out2_ld", C.c_int32), ("pre_s2d", C.c_int32),
        ("in0_off", C.c_int64), ("in1_off", C.c_int64), ("out0_off", C.c_int64), ("out1_off", C.c_int64), ("out2_off", C.c_int64),
        ("w_off", C.c_int64), ("scale0_off", C.c_int64), ("shift0_off", C.c_int64),
        ("scale1_off", C.c_int64), ("shift1_off", C.c_int64),
        ("part_off", C.c_int64), ("part_bytes", C.c_int64), ("ksplit", C.c_int32), ("reserved0", C.c_int32),
    ]


class CseError(RuntimeError):
    pass


_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Loads the C-ABI library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise CseError("libcse_b200.so is not built (%s). Run `python -m cse_b200.build` "
                       "(or __graft_entry__.build()). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.cse_abi_version.restype = C.c_int
    lib.cse_last_error.restype = C.c_char_p
    lib.cse_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.cse_tune.argtypes = [C.c_char_p, i32]
    lib.cse_malloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.cse_free.argtypes = [vp]
    lib.cse_memcpy_h2d.argtypes = [vp, vp, C.c_size_t, vp]
    lib.cse_memcpy_d2h.argtypes = [vp, vp, C.c_size_t, vp]
    lib.cse_stream_synchronize.argtypes = [vp]
    lib.cse_plan_create.argtypes = [C.POINTER(vp), i32, i32]
    lib.cse_plan_add_op.argtypes = [vp, C.POINTER(CseOp)]
    lib.cse_plan_finalize.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, i64, i64]
    lib.cse_plan_run.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.cse_plan_run_range.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.cse_plan_run_from.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
    lib.cse_plan_num_input_ops.argtypes = [vp]
    lib.cse_plan_num_ops.argtypes = [vp]
    lib.cse_plan_last_launches.argtypes = [vp]
    lib.cse_plan_destroy.argtypes = [vp]
    lib.cse_plan_destroy.restype = None
    lib.cse_preprocess.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), vp, i32, i32, vp]
    lib.cse_vote.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.cse_vote_search.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
    lib.cse_assemble_clip.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.cse_resize_u8.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, C.c_double, C.c_double, vp]
    lib.cse_bgr2gray.argtypes = [vp, vp, C.c_longlong, vp]
    lib.cse_farneback_workspace_bytes.argtypes = [i32, i32, i32]
    lib.cse_farneback_workspace_bytes.restype = C.c_size_t
    lib.cse_farneback.argtypes = [vp, i32, i32, i32, C.c_double, i32, i32, i32, i32, C.c_double, vp, vp, C.c_size_t, vp]
    lib.cse_resize_linear_f32.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, vp]
    lib.cse_model_create.argtypes = [C.POINTER(vp), C.c_char_p, i32, i32, i32, i32, i32, i32]
    lib.cse_model_set_option.argtypes = [vp, C.c_char_p, i32]
    lib.cse_model_num_layers.argtypes = [vp]
    lib.cse_model_layer_info.argtypes = [vp, i32, C.c_char_p, i32, C.POINTER(C.c_int)]
    lib.cse_model_tensor_info.argtypes = [vp, i32, i32, C.POINTER(i64), C.POINTER(C.c_int), C.c_char_p, i32]
    lib.cse_model_set_weight.argtypes = [vp, i32, i32, vp, C.POINTER(i64), i32]
    lib.cse_model_lower.argtypes = [vp]
    lib.cse_model_pair_stems.argtypes = [vp, vp]
    lib.cse_model_num_ops.argtypes = [vp]
    lib.cse_model_get_op.argtypes = [vp, i32, C.POINTER(CseOp)]
    lib.cse_model_workspace_bytes.argtypes = [vp]
    lib.cse_model_workspace_bytes.restype = C.c_size_t
    lib.cse_model_weight_bytes.argtypes = [vp]
    lib.cse_model_weight_bytes.restype = C.c_size_t
    lib.cse_model_copy_weight_arena.argtypes = [vp, vp, C.c_size_t]
    lib.cse_model_logits_offset.argtypes = [vp]
    lib.cse_model_logits_offset.restype = i64
    lib.cse_model_probs_offset.argtypes = [vp]
    lib.cse_model_probs_offset.restype = i64
    lib.cse_model_finalize.argtypes = [vp, vp, C.c_size_t]
    lib.cse_model_forward.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.cse_model_forward_shared_input.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.cse_model_destroy.argtypes = [vp]
    lib.cse_model_destroy.restype = None
    for name in EXPORTS:
        getattr(lib, name)
    if lib.cse_abi_version() != ABI_VERSION:
        raise CseError("ABI version mismatch: library %d, binding %d" % (lib.cse_abi_version(), ABI_VERSION))
    _lib = lib
    # experiments: CSE_TUNE="pair_min_tiles=0,twin_min_tiles=2" is applied ONCE here (never read on the launch path)
    for item in filter(None, os.environ.get("CSE_TUNE", "").split(",")):
        key, _, val = item.partition("=")
        check(lib.cse_tune(key.strip().encode(), int(val)))
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().cse_last_error()
        raise CseError("libcse_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CseError("no CUDA device: the cse_b200 hot path has no CPU fallback")
    return torch


def tune(key: str, value: int) -> None:
    """cse_tune: override a launch heuristic (tests force the twin / shared-B conv modes on small shapes)."""
    check(load_library().cse_tune(key.encode(), int(value)))


def device_info():
    lib = load_library()
    sm, mj, mn = C.c_int(), C.c_int(), C.c_int()
    check(lib.cse_device_info(C.byref(sm), C.byref(mj), C.byref(mn)))
    return sm.value, mj.value, mn.value


def current_stream_ptr() -> int:
    torch = require_cuda()
    return torch.cuda.current_stream().cuda_stream


# --------------------------------------------------------------------------- #
# stand-alone kernels
# --------------------------------------------------------------------------- #
def preprocess(clips_u8, out_dtype="bf16", out_channels=None, crop=None, mean=None, scale=None):
    """uint8 NDHWC device tensor -> float NDHWC (channels zero-padded to out_channels).
    Reference behaviour (train.py:466-478): no crop, mean 0, scale 1."""
    torch = require_cuda()
    lib = load_library()
    assert clips_u8.is_cuda and clips_u8.dtype == torch.uint8 and clips_u8.is_contiguous()
    n, t, h, w, c = clips_u8.shape
    t0, h0, w0, to, ho, wo = crop if crop is not None else (0, 0, 0, t, h, w)
    ld = out_channels or c
    tdt = torch.bfloat16 if out_dtype == "bf16" else torch.float32
    out = torch.empty((n, to, ho, wo, ld), dtype=tdt, device=clips_u8.device)
    fm = (C.c_float * c)(*mean) if mean is not None else None
    fs = (C.c_float * c)(*scale) if scale is not None else None
    check(lib.cse_preprocess(clips_u8.data_ptr(), n, t, h, w, c, t0, h0, w0, to, ho, wo, fm, fs,
                             out.data_ptr(), BF16 if out_dtype == "bf16" else F32, ld, current_stream_ptr()))
    return out


def vote(probs, weights=None, mode="SUM", return_summed=False):
    """probs: device tensor [M,N,C] float32 or float64; weights: None (ones) or float64 [M] device
    tensor; mode 'SUM'/'WEIGHTED' (weighted sum) or 'MAXIMUM'.  -> int32 [N] (and fp64 [N,C])."""
    torch = require_cuda()
    lib = load_library()
    assert probs.is_cuda and probs.is_contiguous() and probs.dim() == 3
    assert probs.dtype in (torch.float32, torch.float64)
    m, n, c = probs.shape
    pred = torch.empty((n,), dtype=torch.int32, device=probs.device)
    summed = torch.empty((n, c), dtype=torch.float64, device=probs.device) if return_summed else None
    wptr = None
    if weights is not None:
        weights = weights.to(device=probs.device, dtype=torch.float64).contiguous()
        assert weights.numel() == m
        wptr = weights.data_ptr()
    check(lib.cse_vote(probs.data_ptr(), 1 if probs.dtype == torch.float64 else 0, wptr,
                       1 if mode == "MAXIMUM" else 0, m, n, c, pred.data_ptr(),
                       summed.data_ptr() if summed is not None else None, current_stream_ptr()))
    return (pred, summed) if return_summed else pred


def vote_search(probs_f64, weight_matrix, labels):
    """Batched weighted vote: probs [M,N,C] f64, weight_matrix [W,M] f64, labels [N] int32
    -> int32 [W] number of correctly voted clips per weight vector."""
    torch = require_cuda()
    lib = load_library()
    m, n, c = probs_f64.shape
    w = weight_matrix.shape[0]
    assert probs_f64.dtype == torch.float64 and weight_matrix.dtype == torch.float64
    assert labels.dtype == torch.int32 and weight_matrix.shape[1] == m
    correct = torch.empty((w,), dtype=torch.int32, device=probs_f64.device)
    check(lib.cse_vote_search(probs_f64.contiguous().data_ptr(), weight_matrix.contiguous().data_ptr(),
                              labels.contiguous().data_ptr(), w, m, n, c, correct.data_ptr(),
                              current_stream_ptr()))
    return correct


def assemble_clip(frames_u8, t: int, h: int, w: int, out=None):
    """Decoded frames of one video, uint8 device tensor [n_frames, Hs, Ws(, C)] -> uint8 [T, H, W(, C)]:
    select_frames + cv2.resize (train.py:132-145, 286), bit-exact with OpenCV's 8-bit INTER_LINEAR."""
    torch = require_cuda()
    lib = load_library()
    assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous()
    assert frames_u8.dim() in (3, 4)
    n, hs, ws = frames_u8.shape[:3]
    c = frames_u8.shape[3] if frames_u8.dim() == 4 else 1
    shape = (t, h, w) + ((c,) if frames_u8.dim() == 4 else ())
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=frames_u8.device)
    assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == shape
    check(lib.cse_assemble_clip(frames_u8.data_ptr(), n, hs, ws, c, out.data_ptr(), t, h, w, current_stream_ptr()))
    return out


def resize_u8(images_u8, h: int = None, w: int = None, fx: float = None, fy: float = None):
    """uint8 device tensor [n, Hs, Ws(, C)] -> [n, h, w(, C)]: cv2.resize(img, (w, h)), or with fx / fy
    cv2.resize(img, None, fx=fx, fy=fy) (size = cvRound(size * factor)), bit-exact with OpenCV's 8-bit INTER_LINEAR."""
    torch = require_cuda()
    lib = load_library()
    assert images_u8.is_cuda and images_u8.dtype == torch.uint8 and images_u8.is_contiguous() and images_u8.dim() in (3, 4)
    n, hs, ws = images_u8.shape[:3]
    c = images_u8.shape[3] if images_u8.dim() == 4 else 1
    if fx is not None:
        h, w = round(hs * fy), round(ws * fx)          # cvRound: half to even, like Python's round
    out = torch.empty((n, h, w) + ((c,) if images_u8.dim() == 4 else ()), dtype=torch.uint8, device=images_u8.device)
    check(lib.cse_resize_u8(images_u8.data_ptr(), n, hs, ws, c, out.data_ptr(), h, w, float(fx or 0.0), float(fy or 0.0),
                            current_stream_ptr()))
    return out


def bgr2gray(images_u8):
    """uint8 device tensor [..., 3] (BGR) -> [...]: cv2.cvtColor(COLOR_BGR2GRAY), bit-exact."""
    torch = require_cuda()
    lib = load_library()
    assert images_u8.is_cuda and images_u8.dtype == torch.uint8 and images_u8.is_contiguous() and images_u8.shape[-1] == 3
    out = torch.empty(images_u8.shape[:-1], dtype=torch.uint8, device=images_u8.device)
    check(lib.cse_bgr2gray(images_u8.data_ptr(), out.data_ptr(), out.numel(), current_stream_ptr()))
    return out


def farneback(gray_u8, pyr_scale=0.5, levels=5, winsize=11, iterations=5, poly_n=5, poly_sigma=1.1):
    """uint8 device tensor [F, H, W] of gray frames -> float32 [F - 1, H, W, 2]: cv2.calcOpticalFlowFarneback between
    consecutive frames (defaults = the reference's call, train.py:320-322)."""
    torch = require_cuda()
    lib = load_library()
    assert gray_u8.is_cuda and gray_u8.dtype == torch.uint8 and gray_u8.is_contiguous() and gray_u8.dim() == 3
    f, h, w = gray_u8.shape
    flow = torch.empty((f - 1, h, w, 2), dtype=torch.float32, device=gray_u8.device)
    nbytes = lib.cse_farneback_workspace_bytes(f, h, w)
    work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=gray_u8.device)
    check(lib.cse_farneback(gray_u8.data_ptr(), f, h, w, float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n),
                            float(poly_sigma), flow.data_ptr(), work.data_ptr(), work.numel() * 8, current_stream_ptr()))
    return flow


def resize_linear_f32(images_f32, h: int, w: int):
    """float32 device tensor [n, Hs, Ws(, C)] -> [n, h, w(, C)]: cv2.resize(img, (w, h)) on CV_32F, bit-exact."""
    torch = require_cuda()
    lib = load_library()
    assert images_f32.is_cuda and images_f32.dtype == torch.float32 and images_f32.is_contiguous() and images_f32.dim() in (3, 4)
    n, hs, ws = images_f32.shape[:3]
    c = images_f32.shape[3] if images_f32.dim() == 4 else 1
    out = torch.empty((n, h, w) + ((c,) if images_f32.dim() == 4 else ()), dtype=torch.float32, device=images_f32.device)
    check(lib.cse_resize_linear_f32(images_f32.data_ptr(), n, hs, ws, c, out.data_ptr(), h, w, current_stream_ptr()))
    return out


# --------------------------------------------------------------------------- #
# native member model (cse_model_*): graph construction + lowering inside the library
# --------------------------------------------------------------------------- #
class NativeModel:
    """Thin handle on a ``cse_model``: what a non-Python binder does - name the architecture, hand over the Keras weight
    tensors in ``model.layers`` order, finalize, forward.  (The Python product path lowers in ``lowering.py``;
    ``tests/test_model_api.py`` holds both lowerings to byte-identical plans.)"""

    def __init__(self, model_type: str, shape, nb_classes: int = 11, precision: str = "bf16", max_batch: int = 8,
                 persist_input: bool = False, flow_input_f32: bool = False):
        self.lib = load_library()
        self.handle = C.c_void_p()
        t, h, w = (int(v) for v in shape[:3])
        check(self.lib.cse_model_create(C.byref(self.handle), model_type.encode(), t, h, w, int(nb_classes),
                                        BF16 if precision == "bf16" else F32, int(max_batch)))
        self.nb_classes, self.max_batch = int(nb_classes), int(max_batch)
        if persist_input:
            check(self.lib.cse_model_set_option(self.handle, b"persist_input", 1))
        if flow_input_f32:
            check(self.lib.cse_model_set_option(self.handle, b"flow_input_f32", 1))

    def layers(self):
        """[(layer name, [(tensor name, shape), ...])] of the weighted layers in Keras model.layers order."""
        out = []
        for i in range(self.lib.cse_model_num_layers(self.handle)):
            name = C.create_string_buffer(256)
            nt = C.c_int()
            check(self.lib.cse_model_layer_info(self.handle, i, name, 256, C.byref(nt)))
            tensors = []
            for j in range(nt.value):
                dims = (C.c_int64 * 5)()
                nd = C.c_int()
                tn = C.create_string_buffer(256)
                check(self.lib.cse_model_tensor_info(self.handle, i, j, dims, C.byref(nd), tn, 256))
                tensors.append((tn.value.decode(), tuple(dims[k] for k in range(nd.value))))
            out.append((name.value.decode(), tensors))
        return out

    def set_weights(self, weights):
        """weights: {keras layer name: [arrays in Keras order]} (what hdf5.read_keras_weights + assign_positional give)."""
        import numpy as np
        for i, (lname, tensors) in enumerate(self.layers()):
            arrs = weights[lname]
            if len(arrs) != len(tensors):
                raise CseError("layer %s: %d tensors expected, %d given" % (lname, len(tensors), len(arrs)))
            for j, a in enumerate(arrs):
                a = np.ascontiguousarray(a, dtype=np.float32)
                dims = (C.c_int64 * a.ndim)(*a.shape)
                check(self.lib.cse_model_set_weight(self.handle, i, j, a.ctypes.data, dims, a.ndim))

    def pair_stems(self, follow: "NativeModel"):
        """This member leads, `follow` follows: one N = 128 stem GEMM for both (cse_model_pair_stems).  Weights of both are
        set, neither is lowered; finalize both on one shared workspace, forward the leader first, the follower with
        shared_input=True."""
        check(self.lib.cse_model_pair_stems(self.handle, follow.handle))

    def lower(self):
        check(self.lib.cse_model_lower(self.handle))

    def ops(self):
        out = []
        for i in range(self.lib.cse_model_num_ops(self.handle)):
            s = CseOp()
            check(self.lib.cse_model_get_op(self.handle, i, C.byref(s)))
            out.append(s)
        return out

    def weight_arena(self):
        import numpy as np
        n = self.lib.cse_model_weight_bytes(self.handle)
        buf = np.empty(n, np.uint8)
        check(self.lib.cse_model_copy_weight_arena(self.handle, buf.ctypes.data, n))
        return buf

    def workspace_bytes(self) -> int:
        return int(self.lib.cse_model_workspace_bytes(self.handle))

    def finalize(self, shared_workspace=None):
        require_cuda()
        if shared_workspace is None:
            check(self.lib.cse_model_finalize(self.handle, None, 0))
        else:
            ptr = (shared_workspace.data_ptr() + 1023) // 1024 * 1024
            check(self.lib.cse_model_finalize(self.handle, ptr, shared_workspace.numel() - (ptr - shared_workspace.data_ptr())))
            self._keep = shared_workspace

    def forward(self, rgb, flow=None, shared_input: bool = False):
        """rgb / flow: contiguous CUDA tensors [n,T,H,W,C] -> (logits, probs) fp32 CUDA tensors [n, nb_classes]."""
        torch = require_cuda()
        n = rgb.shape[0]
        logits = torch.empty((n, self.nb_classes), dtype=torch.float32, device=rgb.device)
        probs = torch.empty_like(logits)
        fn = self.lib.cse_model_forward_shared_input if shared_input else self.lib.cse_model_forward
        check(fn(self.handle, rgb.data_ptr(), flow.data_ptr() if flow is not None else None, n, logits.data_ptr(),
                 probs.data_ptr(), current_stream_ptr()))
        return logits, probs

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.cse_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
