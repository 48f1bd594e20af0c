"""Clip input contract of the reference's DataGenerator (train.py:361-488), for evaluation.

The reference decodes every clip with OpenCV, keeps ``frames[::len//T][:T]`` (select_frames,
train.py:132-145), ``cv2.resize``s each frame to (W, H) (bilinear) and stores the BGR uint8 values
into a float32 batch - no crop, no mean/std (train.py:245-291, 466-478).  TwoStream clips add a
2-channel flow volume built from two gray videos (TV-L1, train.py:196-221).

The FarneBack_onTheFly TwoStream variant (train.py:294-332, the SPECIALCASE token of the default global list,
evaluate_ensemble.py:1365-1386) computes a dense float32 flow between consecutive frames in the loader with OpenCV;
``farneback_flow`` makes the same OpenCV calls in the same order (loader-side, like the reference: decode and flow
extraction are CPU I/O in front of the accelerated path), and the float32 flow volume goes to the GPU as is.

``ClipSequence`` keeps the keras.utils.Sequence protocol (``__len__``, ``__getitem__(i) ->
(x | [x_rgb, x_flow], y_onehot)``) but yields uint8 arrays (what the frames are before the float32
store) so a clip costs 1 byte per value on the way to the GPU.  Video decoding is CPU I/O and
outside the accelerated path: clips can equally be pre-decoded ``.npy`` files of shape
[T,H,W,C] uint8, which is what the tests use.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np


def select_frames(frames: Sequence, frames_per_video: int):
    step = len(frames) // frames_per_video
    if step == 0:
        step = 1
    return frames[::step][:frames_per_video]


def _read_video(path: str) -> List[np.ndarray]:
    try:
        import cv2
    except ImportError as e:            # pragma: no cover
        raise RuntimeError("OpenCV is needed to decode %s (or provide pre-decoded .npy clips)" % path) from e
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        cap.open(path)
    frames = []
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        frames.append(frame)
    cap.release()
    if not frames:
        raise IOError("no frames decoded from %s" % path)
    return frames


def _resize(frame: np.ndarray, width: int, height: int) -> np.ndarray:
    if frame.shape[0] == height and frame.shape[1] == width:
        return frame
    import cv2
    return cv2.resize(frame, (width, height))


def _gray(frame: np.ndarray) -> np.ndarray:
    """opticalflow_TVL1_retriever (train.py:334-357): every decoded flow frame goes through BGR2GRAY."""
    if frame.ndim == 2:
        return frame
    import cv2
    return cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)


def decode_frames(path: str, gray: bool = False) -> List[np.ndarray]:
    """All frames of a video (OpenCV) or of a pre-decoded ``.npy`` [F,H,W(,C)] uint8 array."""
    path = path.strip()
    frames = list(np.load(path)) if path.endswith(".npy") else _read_video(path)
    return [_gray(f) for f in frames] if gray else frames


def _select(frames: List[np.ndarray], t: int) -> np.ndarray:
    """select_frames -> one contiguous uint8 array [T', Hs, Ws(, C)] (host only: safe in worker threads)."""
    return np.ascontiguousarray(np.asarray(select_frames(frames, t), dtype=np.uint8))


def _finish(sel: np.ndarray, t: int, h: int, w: int, device=None):
    """Resize the selected frames -> uint8 [T,H,W(,C)].  device=None: on the CPU with cv2 (the reference's
    own path); otherwise the frames are uploaded once and resized by cse_assemble_clip (bit-identical
    output, csrc/ingest.cu) and the clip stays on the GPU."""
    if device is None:
        return np.asarray([_resize(f, w, h) for f in sel], dtype=np.uint8)
    from . import runtime as rt
    torch = rt.require_cuda()
    if len(sel) < t:
        raise ValueError("select_frames keeps %d < %d frames" % (len(sel), t))
    src = torch.from_numpy(sel).to(device)
    with torch.cuda.device(src.device):
        return rt.assemble_clip(src, t, h, w)


def _assemble(frames: List[np.ndarray], t: int, h: int, w: int, device=None):
    return _finish(_select(frames, t), t, h, w, device)


def load_rgb_clip(path: str, t: int, h: int, w: int, device=None):
    """-> uint8 [T,H,W,3] BGR (get_onestream_videoclip, train.py:245-291); numpy, or a CUDA tensor
    when `device` is given."""
    out = _assemble(decode_frames(path), t, h, w, device)
    if tuple(out.shape) != (t, h, w, 3):
        raise ValueError("clip %s decodes to %r, expected %r" % (path, tuple(out.shape), (t, h, w, 3)))
    return out


def load_flow_clip(xpath: str, ypath: str, t: int, h: int, w: int, device=None):
    """-> uint8 [T,H,W,2] from two gray flow videos (TV-L1, train.py:196-221, 334-357)."""
    chans = [_assemble(decode_frames(p, gray=True), t, h, w, device) for p in (xpath, ypath)]
    if device is None:
        out = np.stack(chans, axis=-1)
    else:
        import torch
        out = torch.stack(chans, dim=-1).contiguous()
    if tuple(out.shape) != (t, h, w, 2):
        raise ValueError("flow clip decodes to %r, expected %r" % (tuple(out.shape), (t, h, w, 2)))
    return out


def farneback_flow(frames: Sequence[np.ndarray]) -> np.ndarray:
    """Dense Farneback flow between consecutive frames -> float32 [len(frames) - 1, h', w', 2], as
    opticalflow_FarneBack_extractor does it (train.py:294-332): the first frame is resized so that its longest side
    (of H, W, 3) becomes 224 and THEN converted to gray, every later frame is converted to gray first and then
    resized by the same factor (the order matters for the bytes); flow parameters pyr_scale 0.5, 5 levels, window 11,
    5 iterations, poly_n 5, poly_sigma 1.1, no flags."""
    import cv2
    factor = 224 / max(frames[0].shape)
    previous = cv2.cvtColor(cv2.resize(frames[0], None, fx=factor, fy=factor), cv2.COLOR_BGR2GRAY)
    flows = []
    for frame in frames[1:]:
        if frame is None:
            continue
        current = cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), None, fx=factor, fy=factor)
        flows.append(cv2.calcOpticalFlowFarneback(previous, current, None, pyr_scale=0.5, levels=5, winsize=11,
                                                  iterations=5, poly_n=5, poly_sigma=1.1, flags=0))
        previous = current
    return np.asarray(flows)


def farneback_flow_clip(frames: Sequence[np.ndarray], t: int, h: int, w: int) -> np.ndarray:
    """-> float32 [T,H,W,2]: select_frames over the len(frames) - 1 flow fields, each resized to (W, H) (bilinear,
    float) - the FarneBack branch of get_twostream_videoclip (train.py:223-239)."""
    import cv2
    sel = select_frames(list(farneback_flow(frames)), t)
    out = np.asarray([cv2.resize(f, (w, h)) for f in sel], dtype=np.float32)
    if tuple(out.shape) != (t, h, w, 2):
        raise ValueError("flow clip has shape %r, expected %r" % (tuple(out.shape), (t, h, w, 2)))
    return out


def farneback_flow_clip_device(frames: Sequence[np.ndarray], t: int, h: int, w: int, device, frames_dev=None):
    """farneback_flow_clip on the GPU -> float32 CUDA tensor [T,H,W,2].  The decoded frames are uploaded once (or taken
    from `frames_dev`, uint8 [F,Hs,Ws,3]); frame scaling and gray conversion are bit-exact with cv2 (cse_resize_u8,
    cse_bgr2gray, in the extractor's order: first frame resize -> gray, later frames gray -> resize), the flow of all
    consecutive pairs is one cse_farneback call (bit-identical to oracle/farneback.py, within 1e-4 pixel of cv2 - its
    SIMD summation order is not reproducible), select_frames is a strided slice, the final resize is cse_resize_linear_f32
    (bit-exact with cv2.resize on floats)."""
    import torch
    from . import runtime as rt
    if frames_dev is None:
        frames_dev = torch.from_numpy(np.ascontiguousarray(np.stack(frames))).to(device)
    if frames_dev.dim() != 4 or frames_dev.shape[-1] != 3 or frames_dev.shape[0] < 2:
        raise ValueError("the Farneback extractor needs >= 2 BGR frames, got %r" % (tuple(frames_dev.shape),))
    factor = 224 / max(frames_dev.shape[1:])
    with torch.cuda.device(frames_dev.device):
        first = rt.bgr2gray(rt.resize_u8(frames_dev[:1].contiguous(), fx=factor, fy=factor))
        rest = rt.resize_u8(rt.bgr2gray(frames_dev[1:].contiguous()), fx=factor, fy=factor)
        flows = rt.farneback(torch.cat([first, rest]))                      # [F-1,h',w',2]
        step = max(1, flows.shape[0] // t)
        sel = flows[::step][:t].contiguous()
        if sel.shape[0] != t:
            raise ValueError("flow clip has %d fields, expected %d" % (sel.shape[0], t))
        return rt.resize_linear_f32(sel, h, w)


def load_farneback_twostream_clip(path: str, t: int, h: int, w: int, device=None):
    """-> (uint8 [T,H,W,3] BGR, float32 [T,H,W,2] flow) of one video, decoded once; with `device` both are CUDA tensors
    and the flow is computed on that GPU."""
    frames = decode_frames(path)
    if device is None:
        return _assemble(frames, t, h, w), farneback_flow_clip(frames, t, h, w)
    return _assemble(frames, t, h, w, device), farneback_flow_clip_device(frames, t, h, w, device)


class ClipSequence:
    """DataGenerator for evaluation: ordered, not shuffled, non-augmented (evaluate_ensemble.py:1032-1040)."""

    def __init__(self, video_data, model_type, input_shape, num_classes, batch_size=1,
                 optical_flow_status="TVL1_precomputed", augmentation_status="non_augmented",
                 augmentation_frequency=0, shuffle=False, device=None):
        """device: None = clips are assembled on the CPU and returned as numpy arrays (the reference's
        behaviour); a CUDA device = frames are resized on that GPU and x is returned as uint8 CUDA
        tensors, which Member.predict / predict_generator take directly (FarneBack_onTheFly: the flow is computed on
        that GPU too, cse_farneback, and returned as a float32 CUDA tensor)."""
        if shuffle or augmentation_status != "non_augmented":
            raise ValueError("evaluation clips are ordered and non-augmented")
        if optical_flow_status not in ("TVL1_precomputed", "FarneBack_onTheFly"):
            raise ValueError("optical_flow_status must be TVL1_precomputed or FarneBack_onTheFly")
        self.farneback = model_type == "TWOSTREAM_I3D" and optical_flow_status == "FarneBack_onTheFly"
        self.video_data = video_data
        self.model_type = model_type
        self.input_shape = tuple(input_shape)
        self.num_classes = num_classes
        self.batch_size = int(batch_size)
        self.device = device
        # CSE_CPU_FLOW=1 keeps the reference's own OpenCV flow (bit-identical to its extractor) with GPU clip assembly
        self.flow_on_device = device is not None and os.environ.get("CSE_CPU_FLOW", "0") != "1"
        self.n = int(video_data.count().iloc[0]) if hasattr(video_data.count(), "iloc") else int(video_data.count()[0])

    def __len__(self):
        return int(np.ceil(self.n / self.batch_size))

    # The work of one batch is split in two so that a pool of worker threads can decode ahead
    # (cv2 releases the GIL) while the main thread keeps every CUDA call: load_raw = decode + select_frames
    # (host only), assemble = resize (+ upload) and stacking.  __getitem__ = assemble(load_raw(i)).
    def load_raw(self, index):
        idx = range(index * self.batch_size, min((index + 1) * self.batch_size, self.n))
        t = self.input_shape[0]
        vd = self.video_data
        labels = np.asarray([vd["class"].values[i] for i in idx], dtype=int)
        onehot = np.zeros((len(labels), self.num_classes), np.float32)
        onehot[np.arange(len(labels)), labels % self.num_classes] = 1.0
        raw = []
        for i in idx:
            frames = decode_frames(vd["rgbclips_path"].values[i])
            item = {"rgb": _select(frames, t)}
            if self.farneback and not self.flow_on_device:   # dense flow in the loader thread, like the reference's workers
                item["flow"] = farneback_flow_clip(frames, t, self.input_shape[1], self.input_shape[2])
            elif self.farneback:                          # computed on the GPU in assemble()
                item["frames"] = frames
            elif self.model_type == "TWOSTREAM_I3D":
                item["fx"] = _select(decode_frames(vd["x_axis_flowclips_path"].values[i], gray=True), t)
                item["fy"] = _select(decode_frames(vd["y_axis_flowclips_path"].values[i], gray=True), t)
            raw.append(item)
        return raw, onehot

    def assemble(self, raw):
        t, h, w = self.input_shape[:3]
        if self.device is None:
            stack, last = np.stack, (lambda a, b: np.stack([a, b], axis=-1))
        else:
            import torch
            stack, last = torch.stack, (lambda a, b: torch.stack([a, b], dim=-1).contiguous())
        rgb = stack([_finish(r["rgb"], t, h, w, self.device) for r in raw])
        if tuple(rgb.shape[1:]) != (t, h, w, 3):
            raise ValueError("clips decode to %r, expected %r" % (tuple(rgb.shape[1:]), (t, h, w, 3)))
        if self.farneback:
            if self.flow_on_device:
                return [rgb, stack([farneback_flow_clip_device(r["frames"], t, h, w, self.device) for r in raw])]
            flow = np.stack([r["flow"] for r in raw])
            if self.device is not None:
                import torch
                flow = torch.from_numpy(flow).to(self.device)
            return [rgb, flow]
        if self.model_type == "TWOSTREAM_I3D":
            flow = stack([last(_finish(r["fx"], t, h, w, self.device), _finish(r["fy"], t, h, w, self.device))
                          for r in raw])
            if tuple(flow.shape[1:]) != (t, h, w, 2):
                raise ValueError("flow clips decode to %r, expected %r" % (tuple(flow.shape[1:]), (t, h, w, 2)))
            return [rgb, flow]
        return rgb

    def __getitem__(self, index):
        raw, onehot = self.load_raw(index)
        return self.assemble(raw), onehot


def iterate_batches(generator, first: int, last: int, workers: int = 1, depth: int = 0):
    """Yields generator[first] .. generator[last] in order.  With workers > 1 and a generator that offers
    load_raw / assemble (ClipSequence), decoding runs `depth` batches ahead in a thread pool - the
    reference's `workers` (evaluate_ensemble.py:1053-1056 passes it to predict_generator) - while the
    results are consumed strictly in order on the calling thread."""
    n = last - first + 1
    if n <= 0:
        return
    if workers is None or workers <= 1 or not hasattr(generator, "load_raw"):
        for b in range(first, last + 1):
            yield generator[b]
        return
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    depth = depth or 2 * workers
    with ThreadPoolExecutor(max_workers=workers) as pool:
        pending = deque()
        nxt = first
        while nxt <= last and len(pending) < depth:
            pending.append(pool.submit(generator.load_raw, nxt))
            nxt += 1
        while pending:
            raw, onehot = pending.popleft().result()
            if nxt <= last:
                pending.append(pool.submit(generator.load_raw, nxt))
                nxt += 1
            yield generator.assemble(raw), onehot
