"""Clip input contract of the reference's DataGenerator (train.py:361-488), for evaluation.

The reference decodes every clip with OpenCV, keeps ``frames[::len//T][:T]`` (select_frames,
train.py:132-145), ``cv2.resize``s each frame to (W, H) (bilinear) and stores the BGR uint8 values
into a float32 batch - no crop, no mean/std (train.py:245-291, 466-478).  TwoStream clips add a
2-channel flow volume built from two gray videos (TV-L1, train.py:196-221).

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


def load_rgb_clip(path: str, t: int, h: int, w: int) -> np.ndarray:
    """-> uint8 [T,H,W,3] BGR (get_onestream_videoclip, train.py:245-291)."""
    path = path.strip()
    if path.endswith(".npy"):
        clip = np.load(path)
        frames = select_frames(list(clip), t)
    else:
        frames = select_frames(_read_video(path), t)
    out = np.asarray([_resize(f, w, h) for f in frames], dtype=np.uint8)
    if out.shape != (t, h, w, 3):
        raise ValueError("clip %s decodes to %r, expected %r" % (path, out.shape, (t, h, w, 3)))
    return out


def load_flow_clip(xpath: str, ypath: str, t: int, h: int, w: int) -> np.ndarray:
    """-> uint8 [T,H,W,2] from two gray flow videos (TV-L1, train.py:196-221)."""
    chans = []
    for p in (xpath.strip(), ypath.strip()):
        if p.endswith(".npy"):
            frames = list(np.load(p))
        else:
            frames = [f[..., 0] if f.ndim == 3 else f for f in _read_video(p)]
        frames = select_frames(frames, t)
        chans.append(np.asarray([_resize(f, w, h) for f in frames], dtype=np.uint8))
    out = np.stack(chans, axis=-1)
    if out.shape != (t, h, w, 2):
        raise ValueError("flow clip decodes to %r, expected %r" % (out.shape, (t, h, w, 2)))
    return out


class ClipSequence:
    """DataGenerator for evaluation: ordered, not shuffled, non-augmented (evaluate_ensemble.py:1032-1040)."""

    def __init__(self, video_data, model_type, input_shape, num_classes, batch_size=1,
                 optical_flow_status="TVL1_precomputed", augmentation_status="non_augmented",
                 augmentation_frequency=0, shuffle=False):
        if shuffle or augmentation_status != "non_augmented":
            raise ValueError("evaluation clips are ordered and non-augmented")
        if model_type == "TWOSTREAM_I3D" and optical_flow_status != "TVL1_precomputed":
            raise NotImplementedError("on-the-fly Farneback flow is outside the accelerated path (SURVEY §8f)")
        self.video_data = video_data
        self.model_type = model_type
        self.input_shape = tuple(input_shape)
        self.num_classes = num_classes
        self.batch_size = int(batch_size)
        self.n = int(video_data.count().iloc[0]) if hasattr(video_data.count(), "iloc") else int(video_data.count()[0])

    def __len__(self):
        return int(np.ceil(self.n / self.batch_size))

    def __getitem__(self, index):
        idx = range(index * self.batch_size, min((index + 1) * self.batch_size, self.n))
        t, h, w = self.input_shape[:3]
        vd = self.video_data
        labels = np.asarray([vd["class"].values[i] for i in idx], dtype=int)
        onehot = np.zeros((len(labels), self.num_classes), np.float32)
        onehot[np.arange(len(labels)), labels % self.num_classes] = 1.0
        rgb = np.stack([load_rgb_clip(vd["rgbclips_path"].values[i], t, h, w) for i in idx])
        if self.model_type == "TWOSTREAM_I3D":
            flow = np.stack([load_flow_clip(vd["x_axis_flowclips_path"].values[i],
                                            vd["y_axis_flowclips_path"].values[i], t, h, w) for i in idx])
            return [rgb, flow], onehot
        return rgb, onehot
