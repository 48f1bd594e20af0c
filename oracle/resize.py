"""ORACLE (test infrastructure, not product code) - numpy restatement of the clip assembly that feeds
the hot path: frame selection and ``cv2.resize`` (bilinear, uint8).

PINNED: (1) tests/test_oracle_resize.py compares ``resize_linear_u8`` with the real ``cv2.resize`` of
the container's OpenCV on random shapes (bit-exact); (2) tests/golden/clips_golden.npz holds outputs of
the reference's OWN ``get_onestream_videoclip`` / ``get_twostream_videoclip`` / ``select_frames``
(extracted with ``ast`` from /root/reference/train.py and executed by tools/make_golden_clips.py on the
small videos committed next to it).

Restated:
  select_frames               train.py:132-145
  cv2.resize(frame, (W, H))   train.py:286, 209-214  - OpenCV (requirements.txt:5 pins 4.0.1; the
      container has 4.13) imgproc/resize.cpp, INTER_LINEAR on CV_8U: fixed point with
      INTER_RESIZE_COEF_BITS = 11.  Horizontal taps: fx = float((dx + 0.5) * scale - 0.5), sx =
      floor(fx), fx -= sx; sx < 0 -> (0, fx = 0); sx >= W-1 -> (W-1, fx = 0); weights
      cvRound(f * 2048) as int16.  Vertical taps: same fy WITHOUT the reset, the two row indices are
      clipped to [0, H-1] instead.  Pixel: ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
      with r = a0 * S[sx] + a1 * S[sx+1] (VResizeLinear<uchar,int,short,...>).
  TV-L1 flow frames           train.py:334-357 (BGR2GRAY of each decoded frame), :196-221
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def select_frame_indices(n_frames: int, frames_per_video: int):
    """Indices kept by select_frames (train.py:132-145): frames[::step][:T], step = max(1, n // T)."""
    step = n_frames // frames_per_video
    if step == 0:
        step = 1
    return list(range(0, n_frames, step))[:frames_per_video]


def linear_taps(src: int, dst: int, horizontal: bool, inv_scale: float = None):
    """-> (i0, i1, w0, w1): source indices and int16 fixed-point weights of every destination index.  `inv_scale` is
    the fx / fy of cv2.resize(img, None, fx=, fy=), which OpenCV uses as given instead of dst / src."""
    scale = 1.0 / (float(dst) / float(src) if inv_scale is None else float(inv_scale))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0, src - 1
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int64)                      # cvRound: half to even
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int64)
    i1 = np.clip(s + 1, 0, src - 1)
    i0 = np.clip(s, 0, src - 1)
    return i0, i1, w0, w1


def resize_linear_u8(img: np.ndarray, width: int = None, height: int = None, fx: float = None, fy: float = None) -> np.ndarray:
    """cv2.resize(img, (width, height)), or cv2.resize(img, None, fx=fx, fy=fy) (output size = cvRound(size * factor),
    taps from 1 / factor), for uint8 [H,W] or [H,W,C] images, bit for bit."""
    img = np.asarray(img)
    if img.dtype != np.uint8:
        raise TypeError("uint8 frames only")
    hs, ws = img.shape[:2]
    if fx is not None:
        width, height = int(np.rint(ws * fx)), int(np.rint(hs * fy))
    x0, x1, ax0, ax1 = linear_taps(ws, width, True, fx)
    y0, y1, by0, by1 = linear_taps(hs, height, False, fy)
    im = img.astype(np.int64).reshape(hs, ws, -1)
    rows = im[:, x0] * ax0[None, :, None] + im[:, x1] * ax1[None, :, None]
    r0, r1 = rows[y0], rows[y1]
    out = (((by0[:, None, None] * (r0 >> 4)) >> 16) + ((by1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8).reshape((height, width) + img.shape[2:])


def assemble_clip(frames, t: int, h: int, w: int) -> np.ndarray:
    """get_onestream_videoclip after decoding (train.py:279-291): select, resize, stack -> uint8 [T,H,W(,C)]."""
    idx = select_frame_indices(len(frames), t)
    return np.asarray([resize_linear_u8(frames[i], w, h) for i in idx], dtype=np.uint8)
