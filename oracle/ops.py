"""ORACLE (test infrastructure, not product code) - layer semantics of Keras 2.2.4 /
TensorFlow 1.15 (channels_last) restated with torch CPU ops.

PARITY UNPINNED for the neural-network part: the arithmetic of the reference lives
in un-vendored third-party packages (keras==2.2.4, tensorflow-gpu==1.15.0,
/root/reference/requirements.txt:3-4) that cannot be installed here, and the
reference ships no tests, golden vectors or weights.  Every rule below is therefore
pinned only by hand-computed known-answer tests (tests/test_oracle_semantics.py)
and by an independent naive numpy loop implementation (``naive_*`` below).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.

Call sites restated (all in /root/reference/train.py):
  Conv3D                 653-658, 1230-1258, 1294-1298, 1316-1320, 1338-1345, 1374-1379
  BatchNormalization     665 (scale=False), 1280 (default, eps=1e-3)
  MaxPooling3D           1029-1187, 1233-1261, 1487
  AveragePooling3D       1215-1217, 1504-1507
  ZeroPadding3D          1259
  Flatten/Dense/softmax  838-841, 1006-1007, 1262-1268, 1508-1515
All tensors are NDHWC ``[N,D,H,W,C]``; kernels ``[kd,kh,kw,Cin,Cout]``; Dense ``[in,out]``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


def _same(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def _pads(dhw, k, s, padding):
    if padding == "valid":
        return [(0, 0)] * 3
    assert padding == "same"
    return [_same(dhw[i], k[i], s[i])[1:] for i in range(3)]


def _to_ncdhw(x):
    return x.permute(0, 4, 1, 2, 3)


def _to_ndhwc(x):
    return x.permute(0, 2, 3, 4, 1).contiguous()


def conv3d(x, kernel, bias=None, strides=(1, 1, 1), padding="same"):
    """Cross-correlation, zero padding; TF SAME puts the extra pad at the end."""
    k = kernel.shape[:3]
    (d0, d1), (h0, h1), (w0, w1) = _pads(x.shape[1:4], k, strides, padding)
    xc = F.pad(_to_ncdhw(x), (w0, w1, h0, h1, d0, d1))
    w = kernel.permute(4, 3, 0, 1, 2).contiguous()
    y = F.conv3d(xc, w, bias, stride=tuple(strides))
    return _to_ndhwc(y)


def batchnorm(x, gamma, beta, mean, var, eps=BN_EPS):
    """Inference BN, TF non-fused form: x*inv + (beta - mean*inv), inv = gamma*rsqrt(var+eps)."""
    inv = torch.rsqrt(var + eps)
    if gamma is not None:
        inv = inv * gamma
    return x * inv + (beta - mean * inv)


def relu(x):
    return torch.clamp_min(x, 0)


def maxpool3d(x, k, strides, padding="valid"):
    """Padded taps are ignored (-inf padding)."""
    (d0, d1), (h0, h1), (w0, w1) = _pads(x.shape[1:4], k, strides, padding)
    xc = F.pad(_to_ncdhw(x), (w0, w1, h0, h1, d0, d1), value=float("-inf"))
    return _to_ndhwc(F.max_pool3d(xc, tuple(k), tuple(strides)))


def avgpool3d(x, k, strides=(1, 1, 1)):
    return _to_ndhwc(F.avg_pool3d(_to_ncdhw(x), tuple(k), tuple(strides)))


def zeropad3d(x, pads):
    (d0, d1), (h0, h1), (w0, w1) = pads
    return F.pad(x, (0, 0, w0, w1, h0, h1, d0, d1))


def flatten(x):
    """Row-major over (D,H,W,C)."""
    return x.reshape(x.shape[0], -1)


def dense(x, kernel, bias):
    return x @ kernel + bias


def softmax(x):
    return torch.softmax(x, dim=-1)


# --------------------------------------------------------------------------- #
# independent naive numpy restatements (tiny inputs only) used to cross-check
# the torch formulations above
# --------------------------------------------------------------------------- #
def naive_conv3d(x, kernel, bias, strides, padding):
    x = np.asarray(x, np.float64)
    kernel = np.asarray(kernel, np.float64)
    n, d, h, w, c = x.shape
    kd, kh, kw, ci, co = kernel.shape
    pads = _pads((d, h, w), (kd, kh, kw), strides, padding)
    xp = np.zeros((n, d + sum(pads[0]), h + sum(pads[1]), w + sum(pads[2]), c))
    xp[:, pads[0][0]:pads[0][0] + d, pads[1][0]:pads[1][0] + h, pads[2][0]:pads[2][0] + w] = x
    od = (xp.shape[1] - kd) // strides[0] + 1
    oh = (xp.shape[2] - kh) // strides[1] + 1
    ow = (xp.shape[3] - kw) // strides[2] + 1
    y = np.zeros((n, od, oh, ow, co))
    for a in range(od):
        for b in range(oh):
            for e in range(ow):
                win = xp[:, a * strides[0]:a * strides[0] + kd, b * strides[1]:b * strides[1] + kh,
                         e * strides[2]:e * strides[2] + kw, :]
                y[:, a, b, e, :] = np.tensordot(win, kernel, axes=([1, 2, 3, 4], [0, 1, 2, 3]))
    if bias is not None:
        y += np.asarray(bias, np.float64)
    return y


def naive_pool3d(x, k, strides, padding, mode):
    x = np.asarray(x, np.float64)
    n, d, h, w, c = x.shape
    pads = _pads((d, h, w), k, strides, padding)
    fill = -np.inf if mode == "max" else 0.0
    xp = np.full((n, d + sum(pads[0]), h + sum(pads[1]), w + sum(pads[2]), c), fill)
    xp[:, pads[0][0]:pads[0][0] + d, pads[1][0]:pads[1][0] + h, pads[2][0]:pads[2][0] + w] = x
    od = (xp.shape[1] - k[0]) // strides[0] + 1
    oh = (xp.shape[2] - k[1]) // strides[1] + 1
    ow = (xp.shape[3] - k[2]) // strides[2] + 1
    y = np.zeros((n, od, oh, ow, c))
    for a in range(od):
        for b in range(oh):
            for e in range(ow):
                win = xp[:, a * strides[0]:a * strides[0] + k[0], b * strides[1]:b * strides[1] + k[1],
                         e * strides[2]:e * strides[2] + k[2], :]
                y[:, a, b, e, :] = win.max(axis=(1, 2, 3)) if mode == "max" else win.mean(axis=(1, 2, 3))
    return y
