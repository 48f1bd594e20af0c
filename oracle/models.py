"""ORACLE (test infrastructure, not product code) - whole-model forward passes of the
reference's member architectures, written directly against /root/reference/train.py
and independently of the product's graph builders (cse_b200/graph.py), so that a
mistake in either shows up as a parity failure.

PARITY UNPINNED (see oracle/ops.py): Keras 2.2.4 / TF 1.15 cannot run here and the
reference has no golden vectors.

Weights are ``{keras_layer_name: [arrays in Keras order]}``.  R3D layers carry Keras
auto-names (``conv3d_7``, ``batch_normalization_3`` ...) generated in *creation order*
in a fresh session; the counters below follow the statement order of
Resnet3DBuilder.build / basic_block / bottleneck / _shortcut3d (train.py:1278-1524).

Every function returns ``(logits, probs)``; the reference models only emit ``probs``
(Dense(..., activation='softmax'), train.py:1268, 840, 1007, 1510-1515).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops

_I3D_BLOCKS = {  # train.py:1041-1193 (b0, b1a, b1b, b2a, b2b, b3)
    "3b": (64, 96, 128, 16, 32, 32), "3c": (128, 128, 192, 32, 96, 64),
    "4b": (192, 96, 208, 16, 48, 64), "4c": (160, 112, 224, 24, 64, 64),
    "4d": (128, 128, 256, 24, 64, 64), "4e": (112, 144, 288, 32, 64, 64),
    "4f": (256, 160, 320, 32, 128, 128),
    "5b": (256, 160, 320, 32, 128, 128), "5c": (384, 192, 384, 48, 128, 128),
}


class _W:
    """dtype-converting view of a weight dict."""

    def __init__(self, weights, dtype):
        self.w, self.dtype = weights, dtype

    def __call__(self, layer):
        return [torch.as_tensor(np.asarray(a), dtype=self.dtype) for a in self.w[layer]]


def _as_input(x, dtype):
    """Network input = raw BGR frames 0..255 cast to float (train.py:257-291, 466-478):
    no crop, no mean/std."""
    return torch.as_tensor(np.asarray(x)).to(dtype)


# --------------------------------------------------------------------------- #
def c3d_forward(weights, clips, dtype=torch.float64, last_name="fc8"):
    """ConvNets3D, train.py:1224-1273."""
    W = _W(weights, dtype)
    x = _as_input(clips, dtype)

    def conv(x, name):
        k, b = W(name)
        return ops.relu(ops.conv3d(x, k, b, (1, 1, 1), "same"))

    x = conv(x, "conv1")
    x = ops.maxpool3d(x, (1, 2, 2), (1, 2, 2), "valid")
    x = conv(x, "conv2")
    x = ops.maxpool3d(x, (2, 2, 2), (2, 2, 2), "valid")
    x = conv(conv(x, "conv3a"), "conv3b")
    x = ops.maxpool3d(x, (2, 2, 2), (2, 2, 2), "valid")
    x = conv(conv(x, "conv4a"), "conv4b")
    x = ops.maxpool3d(x, (2, 2, 2), (2, 2, 2), "valid")
    x = conv(conv(x, "conv5a"), "conv5b")
    x = ops.zeropad3d(x, ((0, 0), (0, 1), (0, 1)))
    x = ops.maxpool3d(x, (2, 2, 2), (2, 2, 2), "valid")
    x = ops.flatten(x)
    x = ops.relu(ops.dense(x, *W("fc6")))
    x = ops.relu(ops.dense(x, *W("fc7")))
    logits = ops.dense(x, *W(last_name))
    return logits, ops.softmax(logits)


# --------------------------------------------------------------------------- #
def _i3d_tower(W, x, ext):
    """Inception_architecture(include_top=False), train.py:1013-1217."""

    def unit(x, name, strides=(1, 1, 1)):      # conv3d_bn, train.py:615-670
        (k,) = W(name + ext + "_conv")
        beta, mean, var = W(name + ext + "_bn")
        y = ops.conv3d(x, k, None, strides, "same")
        return ops.relu(ops.batchnorm(y, None, beta, mean, var))

    x = unit(x, "Conv3d_1a_7x7", (2, 2, 2))
    x = ops.maxpool3d(x, (1, 3, 3), (1, 2, 2), "same")
    x = unit(x, "Conv3d_2b_1x1")
    x = unit(x, "Conv3d_2c_3x3")
    x = ops.maxpool3d(x, (1, 3, 3), (1, 2, 2), "same")
    for tag in ("3b", "3c", "P4", "4b", "4c", "4d", "4e", "4f", "P5", "5b", "5c"):
        if tag == "P4":
            x = ops.maxpool3d(x, (3, 3, 3), (2, 2, 2), "same")
            continue
        if tag == "P5":
            x = ops.maxpool3d(x, (2, 2, 2), (2, 2, 2), "same")
            continue
        b0 = unit(x, "Conv3d_%s_0a_1x1" % tag)
        b1 = unit(unit(x, "Conv3d_%s_1a_1x1" % tag), "Conv3d_%s_1b_3x3" % tag)
        b2 = unit(unit(x, "Conv3d_%s_2a_1x1" % tag), "Conv3d_%s_2b_3x3" % tag)
        b3 = unit(ops.maxpool3d(x, (3, 3, 3), (1, 1, 1), "same"), "Conv3d_%s_3b_1x1" % tag)
        for got, want in zip((b0, b1, b2, b3), (_I3D_BLOCKS[tag][i] for i in (0, 2, 4, 5))):
            assert got.shape[-1] == want, (tag, got.shape, want)
        x = torch.cat([b0, b1, b2, b3], dim=-1)
    h, w = x.shape[2], x.shape[3]
    return ops.avgpool3d(x, (2, h, w), (1, 1, 1))


def i3d_forward(weights, clips, dtype=torch.float64):
    """Inception_Inflated3d(include_top=False, weights=None), train.py:673-843."""
    W = _W(weights, dtype)
    f = ops.flatten(_i3d_tower(W, _as_input(clips, dtype), "_rgb"))
    logits = ops.dense(f, *W("predictions"))
    return logits, ops.softmax(logits)


def twostream_forward(weights, rgb, flow, dtype=torch.float64):
    """TwoStream_Inception_Inflated3d, train.py:857-1011: concatenate([Flatten(rgb),
    Flatten(flow)]) -> Dense softmax 'predictions'; inputs [rgb, flow]."""
    W = _W(weights, dtype)
    fr = ops.flatten(_i3d_tower(W, _as_input(rgb, dtype), "_rgb"))
    ff = ops.flatten(_i3d_tower(W, _as_input(flow, dtype), "_flow"))
    logits = ops.dense(torch.cat([fr, ff], dim=-1), *W("predictions"))
    return logits, ops.softmax(logits)


# --------------------------------------------------------------------------- #
_R3D = {"R3D_18": ("basic", [2, 2, 2, 2]), "R3D_34": ("basic", [3, 4, 6, 3]),
        "R3D_50": ("bottleneck", [3, 4, 6, 3]), "R3D_101": ("bottleneck", [3, 4, 23, 3]),
        "R3D_152": ("bottleneck", [3, 8, 36, 3])}


def r3d_forward(weights, clips, model_type="R3D_34", dtype=torch.float64, emulate_bf16=False):
    """Resnet3DBuilder.build, train.py:1459-1524 (pre-activation ResNet-3D).

    emulate_bf16=True restates the SAME graph with the storage roundings of the bf16 device path
    (conv/dense inputs and kernels rounded to bfloat16, every stored activation rounded once; the
    BN-ReLU that follows a residual add is taken from the un-rounded sum, as the fused epilogue
    does), accumulating in `dtype`.  It separates kernel errors from the bf16 quantisation floor of
    a deep random-weight network (tests/test_gpu_models.py)."""
    W = _W(weights, dtype)
    kind, reps = _R3D[model_type]
    cnt = {"conv3d": 0, "batch_normalization": 0}
    q = (lambda t: t.to(torch.bfloat16).to(dtype)) if emulate_bf16 else (lambda t: t)

    def new(prefix):
        cnt[prefix] += 1
        return "%s_%d" % (prefix, cnt[prefix])

    def conv(x, k, strides, padding="same"):
        name = new("conv3d")
        kern, b = W(name)
        assert tuple(kern.shape[:3]) == tuple(k), (name, kern.shape, k)
        return ops.conv3d(q(x), q(kern), b, strides, padding)

    def bn_relu(x):                                      # _bn_relu, :1278
        g, b, m, v = W(new("batch_normalization"))
        return q(ops.relu(ops.batchnorm(x, g, b, m, v)))

    def bn_relu_conv(x, k, strides=(1, 1, 1)):           # _bn_relu_conv3d, :1303
        a = bn_relu(x)
        return conv(a, k, strides)

    def shortcut(x, residual):                           # _shortcut3d, :1324
        s = tuple(math.ceil(x.shape[1 + i] / residual.shape[1 + i]) for i in range(3))
        sc = x
        if max(s) > 1 or x.shape[-1] != residual.shape[-1]:
            name = new("conv3d")
            kern, b = W(name)
            assert kern.shape[-1] == residual.shape[-1]
            sc = q(ops.conv3d(q(x), q(kern), b, s, "valid"))
        return sc + residual

    x = _as_input(clips, dtype)
    x = conv(x, (7, 7, 7), (2, 2, 2))                    # _conv_bn_relu3D, :1283
    x = bn_relu(x)
    x = ops.maxpool3d(x, (3, 3, 3), (2, 2, 2), "same")
    filters = 64
    exact = x          # the block input before its storage rounding (what the fused BN-ReLU output sees)
    for i, r in enumerate(reps):
        for j in range(r):
            strides = (2, 2, 2) if (j == 0 and i != 0) else (1, 1, 1)
            first = (i == 0 and j == 0)
            if kind == "basic":                          # basic_block, :1368
                c1 = conv(x, (3, 3, 3), strides) if first else bn_relu_conv(exact, (3, 3, 3), strides)
                res = bn_relu_conv(c1, (3, 3, 3))
            else:                                        # bottleneck, :1396
                c1 = conv(x, (1, 1, 1), strides) if first else bn_relu_conv(exact, (1, 1, 1), strides)
                c3 = bn_relu_conv(c1, (3, 3, 3))
                res = bn_relu_conv(c3, (1, 1, 1))
                assert res.shape[-1] == filters * 4
            exact = shortcut(x, res)
            x = q(exact)
        filters *= 2
    x = bn_relu(exact)
    x = q(ops.avgpool3d(x, x.shape[1:4], (1, 1, 1)))
    kern, b = W("dense_1")
    logits = ops.dense(ops.flatten(x), q(kern), b)
    return logits, ops.softmax(logits)


def forward(model_type, weights, inputs, dtype=torch.float64, emulate_bf16=False):
    """inputs: one NDHWC array, or [rgb, flow] for TWOSTREAM_I3D (train.py:1009)."""
    if emulate_bf16 and model_type not in _R3D:
        raise NotImplementedError("bf16 storage emulation is only restated for the R3D family")
    with torch.no_grad():
        if model_type == "C3D":
            last = "fc8" if "fc8" in weights else "predictions"
            return c3d_forward(weights, inputs, dtype, last)
        if model_type == "I3D":
            return i3d_forward(weights, inputs, dtype)
        if model_type == "TWOSTREAM_I3D":
            return twostream_forward(weights, inputs[0], inputs[1], dtype)
        if model_type in _R3D:
            return r3d_forward(weights, inputs, model_type, dtype, emulate_bf16)
    raise ValueError("Unknown model %r" % (model_type,))
