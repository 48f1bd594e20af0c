"""Declarative layer graph for the ensemble members (C3D, I3D, TwoStream-I3D, R3D-18..152).

One node per Keras layer of the reference model (including weight-less ones:
InputLayer, Activation, Dropout, Flatten, ...), because

* the tests compare it with the independently written CPU oracle (``oracle/models.py``; traces, FLOPs, parameters),
* the device lowering (``lowering.py``) pattern-matches it into fused kernels,
* Keras' ``model.layers`` order (needed to read ``*_weights.hdf5`` positionally,
  the way ``model.load_weights`` of the reference does, train.py:1731-1769)
  depends on every layer, weight-less or not.

The builders restate, layer by layer, the reference's model code:

* ``build_c3d``        <- ConvNets3D                      train.py:1224-1273
* ``build_i3d``        <- conv3d_bn / Inception_architecture / Inception_Inflated3d
                          (include_top=False, weights=None path)
                          train.py:615-670, 1013-1219, 837-841
* ``build_twostream``  <- TwoStream_Inception_Inflated3d  train.py:857-1011 (999-1009)
* ``build_r3d``        <- Resnet3DBuilder.build + basic_block / bottleneck / _shortcut3d
                          train.py:1278-1559
* ``define_input_shape`` <- define_input                  train.py:1566-1616

Library semantics (Keras 2.2.4 / TF 1.15, channels_last) are listed in SURVEY.md
App. A.0; ``same_pads`` below is TF's SAME rule (extra padding at the end).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon


# --------------------------------------------------------------------------- #
# shape helpers
# --------------------------------------------------------------------------- #
def same_pads(size: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF 'SAME': out=ceil(in/s); total=max((out-1)*s+k-in,0); before=total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    before = total // 2
    return out, before, total - before


def valid_out(size: int, k: int, s: int) -> int:
    return (size - k) // s + 1


def window_geometry(in_dhw, k, s, padding):
    """-> (out_dhw, pads_before, pads_after) for a conv/pool window."""
    out, pb, pa = [], [], []
    for i in range(3):
        if padding == "same":
            o, b, a = same_pads(in_dhw[i], k[i], s[i])
        elif padding == "valid":
            o, b, a = valid_out(in_dhw[i], k[i], s[i]), 0, 0
        else:
            raise ValueError("padding must be 'same' or 'valid', got %r" % (padding,))
        if o <= 0:
            raise ValueError("window %r/%r does not fit input %r" % (k, s, in_dhw))
        out.append(o), pb.append(b), pa.append(a)
    return tuple(out), tuple(pb), tuple(pa)


# --------------------------------------------------------------------------- #
# IR
# --------------------------------------------------------------------------- #
@dataclass
class Node:
    name: str
    op: str                      # input|conv3d|bn|relu|dropout|add|concat|maxpool|avgpool|zeropad|flatten|dense
    inputs: List[str]
    attrs: dict = field(default_factory=dict)
    out_shape: Tuple[int, ...] = ()      # (D,H,W,C) or (F,)
    weights: List[Tuple[str, Tuple[int, ...]]] = field(default_factory=list)  # Keras order


class Graph:
    """Insertion-ordered DAG of Nodes.  ``kind`` is 'sequential' or 'functional'."""

    def __init__(self, name: str, kind: str):
        self.name = name
        self.kind = kind
        self.nodes: "OrderedDict[str, Node]" = OrderedDict()
        self.inputs: List[str] = []
        self.output: Optional[str] = None
        self._uid: Dict[str, int] = {}

    # Keras auto-naming: <snake_case_class>_<n>, counters per prefix, creation order
    def auto(self, prefix: str) -> str:
        self._uid[prefix] = self._uid.get(prefix, 0) + 1
        return "%s_%d" % (prefix, self._uid[prefix])

    def add(self, node: Node) -> str:
        if node.name in self.nodes:
            raise ValueError("duplicate layer name %s" % node.name)
        self.nodes[node.name] = node
        return node.name

    def shape(self, name: str):
        return self.nodes[name].out_shape

    # ---- layer constructors (each returns the new node's name) ----------- #
    def input(self, shape, name=None):
        name = name or self.auto("input")
        self.add(Node(name, "input", [], {}, tuple(shape)))
        self.inputs.append(name)
        return name

    def conv3d(self, x, filters, k, strides=(1, 1, 1), padding="same", use_bias=True,
               activation=None, name=None):
        name = name or self.auto("conv3d")
        d, h, w, c = self.shape(x)
        out, pb, pa = window_geometry((d, h, w), k, strides, padding)
        wts = [(name + "/kernel:0", (k[0], k[1], k[2], c, filters))]
        if use_bias:
            wts.append((name + "/bias:0", (filters,)))
        self.add(Node(name, "conv3d", [x],
                      dict(filters=filters, k=tuple(k), s=tuple(strides), padding=padding,
                           pads_before=pb, pads_after=pa, use_bias=use_bias, activation=activation),
                      out + (filters,), wts))
        return name

    def bn(self, x, scale=True, name=None):
        name = name or self.auto("batch_normalization")
        c = self.shape(x)[-1]
        wts = []
        if scale:
            wts.append((name + "/gamma:0", (c,)))
        wts += [(name + "/beta:0", (c,)), (name + "/moving_mean:0", (c,)),
                (name + "/moving_variance:0", (c,))]
        self.add(Node(name, "bn", [x], dict(scale=scale, eps=BN_EPS), self.shape(x), wts))
        return name

    def relu(self, x, name=None):
        name = name or self.auto("activation")
        self.add(Node(name, "relu", [x], {}, self.shape(x)))
        return name

    def dropout(self, x, name=None):
        name = name or self.auto("dropout")
        self.add(Node(name, "dropout", [x], {}, self.shape(x)))
        return name

    def add_(self, xs, name=None):
        name = name or self.auto("add")
        shp = self.shape(xs[0])
        for x in xs[1:]:
            if self.shape(x) != shp:
                raise ValueError("add: shape mismatch %r vs %r" % (self.shape(x), shp))
        self.add(Node(name, "add", list(xs), {}, shp))
        return name

    def concat(self, xs, name=None):
        name = name or self.auto("concatenate")
        shp = self.shape(xs[0])
        ctot = 0
        for x in xs:
            if self.shape(x)[:-1] != shp[:-1]:
                raise ValueError("concat: shape mismatch")
            ctot += self.shape(x)[-1]
        self.add(Node(name, "concat", list(xs), {}, shp[:-1] + (ctot,)))
        return name

    def _pool(self, op, prefix, x, k, strides, padding, name):
        name = name or self.auto(prefix)
        d, h, w, c = self.shape(x)
        out, pb, pa = window_geometry((d, h, w), k, strides, padding)
        self.add(Node(name, op, [x], dict(k=tuple(k), s=tuple(strides), padding=padding,
                                          pads_before=pb, pads_after=pa), out + (c,)))
        return name

    def maxpool(self, x, k, strides, padding="valid", name=None):
        return self._pool("maxpool", "max_pooling3d", x, k, strides, padding, name)

    def avgpool(self, x, k, strides=(1, 1, 1), padding="valid", name=None):
        if padding != "valid":
            raise ValueError("AveragePooling3D is only used with 'valid' in the reference")
        return self._pool("avgpool", "average_pooling3d", x, k, strides, padding, name)

    def zeropad(self, x, pads, name=None):
        """pads = ((d0,d1),(h0,h1),(w0,w1)) as in keras ZeroPadding3D."""
        name = name or self.auto("zero_padding3d")
        d, h, w, c = self.shape(x)
        (d0, d1), (h0, h1), (w0, w1) = pads
        self.add(Node(name, "zeropad", [x], dict(pads=tuple(map(tuple, pads))),
                      (d + d0 + d1, h + h0 + h1, w + w0 + w1, c)))
        return name

    def flatten(self, x, name=None):
        name = name or self.auto("flatten")
        n = 1
        for v in self.shape(x):
            n *= v
        self.add(Node(name, "flatten", [x], {}, (n,)))
        return name

    def dense(self, x, units, activation=None, name=None):
        name = name or self.auto("dense")
        (fin,) = self.shape(x)
        self.add(Node(name, "dense", [x], dict(units=units, activation=activation), (units,),
                      [(name + "/kernel:0", (fin, units)), (name + "/bias:0", (units,))]))
        return name

    # ---- queries ---------------------------------------------------------- #
    def weighted_layers(self) -> List[Node]:
        """Layers that own weights, in Keras ``model.layers`` order."""
        return [self.nodes[n] for n in self.keras_layer_order() if self.nodes[n].weights]

    def keras_layer_order(self) -> List[str]:
        """``model.layers`` order.

        Sequential: insertion order.  Functional (keras/engine/network.py,
        2.2.4): DFS from the outputs assigns ``layer_index`` in first-visit
        order (inbound layers in call order); depth = longest distance to an
        output; layers sorted by depth descending, ties by ``layer_index``.
        """
        if self.kind == "sequential":
            return list(self.nodes)
        layer_index: Dict[str, int] = {}
        post: List[str] = []
        seen = set()
        # iterative DFS replicating build_map's pre-order index / post-order node list
        stack = [(self.output, 0)]
        while stack:
            name, i = stack.pop()
            node = self.nodes[name]
            if i == 0:
                if name in seen:
                    continue
                if name not in layer_index:
                    layer_index[name] = len(layer_index)
            if i < len(node.inputs):
                stack.append((name, i + 1))
                nxt = node.inputs[i]
                if nxt not in seen:
                    stack.append((nxt, 0))
            else:
                if name not in seen:
                    seen.add(name)
                    post.append(name)
        depth: Dict[str, int] = {}
        for name in reversed(post):
            d = depth.setdefault(name, 0)
            for inp in self.nodes[name].inputs:
                depth[inp] = max(depth.get(inp, 0), d + 1)
        order = sorted(post, key=lambda n: (-depth[n], layer_index[n]))
        return order

    def conv_dense_flops(self) -> Dict[str, float]:
        """2*MACs per clip for every conv3d / dense node at un-padded shapes."""
        out = {}
        for n in self.nodes.values():
            if n.op == "conv3d":
                d, h, w, co = n.out_shape
                ci = self.shape(n.inputs[0])[-1]
                k = n.attrs["k"]
                out[n.name] = 2.0 * d * h * w * co * ci * k[0] * k[1] * k[2]
            elif n.op == "dense":
                (fin,) = self.shape(n.inputs[0])
                out[n.name] = 2.0 * fin * n.attrs["units"]
        return out

    def total_flops(self) -> float:
        return sum(self.conv_dense_flops().values())

    def param_count(self) -> int:
        tot = 0
        for n in self.nodes.values():
            for _, shp in n.weights:
                tot += math.prod(shp)
        return tot


# --------------------------------------------------------------------------- #
# input prototypes                                     (train.py:1566-1616)
# --------------------------------------------------------------------------- #
R3D_REPETITIONS = {          # train.py:1527-1559
    "R3D_18": ("basic_block", [2, 2, 2, 2]),
    "R3D_34": ("basic_block", [3, 4, 6, 3]),
    "R3D_50": ("bottleneck", [3, 4, 6, 3]),
    "R3D_101": ("bottleneck", [3, 4, 23, 3]),
    "R3D_152": ("bottleneck", [3, 8, 36, 3]),
}
MODEL_TYPES = ("C3D", "I3D", "TWOSTREAM_I3D") + tuple(R3D_REPETITIONS)


def define_input_shape(model_type: str) -> Tuple[int, int, int, int]:
    """(T,H,W,C) prototype per architecture.  TWOSTREAM has C=0 (rewritten to 3 / 2
    per tower, train.py:880, 891)."""
    if model_type == "I3D":
        return (20, 224, 224, 3)
    if model_type == "TWOSTREAM_I3D":
        return (20, 224, 224, 0)
    if model_type == "C3D" or model_type in R3D_REPETITIONS:
        return (16, 112, 112, 3)
    raise ValueError("Unknown model %r" % (model_type,))


# --------------------------------------------------------------------------- #
# C3D                                                  (train.py:1224-1273)
# --------------------------------------------------------------------------- #
def build_c3d(input_shape=(16, 112, 112, 3), num_classes=11, last_name="fc8") -> Graph:
    g = Graph("C3D", "sequential")
    x = g.input(input_shape, name="conv1_input")
    k3, s1 = (3, 3, 3), (1, 1, 1)
    x = g.conv3d(x, 64, k3, s1, "same", True, "relu", name="conv1")
    x = g.maxpool(x, (1, 2, 2), (1, 2, 2), "valid", name="pool1")
    x = g.conv3d(x, 128, k3, s1, "same", True, "relu", name="conv2")
    x = g.maxpool(x, (2, 2, 2), (2, 2, 2), "valid", name="pool2")
    x = g.conv3d(x, 256, k3, s1, "same", True, "relu", name="conv3a")
    x = g.conv3d(x, 256, k3, s1, "same", True, "relu", name="conv3b")
    x = g.maxpool(x, (2, 2, 2), (2, 2, 2), "valid", name="pool3")
    x = g.conv3d(x, 512, k3, s1, "same", True, "relu", name="conv4a")
    x = g.conv3d(x, 512, k3, s1, "same", True, "relu", name="conv4b")
    x = g.maxpool(x, (2, 2, 2), (2, 2, 2), "valid", name="pool4")
    x = g.conv3d(x, 512, k3, s1, "same", True, "relu", name="conv5a")
    x = g.conv3d(x, 512, k3, s1, "same", True, "relu", name="conv5b")
    x = g.zeropad(x, ((0, 0), (0, 1), (0, 1)), name="zeropad5")
    x = g.maxpool(x, (2, 2, 2), (2, 2, 2), "valid", name="pool5")
    x = g.flatten(x)
    x = g.dense(x, 4096, "relu", name="fc6")
    x = g.dropout(x)
    x = g.dense(x, 4096, "relu", name="fc7")
    x = g.dropout(x)
    x = g.dense(x, num_classes, "softmax", name=last_name)
    g.output = x
    return g


# --------------------------------------------------------------------------- #
# I3D                                     (train.py:615-670, 1013-1219)
# --------------------------------------------------------------------------- #
# (b0, b1a, b1b, b2a, b2b, b3) per Mixed block, in order
_I3D_MIXED = [
    ("3b", (64, 96, 128, 16, 32, 32)), ("3c", (128, 128, 192, 32, 96, 64)),
    ("POOL4a", None),
    ("4b", (192, 96, 208, 16, 48, 64)), ("4c", (160, 112, 224, 24, 64, 64)),
    ("4d", (128, 128, 256, 24, 64, 64)), ("4e", (112, 144, 288, 32, 64, 64)),
    ("4f", (256, 160, 320, 32, 128, 128)),
    ("POOL5a", None),
    ("5b", (256, 160, 320, 32, 128, 128)), ("5c", (384, 192, 384, 48, 128, 128)),
]


def _conv3d_bn(g: Graph, x, filters, k, strides=(1, 1, 1), name=None):
    """conv3d_bn: Conv3D(no bias,'same') '<name>_conv' -> BN(scale=False) '<name>_bn'
    -> ReLU '<name>'   (train.py:646-668)."""
    x = g.conv3d(x, filters, k, strides, "same", False, None, name=name + "_conv")
    x = g.bn(x, scale=False, name=name + "_bn")
    return g.relu(x, name=name)


def _i3d_tower(g: Graph, x, ext: str):
    """Inception_architecture, include_top=False branch (train.py:1013-1217)."""
    k1, k3, k7 = (1, 1, 1), (3, 3, 3), (7, 7, 7)
    s1 = (1, 1, 1)
    x = _conv3d_bn(g, x, 64, k7, (2, 2, 2), "Conv3d_1a_7x7" + ext)
    x = g.maxpool(x, (1, 3, 3), (1, 2, 2), "same", name="MaxPool2d_2a_3x3" + ext)
    x = _conv3d_bn(g, x, 64, k1, s1, "Conv3d_2b_1x1" + ext)
    x = _conv3d_bn(g, x, 192, k3, s1, "Conv3d_2c_3x3" + ext)
    x = g.maxpool(x, (1, 3, 3), (1, 2, 2), "same", name="MaxPool2d_3a_3x3" + ext)
    for tag, f in _I3D_MIXED:
        if tag == "POOL4a":
            x = g.maxpool(x, (3, 3, 3), (2, 2, 2), "same", name="MaxPool2d_4a_3x3" + ext)
            continue
        if tag == "POOL5a":
            x = g.maxpool(x, (2, 2, 2), (2, 2, 2), "same", name="MaxPool2d_5a_2x2" + ext)
            continue
        b0 = _conv3d_bn(g, x, f[0], k1, s1, "Conv3d_%s_0a_1x1%s" % (tag, ext))
        b1 = _conv3d_bn(g, x, f[1], k1, s1, "Conv3d_%s_1a_1x1%s" % (tag, ext))
        b1 = _conv3d_bn(g, b1, f[2], k3, s1, "Conv3d_%s_1b_3x3%s" % (tag, ext))
        b2 = _conv3d_bn(g, x, f[3], k1, s1, "Conv3d_%s_2a_1x1%s" % (tag, ext))
        b2 = _conv3d_bn(g, b2, f[4], k3, s1, "Conv3d_%s_2b_3x3%s" % (tag, ext))
        b3 = g.maxpool(x, k3, s1, "same", name="MaxPool2d_%s_3a_3x3%s" % (tag, ext))
        b3 = _conv3d_bn(g, b3, f[5], k1, s1, "Conv3d_%s_3b_1x1%s" % (tag, ext))
        x = g.concat([b0, b1, b2, b3], name="Mixed_%s%s" % (tag, ext))
    d, h, w, c = g.shape(x)
    x = g.avgpool(x, (2, h, w), (1, 1, 1), "valid", name="global_avg_pool" + ext)
    return x


def build_i3d(input_shape=(20, 224, 224, 3), num_classes=11) -> Graph:
    """Inception_Inflated3d(include_top=False, weights=None): tower (always the
    '_rgb' suffix, type='rgb' hard-coded, train.py:766) -> Flatten -> Dense softmax
    'predictions' (train.py:837-841)."""
    g = Graph("I3D", "functional")
    x = g.input(input_shape, name="input_1")
    x = _i3d_tower(g, x, "_rgb")
    x = g.flatten(x)
    g.output = g.dense(x, num_classes, "softmax", name="predictions")
    return g


def build_twostream(input_shape=(20, 224, 224, 0), num_classes=11) -> Graph:
    """TwoStream_Inception_Inflated3d: flow tower is constructed first (train.py:919),
    rgb second (:927); features are concatenated [rgb, flow] (:1006) and fed to one
    Dense softmax 'predictions'; model inputs are [rgb, flow] (:1009)."""
    t, h, w, _ = input_shape
    g = Graph("TWOSTREAM_I3D", "functional")
    rgb_in = g.input((t, h, w, 3), name="input_1")    # Input()s: rgb first (train.py:902-903)
    flow_in = g.input((t, h, w, 2), name="input_2")
    y = _i3d_tower(g, flow_in, "_flow")
    x = _i3d_tower(g, rgb_in, "_rgb")
    x = g.flatten(x)
    y = g.flatten(y)
    z = g.concat([x, y])
    g.output = g.dense(z, num_classes, "softmax", name="predictions")
    g.inputs = [rgb_in, flow_in]          # Model(input=[rgb, flow])
    return g


# --------------------------------------------------------------------------- #
# R3D (pre-activation ResNet-3D)                       (train.py:1278-1559)
# --------------------------------------------------------------------------- #
def _bn_relu(g: Graph, x):
    return g.relu(g.bn(x, scale=True))


def _bn_relu_conv(g: Graph, x, filters, k, strides=(1, 1, 1)):
    a = _bn_relu(g, x)
    return g.conv3d(a, filters, k, strides, "same", True, None)


def _shortcut(g: Graph, x, residual):
    """_shortcut3d (train.py:1324-1346): 1x1x1 'valid' strided conv on the RAW block
    input when spatial dims or channels differ; then add([shortcut, residual])."""
    xi, xr = g.shape(x), g.shape(residual)
    strides = tuple(math.ceil(xi[i] / xr[i]) for i in range(3))
    sc = x
    if any(s > 1 for s in strides) or xi[3] != xr[3]:
        sc = g.conv3d(x, xr[3], (1, 1, 1), strides, "valid", True, None)
    return g.add_([sc, residual])


def _basic_block(g, x, filters, strides, first_of_first):
    k3 = (3, 3, 3)
    if first_of_first:
        c1 = g.conv3d(x, filters, k3, strides, "same", True, None)
    else:
        c1 = _bn_relu_conv(g, x, filters, k3, strides)
    res = _bn_relu_conv(g, c1, filters, k3)
    return _shortcut(g, x, res)


def _bottleneck(g, x, filters, strides, first_of_first):
    k1, k3 = (1, 1, 1), (3, 3, 3)
    if first_of_first:
        c1 = g.conv3d(x, filters, k1, strides, "same", True, None)
    else:
        c1 = _bn_relu_conv(g, x, filters, k1, strides)
    c3 = _bn_relu_conv(g, c1, filters, k3)
    res = _bn_relu_conv(g, c3, filters * 4, k1)
    return _shortcut(g, x, res)


def build_r3d(model_type="R3D_34", input_shape=(16, 112, 112, 3), num_classes=11) -> Graph:
    block, reps = R3D_REPETITIONS[model_type]
    fn = _basic_block if block == "basic_block" else _bottleneck
    g = Graph(model_type, "functional")
    x = g.input(input_shape, name="input_1")
    x = g.conv3d(x, 64, (7, 7, 7), (2, 2, 2), "same", True, None)      # _conv_bn_relu3D
    x = _bn_relu(g, x)
    x = g.maxpool(x, (3, 3, 3), (2, 2, 2), "same")
    filters = 64
    for i, r in enumerate(reps):
        for j in range(r):
            strides = (2, 2, 2) if (j == 0 and i != 0) else (1, 1, 1)
            x = fn(g, x, filters, strides, first_of_first=(i == 0 and j == 0))
        filters *= 2
    block_shape = g.shape(x)
    x = _bn_relu(g, x)
    x = g.avgpool(x, block_shape[:3], (1, 1, 1), "valid")
    x = g.flatten(x)
    if num_classes > 1:
        g.output = g.dense(x, num_classes, "softmax")
    else:
        g.output = g.dense(x, num_classes, "sigmoid")
    return g


def build_model_graph(model_type: str, input_shape=None, num_classes=11) -> Graph:
    """Dispatch of evaluate_load_model (train.py:1712-1772), graph part only."""
    if input_shape is None:
        input_shape = define_input_shape(model_type)
    input_shape = tuple(int(v) for v in input_shape)
    if model_type == "C3D":
        return build_c3d(input_shape, num_classes)
    if model_type == "I3D":
        return build_i3d(input_shape, num_classes)
    if model_type == "TWOSTREAM_I3D":
        return build_twostream(input_shape, num_classes)
    if model_type in R3D_REPETITIONS:
        return build_r3d(model_type, input_shape, num_classes)
    raise ValueError("Unknown model %r" % (model_type,))
