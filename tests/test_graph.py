"""Graph builders vs the reference's architecture facts."""
import math

import numpy as np
import pytest

from cse_b200 import graph as G
from cse_b200.weights import synthetic_weights, check_weights, assign_positional


def test_same_padding_worked_examples():
    # SURVEY App. A.0 worked examples (TF SAME: extra pad at the end)
    assert G.same_pads(224, 7, 2) == (112, 2, 3)
    assert G.same_pads(112, 3, 2) == (56, 0, 1)
    assert G.same_pads(28, 3, 2) == (14, 0, 1)
    assert G.same_pads(7, 3, 2) == (4, 1, 1)
    assert G.same_pads(1, 3, 2) == (1, 1, 1)
    assert G.same_pads(5, 2, 2) == (3, 0, 1)
    assert G.same_pads(4, 3, 2) == (2, 0, 1)
    assert G.same_pads(9, 3, 1) == (9, 1, 1)
    assert G.same_pads(9, 1, 1) == (9, 0, 0)


def test_define_input_shapes():
    # train.py:1566-1616
    assert G.define_input_shape("C3D") == (16, 112, 112, 3)
    assert G.define_input_shape("I3D") == (20, 224, 224, 3)
    assert G.define_input_shape("TWOSTREAM_I3D") == (20, 224, 224, 0)
    for m in ("R3D_18", "R3D_34", "R3D_50", "R3D_101", "R3D_152"):
        assert G.define_input_shape(m) == (16, 112, 112, 3)
    with pytest.raises(ValueError):
        G.define_input_shape("VGG")


def test_c3d_trace_and_flops():
    g = G.build_c3d()
    s = g.shape
    assert s("conv1") == (16, 112, 112, 64) and s("pool1") == (16, 56, 56, 64)
    assert s("pool2") == (8, 28, 28, 128) and s("pool3") == (4, 14, 14, 256)
    assert s("pool4") == (2, 7, 7, 512) and s("zeropad5") == (2, 8, 8, 512)
    assert s("pool5") == (1, 4, 4, 512) and s("flatten_1") == (8192,)
    assert s("fc8") == (11,)
    assert abs(g.total_flops() / 1e9 - 77.094) < 1e-3
    assert abs(g.param_count() / 1e6 - 78.04) < 0.01
    assert g.keras_layer_order()[1] == "conv1"


@pytest.mark.parametrize("t,gf,feat", [(64, 222.301, 7168), (20, 69.702, 2048)])
def test_i3d_trace_and_flops(t, gf, feat):
    g = G.build_i3d((t, 224, 224, 3))
    assert g.shape("Conv3d_1a_7x7_rgb")[:3] == (t // 2, 112, 112)
    assert g.shape("Mixed_3b_rgb")[-1] == 256 and g.shape("Mixed_3c_rgb")[-1] == 480
    assert g.shape("Mixed_4f_rgb")[-1] == 832 and g.shape("Mixed_5c_rgb")[-1] == 1024
    assert g.shape("flatten_1") == (feat,)
    assert abs(g.total_flops() / 1e9 - gf) < 2e-3
    nconv = sum(1 for n in g.nodes.values() if n.op == "conv3d")
    assert nconv == 57


def test_twostream():
    g = G.build_twostream((64, 224, 224, 0))
    assert g.inputs == ["input_1", "input_2"]
    assert g.shape("input_1")[-1] == 3 and g.shape("input_2")[-1] == 2
    assert g.shape("concatenate_1") == (14336,)
    assert abs(g.total_flops() / 1e9 - 426.978) < 3e-3
    g20 = G.build_twostream((20, 224, 224, 0))
    assert g20.shape("concatenate_1") == (4096,)
    assert abs(g20.total_flops() / 1e9 - 133.897) < 2e-3
    # BASELINE.md's 24.57 M leaves out the 3*7280 BN tensors per tower
    assert abs((g20.param_count() - 2 * 3 * 7280) / 1e6 - 24.57) < 0.01
    # rgb layers come before flow layers at equal depth (rgb is the first inbound
    # of the feature concatenate, train.py:1006)
    order = g20.keras_layer_order()
    assert order.index("Conv3d_1a_7x7_rgb_conv") < order.index("Conv3d_1a_7x7_flow_conv")
    assert order[-1] == "predictions"


@pytest.mark.parametrize("mt,gf,mp", [("R3D_18", 8.706, 33.20), ("R3D_34", 13.321, 63.51),
                                      ("R3D_50", 10.221, 46.19), ("R3D_101", 14.043, 85.21),
                                      ("R3D_152", 18.763, 117.35)])
def test_r3d_flops_params(mt, gf, mp):
    g = G.build_r3d(mt)
    assert abs(g.total_flops() / 1e9 - gf) < 2e-3
    # BASELINE.md's counts leave out the BatchNormalization tensors
    bn = sum(n.out_shape[-1] * len(n.weights) for n in g.nodes.values() if n.op == "bn")
    assert abs((g.param_count() - bn) / 1e6 - mp) < 0.01
    assert g.shape(g.output) == (11,)


def test_r3d34_trace_and_names():
    g = G.build_r3d("R3D_34")
    assert g.shape("conv3d_1") == (8, 56, 56, 64)
    assert g.shape("max_pooling3d_1") == (4, 28, 28, 64)
    # first block of first layer: conv3d_2 consumes the pool directly (no BN-ReLU)
    assert g.nodes["conv3d_2"].inputs == ["max_pooling3d_1"]
    # last stage works on 1x4x4, shortcut stride (1,2,2) (ceil(1/1), ceil(7/4))
    last_sc = [n for n in g.nodes.values() if n.op == "conv3d" and n.attrs["k"] == (1, 1, 1)][-1]
    assert last_sc.attrs["s"] == (1, 2, 2) and last_sc.out_shape == (1, 4, 4, 512)
    assert g.shape("average_pooling3d_1") == (1, 1, 1, 512)
    order = g.keras_layer_order()
    assert order[0] == "input_1" and order[1] == "conv3d_1" and order[-1] == "dense_1"
    assert len(order) == len(g.nodes)


def test_keras_order_is_topological():
    for g in (G.build_r3d("R3D_18"), G.build_i3d(), G.build_twostream()):
        pos = {n: i for i, n in enumerate(g.keras_layer_order())}
        for n in g.nodes.values():
            for inp in n.inputs:
                assert pos[inp] < pos[n.name]


def test_weights_positional_assignment_roundtrip():
    g = G.build_r3d("R3D_18")
    w = synthetic_weights(g, seed=3)
    check_weights(g, w)
    file_layers = [w.get(n, []) for n in g.keras_layer_order()]
    w2 = assign_positional(g, file_layers)
    assert set(w2) == set(w)
    for k in w:
        for a, b in zip(w[k], w2[k]):
            assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        assign_positional(g, file_layers[:-1])
