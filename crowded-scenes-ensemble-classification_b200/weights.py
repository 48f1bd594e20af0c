"""Member weights: the Keras-ordered weight dictionary, synthetic generators, loaders.

A *weight set* is ``{layer_name: [np.float32 arrays in Keras order]}`` with
Keras layouts: Conv3D ``[kd,kh,kw,Cin,Cout]`` (+bias ``[Cout]``),
BatchNormalization ``[gamma?, beta, moving_mean, moving_variance]``, Dense
``[in,out]`` + bias.  No weights ship with the reference (evaluate_ensemble.py
reads ``*_weights.hdf5`` written by train.py:1850-1853), so tests and benches
use the seeded generator below (N(0, 2/fan_in) kernels, zero bias,
BN gamma=1, beta=0, mean~N(0,0.1), var~U(0.5,1.5)); an optional non-trivial
bias/gamma/beta variant exercises every epilogue term in the parity tests.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

from .graph import Graph


def synthetic_weights(g: Graph, seed: int = 100, nontrivial: bool = False) -> Dict[str, List[np.ndarray]]:
    rng = np.random.default_rng(seed)
    out: Dict[str, List[np.ndarray]] = {}
    for node in g.nodes.values():
        if not node.weights:
            continue
        arrs = []
        for wname, shp in node.weights:
            kind = wname.rsplit("/", 1)[1].split(":")[0]
            if kind == "kernel":
                fan_in = math.prod(shp[:-1])
                a = rng.standard_normal(shp, dtype=np.float32) * np.float32(math.sqrt(2.0 / fan_in))
            elif kind == "bias":
                a = (rng.standard_normal(shp, dtype=np.float32) * np.float32(0.1)) if nontrivial \
                    else np.zeros(shp, np.float32)
            elif kind == "gamma":
                a = (1.0 + 0.2 * rng.standard_normal(shp)).astype(np.float32) if nontrivial \
                    else np.ones(shp, np.float32)
            elif kind == "beta":
                a = (0.1 * rng.standard_normal(shp)).astype(np.float32) if nontrivial \
                    else np.zeros(shp, np.float32)
            elif kind == "moving_mean":
                a = (0.1 * rng.standard_normal(shp)).astype(np.float32)
            elif kind == "moving_variance":
                a = rng.uniform(0.5, 1.5, shp).astype(np.float32)
            else:
                raise ValueError("unknown weight kind %s" % wname)
            arrs.append(np.ascontiguousarray(a, dtype=np.float32))
        out[node.name] = arrs
    return out


def check_weights(g: Graph, weights: Dict[str, List[np.ndarray]]) -> None:
    """Fail loudly on any missing layer / count / shape mismatch."""
    for node in g.nodes.values():
        if not node.weights:
            continue
        if node.name not in weights:
            raise KeyError("weights for layer %r missing" % node.name)
        arrs = weights[node.name]
        if len(arrs) != len(node.weights):
            raise ValueError("layer %r expects %d weight tensors, got %d"
                             % (node.name, len(node.weights), len(arrs)))
        for (wname, shp), a in zip(node.weights, arrs):
            if tuple(a.shape) != tuple(shp):
                raise ValueError("weight %s expects shape %r, got %r" % (wname, shp, a.shape))


def assign_positional(g: Graph, file_layers: List[List[np.ndarray]]) -> Dict[str, List[np.ndarray]]:
    """``model.load_weights(by_name=False)`` rule: weight-less layers
    dropped on both sides, remaining layers paired by position, tensors by position."""
    mine = g.weighted_layers()
    theirs = [l for l in file_layers if len(l) > 0]
    if len(mine) != len(theirs):
        raise ValueError("file has %d weighted layers, model %s has %d"
                         % (len(theirs), g.name, len(mine)))
    out = {}
    for node, arrs in zip(mine, theirs):
        out[node.name] = [np.ascontiguousarray(a, dtype=np.float32) for a in arrs]
    check_weights(g, out)
    return out
