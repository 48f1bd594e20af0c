"""The model-zoo surface of the reference's evaluation path, backed by libcse_b200.

Same names, argument meaning and order as the reference:

* ``define_input(model_type)``                                         train.py:1566-1616
* ``evaluate_load_model(model_type, model_weights_path, input_shape, nb_classes)``
                                                                       train.py:1712-1772

``evaluate_load_model`` returns a :class:`cse_b200.model.Member` (``compile`` /
``predict_generator`` / ``predict``), built from the architecture graph and the member's Keras
``*_weights.hdf5`` file, which is read positionally exactly like ``model.load_weights(path)``
(by_name=False): weight-less layers dropped, remaining layers paired in ``model.layers`` order.
"""
from __future__ import annotations

import numpy as np

from .graph import MODEL_TYPES, build_model_graph, define_input_shape
from .hdf5 import read_keras_weights
from .weights import assign_positional

# options of the B200 path that the reference has no flag for (process-wide defaults)
DEFAULTS = {"precision": "bf16", "max_batch": 32}


def define_input(model_type):
    """Input prototype (only ``.shape`` is used by callers, evaluate_ensemble.py:1030-1034)."""
    try:
        shape = define_input_shape(model_type)
    except ValueError:
        print("Unknown model")
        raise
    print("# %s sample_input creation :" % ("R3D" if model_type.startswith("R3D") else model_type))
    return np.empty(list(shape), dtype=np.uint8)


def load_member_weights(graph, model_weights_path):
    layer_names, weight_names, arrays = read_keras_weights(model_weights_path)
    return assign_positional(graph, arrays)


def evaluate_load_model(model_type, model_weights_path, input_shape, nb_classes, precision=None, max_batch=None,
                        device=None):
    """Builds the member for `model_type` and loads its weights from `model_weights_path`."""
    from .model import Member
    if model_type not in MODEL_TYPES:
        print("Unknown model")
        raise ValueError("Unknown model %r" % (model_type,))
    print("%s evaluation" % model_type)
    graph = build_model_graph(model_type, tuple(int(v) for v in input_shape), int(nb_classes))
    weights = load_member_weights(graph, model_weights_path)
    return Member(graph, weights, precision=precision or DEFAULTS["precision"],
                  max_batch=int(max_batch or DEFAULTS["max_batch"]), device=device)
