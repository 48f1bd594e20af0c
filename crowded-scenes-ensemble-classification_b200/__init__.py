"""B200-native ensemble-inference hot path of MounirB/Crowded-scenes-Ensemble-classification.

Host-side Python mirrors the reference's ``evaluate_ensemble.py`` / ``train.py``
surface for the inference path; all arithmetic runs in hand-written sm_100a CUDA
behind the C ABI declared in ``include/cse.h`` (``csrc/`` -> ``libcse_b200.so``).
There is no CPU fallback: importing :mod:`cse_b200.runtime` without the built
library, or calling it without a GPU, raises.
"""
__version__ = "0.1.0"
