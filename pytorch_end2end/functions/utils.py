"""Host-side helper kept for import compatibility with the reference's tests
(reference: pytorch_end2end/functions/utils.py:6-24, a numba-vectorised two-argument log-sum-exp used by
tests/test_ctc_decoder.py:11,34).  Not on the hot path: plain numpy."""
import numpy as np


def log_sum_exp(a, b):
    """log(exp(a) + exp(b)) with the reference's -inf conventions (src/utils/math_utils.h:8-16)."""
    return np.logaddexp(a, b)
