"""Portable LCG initial-condition generator (reference chsimpy/mport.py:8-32).

The recurrence x <- (a*x + c) mod 2^31 is evaluated in float64 in the reference, and
a*x exceeds 2^53, so the rounding of that product is part of the specification.  This
version keeps the float64 arithmetic but runs the serial recurrence on plain Python
floats (identical IEEE operations, ~10x faster than the numpy-scalar generator) and
fills the matrix column-major in one reshape."""
import math

import numpy as np

_A = 1103515245.0
_C = 12345.0
_M = 2147483648.0


def matlab_lcg_sample(n1, n2, seed):
    """n1 x n2 matrix of pseudo-random values in [0,1), filled column by column."""
    x = float(seed)
    seq = np.empty(n1 * n2)
    for i in range(n1 * n2):
        x = math.fmod(_A * x + _C, _M)
        seq[i] = x
    return seq.reshape(n2, n1).T / (_M - 1.0)
