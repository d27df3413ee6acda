"""gpflow.config defaults that silently shape the numerics of the reference path:
default_float = float64, default_jitter = 1e-6, positive bijector = softplus, positive_minimum = 0."""
import numpy as np

_JITTER = 1e-6


def default_float():
    return np.float64


def default_int():
    return np.int32


def default_jitter() -> float:
    return _JITTER


def default_positive_minimum() -> float:
    return 0.0
