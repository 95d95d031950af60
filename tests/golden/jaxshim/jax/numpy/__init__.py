"""`jax.numpy` stand-in (see `jax/_core.py`)."""
import numpy as _np

from .._core import (Array as ndarray, abs_ as abs, add, all_ as all, allclose, any_ as any, arange,  # noqa: F401
                     arccos, arctan, argmax, argmin, argsort, array, asarray, atleast_1d, atleast_2d, broadcast_to,
                     c_, cbrt, clip, concatenate, cos, cross, cumsum, diag, dot, einsum, exp,
                     expand_dims, eye, flip, floor, full, hstack, isclose, isfinite, isnan, linspace, log,
                     logical_and, logical_not, logical_or, matmul, maximum, mean, minimum, moveaxis,
                     ones, ones_like, outer, power, prod, r_, ravel, repeat, reshape, roll, setdiff1d, sign,
                     sin, sort, sqrt, squeeze, stack, sum_ as sum, swapaxes, take, tanh, tensordot, tile,
                     trace, transpose, tril, triu, unique, vstack, where, zeros, zeros_like)
from .._core import div as divide, mul as multiply, sub as subtract  # noqa: F401
from . import linalg  # noqa: F401

absolute = abs
inf = _np.inf
pi = _np.pi
nan = _np.nan
newaxis = None
float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
bool_ = _np.bool_
complex128 = _np.complex128
floating = _np.floating
integer = _np.integer


def square(x):
    return x * x


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    from .. import _Missing
    return _Missing(f"jax.numpy.{name}")
