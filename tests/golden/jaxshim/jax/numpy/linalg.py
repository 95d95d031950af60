"""`jax.numpy.linalg` stand-in."""
from .._core import det, inv, norm, solve  # noqa: F401
