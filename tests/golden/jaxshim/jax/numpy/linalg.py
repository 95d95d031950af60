"""`jax.numpy.linalg` stand-in."""
from .._core import det, eigh, inv, norm, solve  # noqa: F401
