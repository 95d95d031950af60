"""`jax.flatten_util` stand-in."""
from ._core import ravel_pytree  # noqa: F401
