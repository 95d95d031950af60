"""`jax.experimental` stand-in (see `jax/_core.py`)."""
from . import sparse  # noqa: F401
