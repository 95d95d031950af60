"""`jax.lax` stand-in: eager control flow."""
from ._core import cond, custom_linear_solve, fori_loop, scan, stop_gradient, switch, while_loop  # noqa: F401


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    from . import _Missing
    return _Missing(f"jax.lax.{name}")
