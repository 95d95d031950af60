"""`jax.core` stand-in: nothing is ever traced by the eager shim, so no value is a Tracer."""


class Tracer:
    pass
