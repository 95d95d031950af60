"""Opt-in collectives of the path.

The path has exactly three exchange steps: the all-reduce of the objective
``(J, grad[, H])``, the interface exchange of the assembled residual, and the all-reduce of the
FE parameter gradient.  Every one of them is a SUM over the ranks that *partitioned* the units
(points or elements).  A caller that did not partition - the reference-signature single-point
objectives, a replica running the same experiment on every rank - must not be summed, so none of
the library's functions touches ``torch.distributed`` unless the caller passes a group:

* ``group=None``          no collective (the default everywhere);
* ``group=WORLD``         the default process group (must be initialised);
* ``group=<ProcessGroup>`` that group.
"""
from __future__ import annotations

WORLD = "world"


def resolve(group):
    """``(active, process_group_or_None)``: ``active`` is False when no collective is wanted
    (``group is None``) or the group has a single rank."""
    if group is None:
        return False, None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("a process group was requested but torch.distributed is not initialised")
    pg = None if (isinstance(group, str) and group == WORLD) else group
    return dist.get_world_size(pg) > 1, pg


def all_reduce_sum(t, group):
    """In-place SUM over ``group`` (see :func:`resolve`); returns ``t``."""
    active, pg = resolve(group)
    if active:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
    return t
