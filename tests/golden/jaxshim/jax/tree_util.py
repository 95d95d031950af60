"""`jax.tree_util` stand-in (sorted dict keys, None = empty subtree)."""
from functools import partial as Partial  # noqa: F401

from ._core import (DictKey, GetAttrKey, SequenceKey, register_pytree_node,  # noqa: F401
                    register_pytree_node_class, tree_flatten, tree_flatten_with_path, tree_leaves,
                    tree_map, tree_structure, tree_unflatten)


def keystr(path):
    return "".join(str(p) for p in path)
