"""Parameter pytrees for the B200 path - host-side mirror of the reference's
``cmad.parameters.parameters.Parameters`` (parameters.py:177-401) without JAX.

Same public names and semantics: ``values`` / active flags / transforms are
parallel nested dicts; the flat order is the JAX dict-pytree order (keys sorted,
depth first; array leaves row-major); ``active_idx`` selects the active flat
entries; canonical<->native transforms are ``bounds`` (two numbers) and ``log``
(one reference value).
"""
from __future__ import annotations

import math
from typing import Any, Callable

import numpy as np


# ---- minimal dict-pytree helpers (sorted-key order, None is a leaf) --------
def tree_leaves_with_path(tree: Any, path: tuple = ()) -> list[tuple[tuple, Any]]:
    if isinstance(tree, dict):
        out: list = []
        for key in sorted(tree):
            out.extend(tree_leaves_with_path(tree[key], path + (key,)))
        return out
    return [(path, tree)]


def tree_leaves(tree: Any) -> list:
    return [leaf for _, leaf in tree_leaves_with_path(tree)]


def tree_map(fn: Callable, tree: Any, *others: Any) -> Any:
    if isinstance(tree, dict):
        return {k: tree_map(fn, tree[k], *(o[k] for o in others)) for k in tree}
    return fn(tree, *others)


def tree_unflatten_like(like: Any, flat: np.ndarray) -> Any:
    """Rebuild a pytree shaped like ``like`` from a flat vector."""
    pos = 0

    def go(node):
        nonlocal pos
        if isinstance(node, dict):
            return {k: go(node[k]) for k in sorted(node)}
        arr = np.asarray(node)
        n = int(arr.size)
        seg = np.asarray(flat[pos:pos + n])
        pos += n
        return seg.reshape(arr.shape).copy() if arr.ndim else seg[0]

    return go(like)


def ravel_pytree(tree: Any) -> np.ndarray:
    leaves = tree_leaves(tree)
    if not leaves:
        return np.zeros(0)
    return np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in leaves])


# ---- transforms (parameters.py:27-54, 90-166) ------------------------------
def bounds_transform(value, bounds, transform_from_canonical=True):
    span = 0.5 * (bounds[1] - bounds[0])
    mean = 0.5 * (bounds[0] + bounds[1])
    if transform_from_canonical:
        return span * value + mean
    return min(1.0, max(-1.0, (value - mean) / span))


def log_transform(value, ref_value, transform_from_canonical=True):
    if transform_from_canonical:
        return ref_value[0] * math.exp(value)
    return math.log(value / ref_value[0])


def first_deriv_transform(value, transform):
    if transform is None:
        return 1.0
    if len(transform) == 2:
        return 0.5 * (transform[1] - transform[0])
    if len(transform) == 1:
        return value
    raise ValueError(f"Unexpected transform shape: {transform}")


def second_deriv_transform(value, transform):
    if transform is None or len(transform) == 2:
        return 0.0
    if len(transform) == 1:
        return value
    raise ValueError(f"Unexpected transform shape: {transform}")


def transform_from_canonical(value, active_flag, transform):
    if active_flag and transform is not None:
        if len(transform) == 2:
            return bounds_transform(value, transform)
        if len(transform) == 1:
            return log_transform(value, transform)
        raise ValueError(f"Unexpected transform shape: {transform}")
    return value


def transform_to_canonical(value, active_flag, transform):
    if active_flag and transform is not None:
        if len(transform) == 2:
            return bounds_transform(value, transform, False)
        if len(transform) == 1:
            return log_transform(value, transform, False)
        raise ValueError(f"Unexpected transform shape: {transform}")
    return value


def get_opt_bounds(transform):
    if transform is None or len(transform) == 1:
        return [None, None]
    return [-1.0, 1.0]


class Parameters:
    """Constitutive-model parameters as pytrees (mirror of the reference class)."""

    def __init__(self, values: dict, active_flags: dict | None = None,
                 transforms: dict | None = None) -> None:
        self.values = values
        self._active_flags = active_flags
        self._transforms = transforms
        self._flat_values = ravel_pytree(values)
        self.num_params = len(self._flat_values)
        leaves = tree_leaves_with_path(values)
        self._names = [str(path[-1]) for path, _ in leaves]
        self._paths = [path for path, _ in leaves]
        self.flat_param_sizes = [int(np.size(np.asarray(v))) for _, v in leaves]
        self.block_shapes = [(x, y) for x in self.flat_param_sizes for y in self.flat_param_sizes]
        if active_flags is not None:
            if transforms is None:
                raise AssertionError("transforms must be supplied when active_flags is set")
            flags, trs = [], []
            for size, a, t in zip(self.flat_param_sizes, tree_leaves(active_flags),
                                  tree_leaves(transforms)):
                flags += [bool(a)] * size
                trs += [t] * size
            self._flat_active_flags = np.array(flags, dtype=bool)
            self.num_active_params = int(self._flat_active_flags.sum())
            self.active_idx = np.arange(self.num_params)[self._flat_active_flags]
            self._flat_transforms = trs
            self._flat_active_transforms = [trs[i] for i in self.active_idx]
            self.opt_bounds = np.array([get_opt_bounds(t) for t in self._flat_active_transforms])
        else:
            assert transforms is None
            self.num_active_params = 0
            self.active_idx = np.zeros(0, dtype=np.intp)
            self._flat_active_flags = np.zeros(self.num_params, dtype=bool)
            self._flat_transforms = [None] * self.num_params
            self._flat_active_transforms = []

    # expanded leaf paths, one per flat entry (array leaves repeat their path)
    def flat_paths(self) -> list[tuple[tuple, int]]:
        out = []
        for path, size in zip(self._paths, self.flat_param_sizes):
            out += [(path, k) for k in range(size)]
        return out

    def set_rotation_matrix(self, rotation_matrix) -> None:
        self.values["rotation matrix"] = np.asarray(rotation_matrix, dtype=np.float64)
        self._flat_values = ravel_pytree(self.values)

    def set_active_values(self, values: dict, are_canonical: bool = True) -> None:
        if are_canonical:
            self.values = tree_map(
                lambda v, a, t: (transform_from_canonical(float(v), a, t)
                                 if np.ndim(v) == 0 else v),
                values, self._active_flags, self._transforms)
        else:
            self.values = values

    def set_active_values_from_flat(self, flat_active_values, are_canonical: bool = True) -> None:
        updated = np.array(self._flat_values, dtype=np.float64)
        updated[self.active_idx] = np.asarray(flat_active_values, dtype=np.float64)
        self.set_active_values(tree_unflatten_like(self.values, updated), are_canonical)

    def flat_active_values(self, return_canonical: bool = False) -> np.ndarray:
        flat = ravel_pytree(self.values)
        if return_canonical:
            return np.array([transform_to_canonical(float(v), a, t) for v, a, t in
                             zip(flat, self._flat_active_flags, self._flat_transforms)]
                            )[self.active_idx]
        return np.asarray(flat[self.active_idx])

    def get_active_from_flat(self, pytree: Any) -> np.ndarray:
        return ravel_pytree(pytree)[self.active_idx]

    def transform_grad(self, grad: np.ndarray) -> None:
        """Chain rule native -> canonical, in place (parameters.py:326-331)."""
        vals = self.get_active_from_flat(self.values)
        for ii in range(self.num_active_params):
            grad[ii] = first_deriv_transform(vals[ii], self._flat_active_transforms[ii]) * grad[ii]

    def transform_hessian(self, hessian: np.ndarray, grad: np.ndarray) -> None:
        """parameters.py:334-358."""
        vals = self.get_active_from_flat(self.values)
        n = self.num_active_params
        for ii in range(n):
            for jj in range(n):
                ti, tj = self._flat_active_transforms[ii], self._flat_active_transforms[jj]
                if ii == jj:
                    hessian[ii, ii] = hessian[ii, ii] * first_deriv_transform(vals[ii], ti) ** 2 \
                        + grad[ii] * second_deriv_transform(vals[ii], ti)
                elif ii < jj:
                    hessian[ii, jj] = hessian[ii, jj] * first_deriv_transform(vals[ii], ti) \
                        * first_deriv_transform(vals[jj], tj)
                else:
                    hessian[ii, jj] = hessian[jj, ii]

    def compute_mixed_block_shapes(self, num_eqs) -> None:
        self.mixed_block_shapes = [(x, y) for x in num_eqs for y in self.flat_param_sizes]

    def get_params_pytree_from_flat_canonical_active(self, flat_canonical_active) -> dict:
        """parameters.py:384-401: pure function flat canonical active -> native pytree."""
        flat = np.array(self._flat_values, dtype=np.float64)
        flat[self.active_idx] = np.asarray(flat_canonical_active, dtype=np.float64)
        tree = tree_unflatten_like(self.values, flat)
        return tree_map(lambda v, a, t: (transform_from_canonical(float(v), a, t)
                                         if np.ndim(v) == 0 else v),
                        tree, self._active_flags, self._transforms)
