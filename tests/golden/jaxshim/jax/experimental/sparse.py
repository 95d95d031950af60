"""`jax.experimental.sparse` stand-in: the BCOO / BCSR containers the reference's global
assembly and embedded-BC code builds (cmad/fem/assembly.py:906-916, cmad/fem/sparse_solve.py:42-66,
1058-1176).  Forward evaluation only (NumPy fp64); TEST INFRASTRUCTURE, not JAX."""
import numpy as np

from .._core import Array, _raw, asarray


def _np(x):
    return np.asarray(_raw(x) if isinstance(x, Array) else x)


class BCOO:
    def __init__(self, args, shape, indices_sorted=False, unique_indices=False):
        data, indices = args
        self.data, self.indices = asarray(data), asarray(indices)
        self.shape = tuple(int(s) for s in shape)
        self.indices_sorted, self.unique_indices = indices_sorted, unique_indices

    @property
    def nse(self):
        return int(_np(self.data).shape[0])

    @property
    def dtype(self):
        return _np(self.data).dtype

    def todense(self):
        out = np.zeros(self.shape)
        idx = _np(self.indices)
        np.add.at(out, (idx[:, 0], idx[:, 1]), _np(self.data))
        return asarray(out)

    def __matmul__(self, x):
        xv = _np(x)
        idx = _np(self.indices)
        out = np.zeros((self.shape[0],) + xv.shape[1:])
        np.add.at(out, idx[:, 0], (_np(self.data).reshape((-1,) + (1,) * (xv.ndim - 1))) * xv[idx[:, 1]])
        return asarray(out)

    @property
    def T(self):
        idx = _np(self.indices)
        return BCOO((self.data, idx[:, ::-1].copy()), shape=self.shape[::-1])

    def sum_duplicates(self, nse=None):
        idx = _np(self.indices)
        key = idx[:, 0].astype(np.int64) * self.shape[1] + idx[:, 1]
        uk, inv = np.unique(key, return_inverse=True)
        d = np.zeros(uk.shape[0])
        np.add.at(d, inv, _np(self.data))
        return BCOO((d, np.stack([uk // self.shape[1], uk % self.shape[1]], axis=-1)), shape=self.shape,
                    indices_sorted=True, unique_indices=True)


class BCSR:
    def __init__(self, args, shape):
        data, indices, indptr = args
        self.data, self.indices, self.indptr = asarray(data), asarray(indices), asarray(indptr)
        self.shape = tuple(int(s) for s in shape)

    def _csr(self):
        import scipy.sparse
        return scipy.sparse.csr_matrix((_np(self.data), _np(self.indices), _np(self.indptr)), shape=self.shape)

    def __matmul__(self, x):
        return asarray(self._csr() @ _np(x))

    def todense(self):
        return asarray(self._csr().toarray())
