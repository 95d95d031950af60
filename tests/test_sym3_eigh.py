"""Symmetric 3x3 eigen-decomposition: the Jacobi routine of the kernels (`eig3_jacobi`, used by the
Yld2004-18p surface and exported as `cmadx_sym3_eigh`) against the reference's own closed-form
solver - `sorted_eigen_decomposition` of cmad/util/jax_eigen_decomposition.py:86-171, EXECUTED
unmodified by tests/golden/make_eig3_golden.py -> tests/golden/ref_eig3.npz.

Eigenvalues: 1e-13 of the tensor's norm (bit-exact is not defined between two algorithms; the
reference itself differs from LAPACK by 1e-15).  Eigenvectors: compared through what is defined -
orthonormality, A V = V diag(w), and |v_k . v_k_ref| = 1 where the eigenvalue is isolated."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.helpers import UP
from tests.test_barlat import _host_lib

D = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_eig3.npz"))


def _check(w, V):
    A, wr, Vr = D["A"], D["w"], D["V"]
    n = A.shape[0]
    nrm = np.abs(wr).max(axis=1) + 1e-300
    assert (np.abs(w - wr).max(axis=1) / nrm).max() < 1e-13
    assert np.all(np.diff(w, axis=1) >= 0)
    isolated = 0
    for i in range(n):
        assert np.abs(V[i].T @ V[i] - np.eye(3)).max() < 1e-14
        assert np.abs(A[i] @ V[i] - V[i] * w[i]).max() <= 1e-14 * nrm[i] + 1e-300
        gaps = np.abs(wr[i][:, None] - wr[i][None, :]) + np.eye(3) * 1e300
        for k in range(3):
            if gaps[k].min() > 1e-4 * nrm[i]:
                assert abs(abs(V[i][:, k] @ Vr[i][:, k]) - 1.0) < 1e-10, (i, k)
                isolated += 1
    assert isolated > 100


def test_host_build_of_the_kernel_routine_vs_reference():
    A = D["A"]
    n = A.shape[0]
    A6 = np.ascontiguousarray(np.array([[a[i, j] for i, j in UP] for a in A]))
    w = np.zeros((n, 3)); V = np.zeros((n, 3, 3))
    lib = _host_lib()
    lib.eig3_host.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.eig3_host.restype = None
    lib.eig3_host(n, A6.ctypes.data, w.ctypes.data, V.ctypes.data)
    _check(w, V)


@pytest.mark.gpu
def test_cuda_sym3_eigh_vs_reference(cuda_device):
    import torch
    from cmad_b200 import mp
    A = D["A"]
    A6 = torch.from_numpy(np.ascontiguousarray(np.array([[a[i, j] for i, j in UP] for a in A]).T)).to(cuda_device)
    w, V = mp.sym3_eigh(A6)
    torch.cuda.synchronize()
    _check(w.cpu().numpy().T.copy(), V.cpu().numpy().transpose(2, 0, 1).copy())
    w2, none = mp.sym3_eigh(A6, vectors=False)
    assert none is None and torch.equal(w2, w)


@pytest.mark.gpu
def test_cuda_sym3_eigh_large_batch_properties(cuda_device):
    """2^22 random tensors: trace, Frobenius norm and determinant are reproduced by the spectrum."""
    import torch
    from cmad_b200 import mp
    g = torch.Generator(device=cuda_device); g.manual_seed(3)
    A6 = torch.randn((6, 1 << 22), dtype=torch.float64, device=cuda_device, generator=g) * 100.0
    w, V = mp.sym3_eigh(A6)
    xx, xy, xz, yy, yz, zz = A6
    tr = xx + yy + zz
    fro = xx * xx + yy * yy + zz * zz + 2 * (xy * xy + xz * xz + yz * yz)
    det = xx * (yy * zz - yz * yz) - xy * (xy * zz - yz * xz) + xz * (xy * yz - yy * xz)
    s = fro.sqrt()
    assert float(((w.sum(0) - tr).abs() / s).max()) < 1e-14
    assert float((((w * w).sum(0) - fro).abs() / fro).max()) < 1e-14
    assert float(((w.prod(0) - det).abs() / s ** 3).max()) < 1e-13
    VtV = torch.einsum("mkn,mln->kln", V, V)
    assert float((VtV - torch.eye(3, dtype=torch.float64, device=cuda_device)[:, :, None]).abs().max()) < 1e-14
