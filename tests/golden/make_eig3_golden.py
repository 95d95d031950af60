"""Golden vectors of the reference's closed-form symmetric 3x3 eigen-decomposition:
EXECUTES /root/reference/cmad/util/jax_eigen_decomposition.py (`sorted_eigen_decomposition`,
unmodified) on the NumPy `jax` stand-in -> tests/golden/ref_eig3.npz.  Inputs: random symmetric
tensors at three magnitudes, stress-like tensors with one dominant direction, nearly coincident
eigenvalue pairs, and diagonal tensors (the reference's `cond` takes its diagonal branch there).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "jaxshim"), "/root/reference"]

import numpy as np  # noqa: E402

import jax  # noqa: E402  (the shim)
from cmad.util.jax_eigen_decomposition import sorted_eigen_decomposition  # noqa: E402


def main():
    rng = np.random.default_rng(17)
    mats = []
    for scale in (1.0, 1e3, 1e-4):
        for _ in range(12):
            A = rng.normal(size=(3, 3)) * scale
            mats.append(A + A.T)
    for _ in range(8):                                   # stress-like: a large principal value
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        A = rng.normal(size=(3, 3)) * 5.0
        mats.append(400.0 * np.outer(v, v) + A + A.T)
    for gap in (1e-3, 1e-6):                             # close pairs (well-defined eigenvalues)
        for _ in range(4):
            Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            mats.append(Q @ np.diag([1.0, 1.0 + gap, 3.0]) @ Q.T)
    mats += [np.diag([3.0, -1.0, 2.0]), np.diag([1.0, 1.0, 5.0]), np.zeros((3, 3))]
    A = np.array([0.5 * (m + m.T) for m in mats])
    w, V = [], []
    for m in A:
        wi, Vi = sorted_eigen_decomposition(jax.numpy.asarray(m))
        w.append(np.asarray(wi)); V.append(np.asarray(Vi))
    np.savez_compressed(os.path.join(HERE, "ref_eig3.npz"), A=A, w=np.array(w), V=np.array(V))
    print("ref_eig3.npz:", A.shape[0], "tensors")


if __name__ == "__main__":
    main()
