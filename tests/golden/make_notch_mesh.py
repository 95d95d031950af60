#!/usr/bin/env python
"""Convert the reference's example mesh examples/meshes/notch.exo (eighth-symmetry notched
plate, 546 nodes, 1550 tet4, one block; CDF-2 classic netCDF, read with SciPy) into
tests/golden/notch_mesh.npz (nodes (546,3) float64, tets (1550,4) int64, 0-based) - the
input DATA of BASELINE.json configs[3] (examples/notch_hosford.yaml), so that the GPU box,
where /root/reference does not exist, can run that deck's problem at its native size.
Run in the build container:  python tests/golden/make_notch_mesh.py
"""
import os

import numpy as np
from scipy.io import netcdf_file

SRC = "/root/reference/examples/meshes/notch.exo"
HERE = os.path.dirname(os.path.abspath(__file__))

with netcdf_file(SRC, "r", mmap=False) as f:
    v = f.variables
    if "coord" in v:
        nodes = np.array(v["coord"][:], dtype=np.float64).T
    else:
        nodes = np.stack([np.array(v[k][:], dtype=np.float64) for k in ("coordx", "coordy", "coordz")], axis=1)
    tets = np.array(v["connect1"][:], dtype=np.int64) - 1
assert tets.shape[1] == 4
# orientation as the reference consumes it: iso_jac_det must be positive (cmad/fem/precompute.py:218)
X = nodes[tets]
det = np.linalg.det(np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=2))
print("nodes", nodes.shape, "tets", tets.shape, "min det", det.min(), "bbox", nodes.min(0), nodes.max(0))
np.savez_compressed(os.path.join(HERE, "notch_mesh.npz"), nodes=nodes, tets=tets)
