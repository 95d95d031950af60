#!/usr/bin/env python
"""Generate tests/golden/ka1_analytic_paths.npz by EXECUTING THE REFERENCE'S OWN CODE.

The reference package cannot be imported here (``import cmad`` needs JAX), but its
known-answer generator is plain NumPy: ``cmad/verification/solutions.py``
(``compute_plastic_fields``) imports nothing else, and the yield functions /
normals it is driven with in the reference's tests (``J2_yield``,
``J2_yield_normal``, ``hill_yield``, ``hill_yield_normal`` in
``cmad/verification/functions.py:7-54``) are NumPy too - only that module's
top-level ``import jax.numpy`` / eigen-solver import (used by the Barlat
functions further down) stand in the way, so two empty stub modules are
registered for the duration of the import.  Nothing of the reference is copied:
the files are loaded from /root/reference and run unmodified.

Cases = tests/models/test_elastic_plastic_models.py:15-61 (``J2AnalyticalProblem``,
tests/support/test_problems.py:142-162): E=200e3, nu=0.3, Y=200, S=200, D=20;
stress masks uniaxial diag(1,0,0) and biaxial diag(1,-1,0); 100 steps,
max_alpha 0.5; J2 and Hill(F..N = 0.5, identical to J2).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/cmad/verification"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    solutions = _load("ref_solutions", os.path.join(REF, "solutions.py"))
    stubs = {"jax": types.ModuleType("jax"), "jax.numpy": types.ModuleType("jax.numpy"),
             "cmad": types.ModuleType("cmad"), "cmad.util": types.ModuleType("cmad.util"),
             "cmad.util.jax_eigen_decomposition": types.ModuleType("cmad.util.jax_eigen_decomposition")}
    stubs["cmad.util.jax_eigen_decomposition"].jax_compute_eigenvalues = None
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        functions = _load("ref_functions", os.path.join(REF, "functions.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    params = np.array([200e3, 0.3, 200., 200., 20.])
    hill = np.full(6, 0.5)
    masks = {"uniaxial": np.diag([1., 0., 0.]), "biaxial": np.diag([1., -1., 0.])}
    out = {"isotropic_params": params, "hill_params": hill, "max_alpha": np.array(0.5),
           "num_steps": np.array(100)}
    for mname, mask in masks.items():
        for yname, (yf, nf) in {
                "J2": (functions.J2_yield, functions.J2_yield_normal),
                "hill": (lambda c: functions.hill_yield(c, hill),
                         lambda c: functions.hill_yield_normal(c, hill))}.items():
            stress, strain, alpha = solutions.compute_plastic_fields(mask, yf, nf, params, 0.5, 100)
            out[f"{yname}_{mname}_mask"] = mask
            out[f"{yname}_{mname}_stress"] = stress
            out[f"{yname}_{mname}_strain"] = strain
            out[f"{yname}_{mname}_alpha"] = alpha
    path = os.path.join(HERE, "ka1_analytic_paths.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
