"""``cmad primal`` on the B200 path: the forward time-step loop with stress / state recording
and the reference's on-disk output layouts.

``run_primal_pass`` mirrors ``cmad/cli/primal.py:129-176`` (same arguments, same return
tuple) for one material point - or a batch of them, which is what the GPU is for - and the
``write_*`` functions produce the files of ``cmad/io/writers.py:63-172`` byte-compatible with
the reference's (``np.save`` / ``np.savetxt`` / ``json.dump`` of arrays in its layouts):

    cauchy.npy          (3, 3, N+1)                 per point
    xi_block_<k>.npy    (N+1, num_eqs_in_block)     one file per residual block
    solver.json         [{"iters", "final_residual"}, ...] per step
    J.json, grad.npy, hess.npy

The loop runs K1 once per load step (imperative ``newton_solve`` flavour, as the reference's
primal pass does); there is no CPU fallback.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Sequence

import numpy as np
import torch

from . import _lib as L
from . import mp
from .material import NewtonSettings
from .objectives import (FULL_3D, PLANE_STRESS, UNIAXIAL_STRESS, Calibration, SmallElasticPlastic,
                         strain_history_from_F)

# the reference's DefType enum (cmad/models/deformation_types.py) -> the library's CMADX_DEF_*
_LIB_DEF_TYPE = {FULL_3D: L.DEF_FULL_3D, PLANE_STRESS: L.DEF_PLANE_STRESS, UNIAXIAL_STRESS: L.DEF_UNIAXIAL_STRESS}

_COMP = np.array([0, 1, 2, 1, 3, 4, 2, 4, 5])       # row-major 3x3 entry -> packed component
_CAUCHY_HEADER = "S11 S12 S13 S21 S22 S23 S31 S32 S33"


def block_sizes(model: SmallElasticPlastic) -> list[int]:
    """Residual blocks of the state vector: plastic strain (6), alpha (1), then the stretch
    block of the PLANE_STRESS (1) / UNIAXIAL_STRESS (2) def-types
    (cmad/models/small_elastic_plastic.py:126-180)."""
    extra = model.num_dofs - 7
    if extra == 5:          # the rate form under UNIAXIAL_STRESS: stretches (2) + off-axis delta strains (3)
        return [6, 1, 2, 3]
    return [6, 1] + ([extra] if extra else [])


def run_primal_pass(model: SmallElasticPlastic, F: np.ndarray, num_steps: int,
                    newton_kwargs: dict[str, Any] | None = None, qoi: Calibration | None = None,
                    device=None):
    """``(cauchy, xi_trajectory, solver_log, J)`` as ``run_primal_pass`` of the reference returns
    them for ``F (nd, nd, N+1)``; for a batch ``F (B, nd, nd, N+1)`` every entry gains a leading
    point axis (``cauchy (B, 3, 3, N+1)``, ``xi_trajectory[step][block] (B, n_eqs)``,
    ``solver_log[step]["iters"] (B,)``, ``J`` summed over the batch)."""
    F = np.asarray(F, dtype=np.float64)
    single = F.ndim == 3
    Fb = F[None] if single else F
    if Fb.shape[-1] != num_steps + 1:
        raise ValueError(f"F holds {Fb.shape[-1] - 1} steps, expected {num_steps}")
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    kw = dict(newton_kwargs or {})
    newton = NewtonSettings(mode="imperative", max_iters=kw.pop("max_iters", 10),
                            abs_tol=kw.pop("abs_tol", 1e-14), rel_tol=kw.pop("rel_tol", 1e-14),
                            max_ls_evals=int(kw.pop("max_ls_evals", 0)))     # legacy line search of newton_solve
    if kw:
        raise ValueError(f"unknown newton_kwargs {sorted(kw)}")
    B, n_xi, dt = Fb.shape[0], model.num_dofs, _LIB_DEF_TYPE[model._def_type]
    strain = torch.from_numpy(strain_history_from_F(Fb)).to(device)           # (N+1, comps, B)
    mat = model.material()
    xi = mp.init_xi(mat, B, device, def_type=dt)                               # stretches start at 1
    assert xi.shape[0] == n_xi
    sizes = block_sizes(model)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    split = lambda x: [x[:, offs[k]:offs[k + 1]].copy() for k in range(len(sizes))]
    cauchy = np.zeros((B, 3, 3, num_steps + 1))
    xi_traj = [split(xi.T.cpu().numpy())]
    solver_log = []
    J = 0.0
    if qoi is not None:
        data = qoi.data() if not single else qoi.data()[None]
        w = qoi._weight
    rate = model.model_name == "small_rate_elastic_plastic"     # its kernel takes the strain INCREMENT
    for step in range(1, num_steps + 1):
        e = (strain[step] - strain[step - 1]) if rate else strain[step]
        out = mp.mp_update(mat, newton, [], xi, e.contiguous(),
                           outputs=("xi", "sigma", "iters", "cnorm"), def_type=dt)
        xi = out["xi"]
        sig = out["sigma"].T.cpu().numpy()[:, _COMP].reshape(B, 3, 3)
        cauchy[..., step] = sig
        xi_traj.append(split(xi.T.cpu().numpy()))
        solver_log.append({"iters": out["iters"].cpu().numpy().copy(),
                           "final_residual": out["cnorm"].cpu().numpy().copy()})
        if qoi is not None:
            mis = w[None] * (sig - data[..., step])
            J += 0.5 * float((mis * mis).sum())
    if single:
        cauchy = cauchy[0]
        xi_traj = [[b[0] for b in blocks] for blocks in xi_traj]
        solver_log = [{"iters": int(s["iters"][0]), "final_residual": float(s["final_residual"][0])}
                      for s in solver_log]
    return cauchy, xi_traj, solver_log, J


# ---- writers: cmad/io/writers.py:63-172 --------------------------------------------------
def _check_fmt(fmt: str) -> None:
    if fmt not in {"npy", "text"}:
        raise ValueError(f"output.format: expected 'npy' or 'text', got {fmt!r}")


def write_cauchy(out_dir, prefix: str, cauchy: np.ndarray, fmt: str) -> None:
    """The ``(3, 3, N+1)`` Cauchy trajectory (writers.py:63-83)."""
    _check_fmt(fmt)
    out_dir = Path(out_dir)
    if fmt == "npy":
        np.save(out_dir / f"{prefix}cauchy.npy", cauchy)
    else:
        np.savetxt(out_dir / f"{prefix}cauchy.csv", cauchy.transpose(2, 0, 1).reshape(-1, 9),
                   header=_CAUCHY_HEADER)


def write_xi(out_dir, prefix: str, xi_trajectory: Sequence[Sequence[np.ndarray]], fmt: str) -> None:
    """One file per residual block, shape ``(N+1, num_eqs_in_block)`` (writers.py:86-112)."""
    _check_fmt(fmt)
    out_dir = Path(out_dir)
    if not xi_trajectory:
        return
    for k in range(len(xi_trajectory[0])):
        per_step = np.stack([xi_trajectory[t][k] for t in range(len(xi_trajectory))])
        if fmt == "npy":
            np.save(out_dir / f"{prefix}xi_block_{k:02d}.npy", per_step)
        else:
            np.savetxt(out_dir / f"{prefix}xi_block_{k:02d}.csv", per_step)


def write_solver_log(out_dir, prefix: str, solver_log) -> None:
    # batched runs carry arrays per step (one entry per experiment): lists in the JSON
    plain = [{k: (np.asarray(v).tolist() if isinstance(v, (np.ndarray, np.generic)) else v) for k, v in s.items()}
             for s in solver_log]
    with (Path(out_dir) / f"{prefix}solver.json").open("w") as f:
        json.dump(plain, f, indent=2)


def write_J(out_dir, prefix: str, J: float) -> None:
    with (Path(out_dir) / f"{prefix}J.json").open("w") as f:
        json.dump({"J": J}, f, indent=2)


def write_grad(out_dir, prefix: str, grad: np.ndarray, fmt: str) -> None:
    _check_fmt(fmt)
    if fmt == "npy":
        np.save(Path(out_dir) / f"{prefix}grad.npy", grad)
    else:
        np.savetxt(Path(out_dir) / f"{prefix}grad.csv", grad)


def write_hessian(out_dir, prefix: str, hessian: np.ndarray, fmt: str) -> None:
    _check_fmt(fmt)
    if fmt == "npy":
        np.save(Path(out_dir) / f"{prefix}hess.npy", hessian)
    else:
        np.savetxt(Path(out_dir) / f"{prefix}hess.csv", hessian)
