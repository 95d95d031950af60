"""Material-point calibration objectives on the B200 path.

Host-side mirror of ``cmad.objectives.mp_objective`` (MPAdjointObjective,
MPDirectObjective; mp_objective.py:22-215), ``cmad.qois.calibration.Calibration``
(calibration.py:21-66) and the model constructor surface they need
(``cmad.models.small_elastic_plastic.SmallElasticPlastic``), *batched*: every
objective evaluates many independent material points ("experiments") at once
and sums J and dJ/dp over them - the reference's per-point Python loop
(and its serial multi-experiment sum,
cmad/calibrations/al7079/multi_experiment_hill_calibration.py:20-32) becomes

    forward:  N launches of K1 (state history kept in HBM)
    gradient: one launch of K2 (adjoint or direct recurrence per point, fused
              with the QoI and a deterministic block reduction)
    multi-GPU: points are sharded by rank; the only collective is ONE allreduce
              (sum) of [J, grad] = 1 + n_active doubles over NCCL.

The single-point constructors keep the reference's signatures so its tests read
the same: ``MPAdjointObjective(Calibration(model, data, weight), F)`` with
``F`` of shape (3, 3, N+1).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, NamedTuple

import numpy as np
import torch

from . import _lib as L
from .comm import all_reduce_sum
from .material import NewtonSettings, active_param_ids, material_from_values
from .parameters import Parameters

# DefType of the reference (cmad/models/deformation_types.py:4-9)
FULL_3D, PLANE_STRAIN, PLANE_STRESS, UNIAXIAL_STRESS, PURE_SHEAR = range(5)
_N_XI = {FULL_3D: 7, PLANE_STRESS: 8, UNIAXIAL_STRESS: 9}
_NDIMS = {FULL_3D: 3, PLANE_STRESS: 2, UNIAXIAL_STRESS: 1}


class GradientResult(NamedTuple):
    J: float
    grad: np.ndarray


class HessianResult(NamedTuple):
    """cmad/objectives/objective.py HessianResult: canonical-coordinate gradient and Hessian."""
    J: float
    grad: np.ndarray
    hessian: np.ndarray


class SmallElasticPlastic:
    """Constructor-compatible stand-in for the reference model class: carries the
    parameters and the deformation type; all evaluation happens in the kernels."""
    model_name = "small_elastic_plastic"

    def __init__(self, parameters: Parameters, def_type: int = FULL_3D, yield_tol: float = 1e-14,
                 uniaxial_stress_idx: int = 0, **unsupported):
        if def_type not in _N_XI:
            raise NotImplementedError(f"DefType {def_type} is not on the B200 path (FULL_3D, PLANE_STRESS, "
                                      "UNIAXIAL_STRESS are)")
        if def_type == UNIAXIAL_STRESS and uniaxial_stress_idx != 0:
            raise NotImplementedError("UNIAXIAL_STRESS: only uniaxial_stress_idx = 0 is on the B200 path")
        if unsupported:
            raise NotImplementedError(f"unsupported model options: {sorted(unsupported)}")
        self.parameters = parameters
        self._def_type = def_type
        self.yield_tol = yield_tol
        self.num_dofs = _N_XI[def_type]                       # small_elastic_plastic.py:126-180
        self.num_residuals = 2 if def_type == FULL_3D else 3

    def material(self) -> L.Material:
        return material_from_values(self.parameters.values, self.model_name, self.yield_tol)


class SmallRateElasticPlastic(SmallElasticPlastic):
    """Stand-in for the rate form (cmad/models/small_rate_elastic_plastic.py): state =
    [cauchy(6), alpha] (+ the out-of-plane stretch under PLANE_STRESS, + two stretches and three
    off-axis delta strains under UNIAXIAL_STRESS: n_xi 8 / 12, :176-199); on the B200 path for the
    material-point update and the ``cmad primal`` loop (:mod:`cmad_b200.primal`) in the three
    def-types, and for the calibration objectives in FULL_3D."""
    model_name = "small_rate_elastic_plastic"

    def __init__(self, parameters: Parameters, def_type: int = FULL_3D, yield_tol: float = 1e-14, **unsupported):
        super().__init__(parameters, def_type, yield_tol, **unsupported)
        if def_type == UNIAXIAL_STRESS:
            self.num_dofs += 3
            self.num_residuals = 4


class Calibration:
    """``Calibration(model, data, weight)``: J = sum_t 1/2 ||weight o (cauchy - data[..., t])||^2.
    ``data`` is (3, 3, N+1) for one point or (B, 3, 3, N+1) for a batch."""

    def __init__(self, model: SmallElasticPlastic, data: np.ndarray, weight: np.ndarray) -> None:
        weight = np.asarray(weight, dtype=np.float64)
        assert weight.shape == (3, 3)
        self._model, self._data, self._weight = model, np.asarray(data, dtype=np.float64), weight

    def model(self):
        return self._model

    def data(self):
        return self._data


class UniaxialCalibration:
    """``UniaxialCalibration(model, data, weight, uniaxial_stress_idx, stretch_var_idx)``
    (cmad/qois/uniaxial_calibration.py:21-85), the QoI of uniaxial-stress tests that also measure
    the two lateral strains: ``J = sum_t 1/2 ||weight[:, t] o ([sigma_axial, lambda_2 - 1,
    lambda_3 - 1] - data[:, t])||^2`` with ``data`` and ``weight`` of shape ``(3, N+1)``
    (``(B, 3, N+1)`` data for a batch).  UNIAXIAL_STRESS models, ``uniaxial_stress_idx = 0`` and
    ``stretch_var_idx = 2`` (the block of off-axis stretches) - what the kernels carry."""
    qoi_kind = L.QOI_UNIAXIAL_CALIBRATION

    def __init__(self, model: SmallElasticPlastic, data: np.ndarray, weight: np.ndarray,
                 uniaxial_stress_idx: int = 0, stretch_var_idx: int = 2) -> None:
        data, weight = np.asarray(data, dtype=np.float64), np.asarray(weight, dtype=np.float64)
        if getattr(model, "_def_type", FULL_3D) != UNIAXIAL_STRESS:
            raise NotImplementedError("UniaxialCalibration needs a UNIAXIAL_STRESS model")
        if uniaxial_stress_idx != 0 or stretch_var_idx != 2:
            raise NotImplementedError("UniaxialCalibration: uniaxial_stress_idx = 0, stretch_var_idx = 2 on the B200 path")
        if weight.ndim != 2 or weight.shape[0] != 3 or data.shape[-2:] != weight.shape:
            raise ValueError("UniaxialCalibration: data (.., 3, N+1) and weight (3, N+1)")
        self._model, self._data, self._weight = model, data, weight

    def model(self):
        return self._model

    def data(self):
        return self._data


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced range of point indices owned by ``rank``."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def strain_history_from_F(F: np.ndarray) -> np.ndarray:
    """(B, nd, nd, N+1) deformation gradients -> (N+1, nd*nd, B) grad_u = F - I slabs
    (nd = 3 FULL_3D, 2 PLANE_STRESS, 1 UNIAXIAL_STRESS: cmad/models/kinematics.py:10-57)."""
    F = np.asarray(F, dtype=np.float64)
    nd = F.shape[1]
    gu = F - np.eye(nd)[None, :, :, None]
    return np.ascontiguousarray(gu.reshape(F.shape[0], nd * nd, F.shape[-1]).transpose(2, 1, 0))


def data_history(data: np.ndarray) -> np.ndarray:
    """(B, 3, 3, N+1) -> (N+1, 9, B)."""
    data = np.asarray(data, dtype=np.float64)
    return np.ascontiguousarray(data.reshape(data.shape[0], 9, data.shape[-1]).transpose(2, 1, 0))


def uniaxial_data_history(data: np.ndarray) -> np.ndarray:
    """UniaxialCalibration data (B, 3, N+1) = (sigma_axial, e_2, e_3) -> (N+1, 9, B) slabs whose
    rows 0..2 carry them (the layout CMADX_QOI_UNIAXIAL_CALIBRATION reads)."""
    data = np.asarray(data, dtype=np.float64)
    out = np.zeros((data.shape[-1], 9, data.shape[0]))
    out[:, :3, :] = data.transpose(2, 1, 0)
    return out


class _DeviceHistories:
    """Device-resident histories of this rank's shard."""

    def __init__(self, strain_hist: np.ndarray, data_hist: np.ndarray, device: torch.device, n_xi: int | None = None):
        self.N = strain_hist.shape[0] - 1
        self.n = strain_hist.shape[2]
        self.strain = torch.from_numpy(np.ascontiguousarray(strain_hist)).to(device)
        self.data = torch.from_numpy(np.ascontiguousarray(data_hist)).to(device)
        self.n_stretch = {9: 0, 6: 0, 4: 1, 3: 1, 1: 2}[strain_hist.shape[1]]
        # (the rate form under uniaxial stress: 12 = 9 + three off-axis delta strains)
        self.n_xi = n_xi if n_xi is not None else 7 + self.n_stretch
        self.xi = torch.zeros((self.N + 1, self.n_xi, self.n), dtype=torch.float64, device=device)
        self.iters = torch.zeros((self.N + 1, self.n), dtype=torch.int32, device=device)
        self.J_point = torch.zeros((self.n,), dtype=torch.float64, device=device)
        self.device = device


def gpu_local_evaluator(model: SmallElasticPlastic, strain_hist: np.ndarray, data_hist: np.ndarray,
                        weight: np.ndarray, strategy: str, device: torch.device,
                        newton: NewtonSettings | None = None,
                        reference_qoi_cross_terms: bool = False) -> Callable[[], torch.Tensor]:
    """Returns ``f() -> tensor[1 + n_active]`` (J, dJ/dp native) for this rank's
    points, evaluated with K1 + K2 on ``device`` at the model's current parameters."""
    lib = L.lib()
    nd = _NDIMS[getattr(model, "_def_type", FULL_3D)]
    if strain_hist.shape[1] not in ((6, 9) if nd == 3 else (nd * nd, 3) if nd == 2 else (1,)):
        raise ValueError(f"strain history with {strain_hist.shape[1]} rows does not fit the model's def_type")
    hist = _DeviceHistories(strain_hist, data_hist, device, getattr(model, "num_dofs", None))
    newton = newton or NewtonSettings(mode="imperative", max_iters=10, abs_tol=1e-14, rel_tol=1e-14)
    adjoint = {"adjoint": True, "direct": False, "direct_adjoint": True}[strategy]
    hessian = strategy == "direct_adjoint"
    weight = np.asarray(weight, dtype=np.float64)
    uniaxial_qoi = weight.shape != (3, 3)              # UniaxialCalibration: per-step weights (3, N+1)
    if uniaxial_qoi:
        if weight.shape != (3, hist.N + 1) or strain_hist.shape[1] != 1:
            raise ValueError("UniaxialCalibration: weight (3, N+1) on a UNIAXIAL_STRESS history")
        w_steps = torch.from_numpy(np.ascontiguousarray(weight.T)).to(device)        # (N+1, 3)
        w = np.zeros(9)
    else:
        w_steps = None
        w = weight.reshape(9)

    def evaluate() -> torch.Tensor:
        pid = active_param_ids(model.parameters)
        na = len(pid)
        mat = model.material()
        nw = newton.to_struct()
        result = torch.zeros((1 + na + (na * na if hessian else 0),), dtype=torch.float64, device=device)
        if hessian:
            ws_bytes = lib.cmadx_mp_hessian_workspace_bytes(C.c_int64(hist.n), C.c_int64(max(hist.n, 1)),
                                                            C.c_int32(hist.N), C.c_int32(na))
        else:
            ws_bytes = lib.cmadx_mp_objective_workspace_bytes(C.c_int64(hist.n), C.c_int32(na))
        ws = torch.empty((max(int(ws_bytes) // 8, 1),), dtype=torch.float64, device=device)
        h = L.MpHistory()
        h.n, h.ld, h.nsteps, h.strain_comps = hist.n, max(hist.n, 1), hist.N, hist.strain.shape[1]
        h.strain, h.data = hist.strain.data_ptr(), hist.data.data_ptr()
        for k in range(9):
            h.weight[k] = float(w[k])
        h.xi_hist, h.iters_hist = hist.xi.data_ptr(), hist.iters.data_ptr()
        h.result, h.workspace, h.J_point = result.data_ptr(), ws.data_ptr(), hist.J_point.data_ptr()
        if uniaxial_qoi:
            h.qoi_kind, h.weight_steps = L.QOI_UNIAXIAL_CALIBRATION, w_steps.data_ptr()
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        hist.xi[0].zero_()
        hist.xi[0, 7:7 + hist.n_stretch] = 1.0             # stretches of the def-type variants start at 1
        with torch.cuda.device(device):
            L.check(lib.cmadx_mp_forward_history(C.byref(mat), C.byref(nw), C.byref(h), stream),
                    "cmadx_mp_forward_history")
            pidp = pid.ctypes.data_as(C.POINTER(C.c_int32))
            if hessian:
                rc = lib.cmadx_mp_objective_hessian(C.byref(mat), pidp, na, C.byref(h),
                                                    C.c_int32(1 if reference_qoi_cross_terms else 0), stream)
            else:
                fn = lib.cmadx_mp_objective_adjoint if adjoint else lib.cmadx_mp_objective_direct
                rc = fn(C.byref(mat), pidp, na, C.byref(h), stream)
            L.check(rc, "cmadx_mp_objective")
        evaluate.histories = hist
        return result

    evaluate.histories = hist
    return evaluate


class BatchedMPObjective:
    """J(p), dJ/dp summed over a batch of material points, optionally sharded
    over the ranks of a ``torch.distributed`` process group.

    ``local_evaluator() -> tensor[1 + n_active]`` returns this rank's partial
    (J, grad) in native parameter coordinates; the class owns parameter
    injection (``set_active_values_from_flat``), the single allreduce, and the
    canonical-coordinate chain rule (``transform_grad``) - exactly the host
    scaffolding of ``MPObjective.evaluate`` (mp_objective.py:53-57, 143-147).

    The all-reduce is opt-in: pass ``group=cmad_b200.comm.WORLD`` (or a process group)
    when the evaluator's points are this rank's shard (:func:`shard_range`).  With
    ``group=None`` (the default, and what the reference-signature constructors use)
    nothing is summed across ranks, whether or not ``torch.distributed`` is initialised:
    every rank evaluating the same experiment gets that experiment's J, not world_size x J.
    """

    def __init__(self, parameters: Parameters, local_evaluator: Callable[[], torch.Tensor],
                 group=None) -> None:
        self._parameters = parameters
        self._local = local_evaluator
        self._group = group

    def evaluate(self, flat_active_values, are_canonical: bool = True) -> GradientResult:
        self._parameters.set_active_values_from_flat(np.asarray(flat_active_values, dtype=np.float64),
                                                     are_canonical)
        return self._evaluate()

    def _evaluate(self) -> GradientResult:
        partial = all_reduce_sum(self._local(), self._group)
        host = partial.detach().cpu().numpy()
        na = self._parameters.num_active_params
        grad = host[1:1 + na].copy()
        if host.size > 1 + na:                       # (J, grad, H): MPDirectAdjointObjective :269-340
            hess = host[1 + na:].reshape(na, na).copy()
            native_grad = grad.copy()
            self._parameters.transform_grad(grad)
            self._parameters.transform_hessian(hess, native_grad)
            return HessianResult(J=float(host[0]), grad=grad, hessian=hess)
        self._parameters.transform_grad(grad)
        return GradientResult(J=float(host[0]), grad=grad)


def _single_point_objective(qoi: Calibration, global_state: np.ndarray, strategy: str,
                            device=None, group=None, **kwargs) -> BatchedMPObjective:
    F = np.asarray(global_state, dtype=np.float64)
    data = qoi.data()
    if F.ndim == 3:
        F, data = F[None], data[None]
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    model = qoi.model()
    if getattr(qoi, "qoi_kind", L.QOI_CALIBRATION) == L.QOI_UNIAXIAL_CALIBRATION:
        dh = uniaxial_data_history(data)
    else:
        dh = data_history(data)
    ev = gpu_local_evaluator(model, strain_history_from_F(F), dh, qoi._weight, strategy, device, **kwargs)
    return BatchedMPObjective(model.parameters, ev, group)


def MPAdjointObjective(qoi: Calibration, global_state: np.ndarray, device=None, group=None):
    """Reference signature (mp_objective.py:92): gradient by the reverse-time adjoint."""
    return _single_point_objective(qoi, global_state, "adjoint", device, group)


def MPDirectAdjointObjective(qoi: Calibration, global_state: np.ndarray, device=None, group=None,
                             reference_qoi_cross_terms: bool = False):
    """Reference signature (mp_objective.py:218): J, gradient and Hessian (canonical
    coordinates) by the direct-adjoint method; ``evaluate`` returns a :class:`HessianResult`.
    Under ``torch.distributed`` the single all-reduce carries ``1 + P_a + P_a^2`` doubles.
    ``reference_qoi_cross_terms=True`` reproduces the reference entry for entry where its QoI
    drops the d2J/dxi dparams block (see CMADX_HESS_F_REFERENCE_QOI_CROSS in the header);
    the default is the complete Hessian."""
    return _single_point_objective(qoi, global_state, "direct_adjoint", device, group,
                                   reference_qoi_cross_terms=reference_qoi_cross_terms)


class MPJVPObjective:
    """The ``JVP`` strategy (cmad/objectives/mp_jvp_objective.py:14-80): J, its gradient and its
    Hessian as ``jax.value_and_grad`` / ``jax.hessian`` of the whole traced time loop give them -
    the TRACED local Newton with its line search (``update_fun = make_newton_solve(...)``,
    cmad/cli/sensitivity.py:98-107) and the exact second derivatives (``jax.hessian`` has no
    omitted block: this is the library's complete Hessian).  Same three entry points, canonical
    coordinates.  ``update_fun`` is accepted for signature compatibility and must be None or
    carry the traced solver's settings as attributes ``max_iters / abs_tol / rel_tol /
    ls_max_evals`` (the kernels implement make_newton_solve themselves)."""

    def __init__(self, qoi: Calibration, global_state: np.ndarray, update_fun=None, device=None, group=None):
        kw = {k: getattr(update_fun, k) for k in ("max_iters", "abs_tol", "rel_tol", "ls_max_evals")
              if update_fun is not None and hasattr(update_fun, k)}
        newton = NewtonSettings(mode="traced", **{"max_iters": 10, "abs_tol": 1e-14, "rel_tol": 1e-14, **kw})
        self._grad = _single_point_objective(qoi, global_state, "adjoint", device, group, newton=newton)
        self._hess = _single_point_objective(qoi, global_state, "direct_adjoint", device, group, newton=newton)

    def evaluate_objective(self, flat_active_values) -> float:
        return self._grad.evaluate(flat_active_values).J

    def evaluate_objective_and_grad(self, flat_active_values):
        r = self._grad.evaluate(flat_active_values)
        return r.J, r.grad

    def evaluate_hessian(self, flat_active_values) -> np.ndarray:
        return self._hess.evaluate(flat_active_values).hessian


def MPDirectObjective(qoi: Calibration, global_state: np.ndarray, device=None, group=None):
    """Reference signature (mp_objective.py:150): gradient by forward sensitivities."""
    return _single_point_objective(qoi, global_state, "direct", device, group)
