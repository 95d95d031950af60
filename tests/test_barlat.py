"""Yld2004-18p ("barlat") effective stress: parity with the reference's own run.

`tests/golden/ref_barlat.npz` was produced by EXECUTING THE REFERENCE'S UNMODIFIED SOURCE
(`make_reference_golden.py --only barlat`: `SmallElasticPlastic` with `effective_stress.py:81-84`
-> `verification/functions.py:71-154`, `jnp.linalg.eigh` with JAX's JVP rule on the NumPy stand-in):
traced Newton states / counts / flags / stress, `Model`'s AD products dC/dxi, dC/dxi_prev and dC/dp
over ALL leaves (18 tensor coefficients + exponent included), IFT tangents on a subset, and the
`MPAdjointObjective` / `MPDirectObjective` calibration objectives with tensor coefficients active.
Materials: the AL7079 fit of `cmad/calibrations/al7079/support.py:80-89` (a = 18.2), the same
tensors with a = 8, and rotated material axes with Voce + linear hardening.

CPU: the torch-AD oracle against the fixture; the CLOSED-FORM routine of the CUDA kernels
(`cmad_b200/csrc/barlat.cuh`, built for the host by `oracle/barlat_host.cpp`) against torch AD and
against the reference's Jacobian / dC/dp entries.  GPU: K1, the forward history and K2 through the
C-ABI against the fixture - counts and flags exact, values 1e-10, derivatives 1e-9.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cmad_b200 import Parameters
from tests.golden.materials import BARLAT_KEYS, NEWTON, const_like, material, objective_trees
from tests.helpers import UP, rel_err
from tests.test_reference_golden import _ROW9, _newton_kw, _sym_cols

G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BR = np.load(os.path.join(G, "ref_barlat.npz"))
CASES = ["barlat.mp", "barlat_rot.fe", "barlat_a8.mp"]
TAN = ["barlat.mp.tan", "barlat_rot.fe.tan"]
HEADER_ORDER = BARLAT_KEYS[:18]          # sp_12 .. sp_66, dp_12 .. dp_66 (cmadx_material_t::barlat)


def _active_sets(values):
    """The kernels carry at most CMADX_MAX_ACTIVE = 16 columns: the 23 differentiable scalar leaves
    of a Barlat material go through in two sets."""
    sets = []
    for keys in (("E", "nu", "Y", "S", "D", "K") + tuple(k for k in HEADER_ORDER if k.startswith("sp")),
                 tuple(k for k in HEADER_ORDER if k.startswith("dp")) + ("a",)):
        act = const_like(values, False)

        def mark(t, a):
            for k in t:
                if isinstance(t[k], dict):
                    mark(t[k], a[k])
                elif k in keys and np.ndim(t[k]) == 0:
                    a[k] = True
        mark(values, act)
        sets.append(Parameters(values, act, const_like(values, None)))
    return sets


# ------------------------------------------------------------------------------------------ #
#  CPU                                                                                        #
# ------------------------------------------------------------------------------------------ #
def _host_lib():
    so = os.path.join(ROOT, "oracle", "_build", "libbarlat_host.so")
    src = os.path.join(ROOT, "oracle", "barlat_host.cpp")
    deps = [src, os.path.join(ROOT, "oracle", "host_shims.h")] + \
        [os.path.join(ROOT, "cmad_b200", "csrc", f) for f in ("barlat.cuh", "point_solver.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-I/usr/local/cuda/include",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "cmad_b200", "csrc"),
                               "-I" + os.path.join(ROOT, "oracle"), "-shared", "-o", so, src])
    return C.CDLL(so)


def _closed_form(coeffs19, sig6):
    """phi, n (6), M (6, 6), dphi/dtheta (19), dn/dtheta (19, 6) from the kernels' routine."""
    out = np.zeros(43 + 7 * 19)
    c = np.ascontiguousarray(coeffs19, dtype=np.float64)
    s = np.ascontiguousarray(sig6, dtype=np.float64)
    rc = _host_lib().barlat_host_eval(c.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    d = out[43:].reshape(19, 7)
    return out[0], out[1:7], out[7:43].reshape(6, 6), d[:, 0], d[:, 1:]


def _coeffs(values):
    b = values["plastic"]["effective stress"]["barlat"]
    return np.array([b[k] for k in HEADER_ORDER] + [b["a"]])


def test_closed_form_vs_torch_ad():
    """Value, normal, Hessian and all 19 parameter derivatives of the kernels' closed form equal
    automatic differentiation of the oracle's restatement (eigh) at random and at special stresses."""
    import torch
    from torch.func import grad, jacfwd
    from oracle import cmad_oracle as co
    coeffs = _coeffs(material("barlat"))
    order = list(HEADER_ORDER) + ["a"]

    def phi(S, c):
        return co.barlat_effective_stress(S, {"effective stress": {"barlat": {k: c[i] for i, k in enumerate(order)}}})

    rng = np.random.default_rng(3)
    trials = [np.array([300., 0, 0, 0, 0, 0]), np.array([100., 50, 0, 100, 0, -30]), np.array([0., 120, 0, 0, 0, 0])]
    trials += [rng.normal(size=6) * 200 for _ in range(9)]
    for sig in trials:
        S = torch.zeros(3, 3, dtype=torch.float64)
        for k, (i, j) in enumerate(UP):
            S[i, j] = sig[k]; S[j, i] = sig[k]
        c = torch.tensor(coeffs)
        n = grad(phi)(S, c); H = jacfwd(grad(phi))(S, c)
        dphi = grad(phi, argnums=1)(S, c).numpy(); dn = jacfwd(grad(phi), argnums=1)(S, c)
        p, n6, M, dp, dn6 = _closed_form(coeffs, sig)
        assert abs(p - float(phi(S, c))) < 1e-13 * abs(p)
        assert rel_err(n6, np.array([float(n[i, j]) for i, j in UP])) < 1e-12
        Mref = np.array([[float(H[i, j, k, l]) + (float(H[i, j, l, k]) if k != l else 0.0) for k, l in UP] for i, j in UP])
        assert rel_err(M, Mref) < 1e-11
        assert rel_err(dp, dphi) < 1e-11
        assert rel_err(dn6, np.array([[float(dn[i, j, k]) for i, j in UP] for k in range(19)])) < 1e-11


def test_closed_form_is_finite_at_coincident_eigenvalues():
    """Isotropic tensors under uniaxial stress: both images have a double eigenvalue; JAX's eigh rule
    gives inf / NaN second derivatives there, the kernels use the limit - which is the Hessian of the
    Hosford / Hershey norm of the deviator (for a = 2: von Mises)."""
    c = np.array([1.0] * 18 + [2.0])
    p, n6, M, _, _ = _closed_form(c, np.array([250.0, 0, 0, 0, 0, 0]))
    assert np.isfinite(M).all() and np.isfinite(n6).all()
    # a = 2, L' = L'' = deviator projector: phi^2 = 1/4 sum_ij (s_i - s_j)^2 = 3/2 s:s
    assert abs(p - 250.0) < 1e-12 * 250
    s = np.array([2 / 3, 0, 0, -1 / 3, 0, -1 / 3]) * 250.0
    sn = np.sqrt(s[0] ** 2 + s[3] ** 2 + s[5] ** 2)
    sh = s / sn
    mult = np.array([1, 2, 2, 1, 2, 1.0])
    Mref = np.sqrt(1.5) / sn * (np.eye(6) - np.outer(sh, sh * mult) - np.array(
        [[(1 / 3 if (a in (0, 3, 5) and b in (0, 3, 5)) else 0.0) for b in range(6)] for a in range(6)]))
    assert rel_err(M, Mref) < 1e-9


@pytest.mark.parametrize("case", ["barlat.mp", "barlat_a8.mp"])
def test_closed_form_vs_reference_jacobian_entries(case):
    """dC/dxi and the Barlat columns of dC/dp the reference obtained by AD, rebuilt from the closed
    form at the reference's converged states: rows a < 6: delta_ab + dgamma 2 mu M(a, b) and
    -dgamma dn_a/dtheta, yield row: -mult(b) n_b and dphi/dtheta / 2 mu."""
    values = material(case.split(".")[0])
    coeffs = _coeffs(values)
    two_mu = values["elastic"]["E"] / (1.0 + values["elastic"]["nu"])
    names = [str(x) for x in BR[f"{case}.param_names"]]
    sizes = BR[f"{case}.param_sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    col = {}
    for nm, o in zip(names, offs):
        leaf = nm.split("'")[-2] if "'" in nm else nm
        col[leaf] = int(o)
    mult = np.array([1, 2, 2, 1, 2, 1.0])
    checked = 0
    for s in range(BR[f"{case}.xi"].shape[0]):
        for i in range(BR[f"{case}.xi"].shape[1]):
            if not (BR[f"{case}.flags"][s, i] & 2):
                continue
            xi, xp, sig = BR[f"{case}.xi"][s, i], BR[f"{case}.xi_prev"][s, i], BR[f"{case}.sigma"][s, i]
            dg = xi[6] - xp[6]
            _, n6, M, dphi, dn = _closed_form(coeffs, sig)
            J = BR[f"{case}.dC_dxi"][s, i]
            assert rel_err(np.eye(6) + dg * two_mu * M, J[:6, :6]) < 1e-9
            assert rel_err(-n6, J[:6, 6]) < 1e-9 and rel_err(-mult * n6, J[6, :6]) < 1e-9
            dCdp = BR[f"{case}.dC_dp"][s, i]
            for k, key in enumerate(list(HEADER_ORDER) + ["a"]):
                ref = dCdp[:, col[key]]
                got = np.concatenate([-dg * dn[k], [dphi[k] / two_mu]])
                assert np.abs(got - ref).max() < 1e-9 * max(np.abs(dCdp).max(), 1e-300), (case, s, i, key)
            checked += 1
    assert checked >= 8


@pytest.mark.parametrize("case", CASES)
def test_torch_oracle_vs_reference_traced_newton(case):
    import torch
    from oracle import cmad_oracle as co
    kind, key = case.split(".")
    tv = co.to_torch_tree(material(kind))
    spec = co.ModelSpec()
    n_plastic = 0
    for s in (0, BR[f"{case}.xi"].shape[0] - 1):
        for i in range(3):
            gu = torch.from_numpy(BR[f"{case}.grad_u"][s, i].reshape(3, 3).copy())
            x, info = co.newton_traced(BR[f"{case}.xi_prev"][s, i].copy(), tv, gu, gu, spec, **NEWTON[key])
            assert info.iters == BR[f"{case}.iters"][s, i]
            assert (info.flag_entry | 2 * info.flag_exit) == BR[f"{case}.flags"][s, i]
            assert rel_err(x.numpy(), BR[f"{case}.xi"][s, i]) < 1e-10
            n_plastic += info.flag_exit
    assert n_plastic > 0


def _objective_inputs(case):
    mode = case.split(".")[-1]
    values, act, tr = objective_trees("barlat", mode == "scaled")
    P = Parameters(values, act, tr)
    assert np.array_equal(P.active_idx, BR[f"{case}.active_idx"])
    P.set_active_values_from_flat(BR[f"{case}.x_canonical"], are_canonical=True)
    return P, BR[f"{case}.F"], BR[f"{case}.data"], BR[f"{case}.weight"]


def test_torch_oracle_objective_vs_reference():
    """Adjoint objective with tensor coefficients and the exponent active (native parameters)."""
    from oracle import cmad_oracle as co
    case = "objective.barlat.native"
    values, act, tr = objective_trees("barlat", False)
    P = co.OracleParameters(values, act, tr)
    P.set_active_values_from_flat(BR[f"{case}.x_canonical"], True)
    r = co.mp_objective_adjoint(P, BR[f"{case}.F"], BR[f"{case}.data"], BR[f"{case}.weight"], co.ModelSpec())
    assert abs(r[0] - BR[f"{case}.J_adjoint"]) < 1e-11 * abs(r[0])
    assert rel_err(np.asarray(r[1]), BR[f"{case}.grad_adjoint"]) < 1e-9
    assert rel_err(BR[f"{case}.grad_direct"], BR[f"{case}.grad_adjoint"]) < 1e-9      # the reference's two strategies


def test_material_and_parameter_ids():
    from cmad_b200 import _lib as L, active_param_ids, material_from_values
    values = material("barlat")
    m = material_from_values(values)
    assert m.yield_ == L.YIELD_BARLAT and m.barlat_a == 18.2
    assert np.allclose(np.array(m.barlat[:]), _coeffs(values)[:18])
    P = _active_sets(values)[1]                               # dp_* and a, in the reference's sorted-key order
    ids = active_param_ids(P)
    assert ids[0] == L.P_BARLAT_A and list(ids[1:]) == [L.P_BARLAT_C0 + 9 + k for k in range(9)]


# ------------------------------------------------------------------------------------------ #
#  GPU: K1 / forward history / K2 through the C-ABI against the reference's own output        #
# ------------------------------------------------------------------------------------------ #
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + TAN)
def test_cuda_vs_reference_traced_newton(cuda_device, case):
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp
    kind, key = case.split(".")[:2]
    values = material(kind)
    tan = case.endswith(".tan")
    want = ("xi", "sigma", "dC_dp", "dC_dxi", "dC_dxi_prev", "iters", "flags") + (("dsig_deps", "dxi_deps") if tan else ())
    g = {k: BR[f"{case}.{k}"] for k in ("grad_u", "xi_prev", "xi", "iters", "flags", "sigma", "dC_dp", "dC_dxi", "dC_dxi_prev")}
    n_plastic = 0
    for P in _active_sets(values):
        aidx = np.asarray(P.active_idx)
        for s in range(g["xi"].shape[0]):
            out = mp.mp_update(material_from_values(values), NewtonSettings(mode="traced", **_newton_kw(key)),
                               active_param_ids(P), torch.from_numpy(g["xi_prev"][s].T.copy()).to(cuda_device),
                               torch.from_numpy(g["grad_u"][s].T.copy()).to(cuda_device), outputs=want)
            torch.cuda.synchronize()
            out = {k: v.cpu().numpy() for k, v in out.items()}
            assert np.array_equal(out["iters"], g["iters"][s]) and np.array_equal(out["flags"], g["flags"][s]), \
                (case, s, out["iters"], g["iters"][s])
            n = out["iters"].size
            n_plastic += int((g["flags"][s] & 2).astype(bool).sum())
            assert rel_err(out["xi"], g["xi"][s].T) < 1e-10 and rel_err(out["sigma"], g["sigma"][s].T) < 1e-10
            assert rel_err(out["dC_dxi"], g["dC_dxi"][s].reshape(n, 49).T) < 1e-9
            assert rel_err(out["dC_dxi_prev"], g["dC_dxi_prev"][s].reshape(n, 49).T) < 1e-9
            assert rel_err(out["dC_dp"], g["dC_dp"][s][:, :, aidx].reshape(n, -1).T) < 1e-9, (case, s)
            if tan:
                assert rel_err(out["dsig_deps"], _sym_cols(BR[f"{case}.dsig_dgradu"][s])[:, _ROW9, :].reshape(n, 36).T) < 1e-9
                assert rel_err(out["dxi_deps"], _sym_cols(BR[f"{case}.dxi_dgradu"][s]).reshape(n, 42).T) < 1e-9
                A = out["dC_dxi"].T.reshape(n, 7, 7)
                dxdp = -np.linalg.solve(A, out["dC_dp"].T.reshape(n, 7, len(aidx)))
                ok = (g["flags"][s] & 2) > 0
                if ok.any():
                    assert rel_err(dxdp[ok], BR[f"{case}.dxi_dp"][s][:, :, aidx][ok]) < 1e-8
    assert n_plastic > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["scaled", "native"])
def test_cuda_objectives_vs_reference(cuda_device, mode):
    """K1 history + K2 adjoint / direct with the reference's constructor signatures; `native` has
    four tensor coefficients and the exponent among its ten active parameters."""
    from cmad_b200.objectives import Calibration, MPAdjointObjective, MPDirectObjective, SmallElasticPlastic
    case = f"objective.barlat.{mode}"
    for strategy, ctor in (("adjoint", MPAdjointObjective), ("direct", MPDirectObjective)):
        P, F, data, w = _objective_inputs(case)
        obj = ctor(Calibration(SmallElasticPlastic(P), data, w), F, device=cuda_device)
        r = obj.evaluate(BR[f"{case}.x_canonical"])
        assert abs(r.J - BR[f"{case}.J_{strategy}"]) < 1e-11 * abs(r.J)
        assert rel_err(r.grad, BR[f"{case}.grad_{strategy}"]) < 1e-9


BFE_PATH = os.path.join(G, "ref_barlat_fe.npz")
BFE = np.load(BFE_PATH) if os.path.exists(BFE_PATH) else None
FE_CASES = sorted({k.rsplit(".", 1)[0] for k in BFE.files}) if BFE is not None else []


def test_fe_fixture_is_plastic_and_complete():
    """`per_element_R_and_K_coupled` of the reference on distorted tet4 / hex8 elements with the
    Yld2004-18p surface, displacement and mixed u-p, rotated axes included."""
    assert FE_CASES == ["hex8.barlat.disp", "hex8.barlat_rot.mixed", "tet4.barlat.mixed", "tet4.barlat_rot.disp"]
    for case in FE_CASES:
        assert BFE[f"{case}.xi"][..., 6].max() > 0
        assert rel_err(BFE[f"{case}.R_only_u"], BFE[f"{case}.R_u"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("case", FE_CASES)
@pytest.mark.parametrize("deterministic", [True, False])
def test_cuda_fe_vs_reference_elements(cuda_device, case, deterministic):
    """K3 / K4 (the any-rule element kernel with the Barlat point routine) against the reference's
    own element residuals, tangent blocks and states."""
    from tests.test_fe_reference_golden import _cuda_fe_vs_fixture
    _cuda_fe_vs_fixture(cuda_device, case, deterministic, BFE)


@pytest.mark.gpu
def test_cuda_barlat_unsupported_entries_say_so(cuda_device):
    """Entry points that do not carry the surface return CMADX_EUNSUPPORTED, never a wrong answer:
    the Hessian pass of the calibration objective."""
    from cmad_b200.objectives import Calibration, MPDirectAdjointObjective, SmallElasticPlastic
    P, F, data, w = _objective_inputs("objective.barlat.scaled")
    obj = MPDirectAdjointObjective(Calibration(SmallElasticPlastic(P), data, w), F, device=cuda_device)
    with pytest.raises(NotImplementedError):
        obj.evaluate(BR["objective.barlat.scaled.x_canonical"])
