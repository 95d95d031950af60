"""CPU-side checks: parameter pytrees (flatten order, active columns, transforms),
material mapping, and that the C-ABI library loads and exports every symbol
include/cmad_b200.h declares (no compute calls without a GPU)."""
import ctypes as C

import numpy as np
import pytest

from cmad_b200 import NewtonSettings, Parameters, _lib, active_param_ids, material_from_values
from cmad_b200 import material as matmod
from oracle import analytic
from tests.helpers import param_tree


def test_flatten_order_matches_reference_sorted_keys():
    # [elastic.E, elastic.nu, effective stress.J2, voce.D, voce.S, Y, Q(9)]
    values, act, tr = analytic.j2_voce_param_tree("J2")
    p = Parameters(values, act, tr)
    assert p.num_params == 15
    assert p._names == ["E", "nu", "J2", "D", "S", "Y", "rotation matrix"]
    assert list(p.active_idx) == [3, 4, 5]                 # D, S, Y
    assert list(active_param_ids(p)) == [_lib.P_VOCE_D, _lib.P_VOCE_S, _lib.P_Y]
    assert p.opt_bounds.tolist() == [[-1.0, 1.0], [-1.0, 1.0], [None, None]]


def test_uniaxial_deck_active_order():
    # examples/elastic_plastic_uniaxial.yaml: active = [E, nu, D, S, Y]
    values, act, tr = param_tree("J2")
    p = Parameters(values, act, tr)
    assert list(active_param_ids(p)) == [_lib.P_EL0, _lib.P_EL1, _lib.P_VOCE_D, _lib.P_VOCE_S, _lib.P_Y]


def test_transforms_round_trip_and_grad_chain_rule():
    values, act, tr = analytic.j2_voce_param_tree("J2")       # log Y, bounds S, D
    p = Parameters(values, act, tr)
    canon = p.flat_active_values(return_canonical=True)
    assert np.allclose(canon, [0.0, 0.0, 0.0])                # D, S mid-bounds; Y = ref
    p.set_active_values_from_flat(np.array([0.5, -0.25, 0.1]))
    nat = p.flat_active_values()
    assert np.allclose(nat, [25.0, 175.0, 200.0 * np.exp(0.1)])
    g = np.array([1.0, 1.0, 1.0]); p.transform_grad(g)
    assert np.allclose(g, [10.0, 100.0, 200.0 * np.exp(0.1)])  # span, span, value
    H = np.eye(3); p.transform_hessian(H, np.array([1.0, 1.0, 1.0]))
    assert np.isclose(H[2, 2], nat[2] ** 2 + nat[2]) and np.isclose(H[0, 0], 100.0)
    tree = p.get_params_pytree_from_flat_canonical_active(np.zeros(3))
    assert tree["plastic"]["flow stress"]["initial yield"]["Y"] == pytest.approx(200.0)


def test_material_mapping_and_errors():
    values, act, tr = param_tree("hill", ("voce", "linear"), hill=(.4, .5, .6, 1.1, 1.2, 1.3))
    m = material_from_values(values)
    assert m.yield_ == _lib.YIELD_HILL and m.hardening_mask == 3
    assert list(m.hill) == [.4, .5, .6, 1.1, 1.2, 1.3] and m.linear_K == 1500.0
    assert _lib.ELASTIC_PAIRS[m.elastic_pair] == ("E", "nu")
    with pytest.raises(NotImplementedError):
        bad = {**values, "plastic": {**values["plastic"], "effective stress": {"hybrid_hill": {}}}}
        material_from_values(bad)
    with pytest.raises(ValueError):
        material_from_values({"elastic": {"E": 1.0}}, model="elastic")
    # an active leaf that does not enter the model is an error, not a silent zero column
    values, act, tr = analytic.j2_voce_param_tree("J2")
    act["plastic"]["effective stress"]["J2"] = True
    with pytest.raises(ValueError):
        active_param_ids(Parameters(values, act, tr))


def test_library_loads_and_exports_declared_symbols(built_lib):
    names = _lib.exported_symbols()
    assert "cmadx_mp_update" in names and "cmadx_mp_update_host" in names
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/cmad_b200.h but not exported"
    assert built_lib.cmadx_version() == 100
    sizes = (C.c_int64 * 5)()
    assert built_lib.cmadx_struct_sizes(sizes) == 0
    assert list(sizes) == [C.sizeof(_lib.Material), C.sizeof(_lib.Newton), C.sizeof(_lib.MpBuffers),
                           C.sizeof(_lib.MpHistory), C.sizeof(_lib.FeBlock)]
    assert built_lib.cmadx_error_string(1) == b"invalid argument"


def test_lame_through_library_all_pairs(built_lib):
    E, nu = 200e3, 0.3
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    vals = {"E": E, "nu": nu, "mu": mu, "kappa": lam + 2 * mu / 3, "lambda": lam}
    for a, b in _lib.ELASTIC_PAIRS:
        m = material_from_values({"elastic": {a: vals[a], b: vals[b]}}, model="elastic")
        out = matmod.lame(m)
        assert out[0] == pytest.approx(lam, rel=1e-10) and out[1] == pytest.approx(mu, rel=1e-10)
        # derivative check by central differences on the library itself
        for j, key in enumerate((a, b)):
            h = 1e-6 * vals[key]
            up = material_from_values({"elastic": {a: vals[a] + (h if j == 0 else 0), b: vals[b] + (h if j == 1 else 0)}}, model="elastic")
            dn = material_from_values({"elastic": {a: vals[a] - (h if j == 0 else 0), b: vals[b] - (h if j == 1 else 0)}}, model="elastic")
            fd = (matmod.lame(up)[:2] - matmod.lame(dn)[:2]) / (2 * h)
            assert out[2 + j] == pytest.approx(fd[0], rel=1e-5, abs=1e-6)
            assert out[4 + j] == pytest.approx(fd[1], rel=1e-5, abs=1e-6)


def test_argument_validation_without_gpu(built_lib):
    values, act, tr = param_tree("J2")
    m = material_from_values(values)
    nw = NewtonSettings().to_struct()
    b = _lib.MpBuffers(); b.n = 4; b.ld = 2; b.strain_comps = 6          # ld < n
    assert built_lib.cmadx_mp_update(C.byref(m), C.byref(nw), None, 0, C.byref(b), None) == _lib.EINVAL
    b.ld = 4; b.strain_comps = 7
    assert built_lib.cmadx_mp_update(C.byref(m), C.byref(nw), None, 0, C.byref(b), None) == _lib.EINVAL
    # rotation-matrix entries are not differentiated in the def-type kernels (the Hosford exponent is,
    # since the def-type kernels carry its dC/dp column)
    b.strain_comps = 3; b.def_type = _lib.DEF_PLANE_STRESS
    pid = (C.c_int32 * 1)(_lib.P_Q00 + 4)
    assert built_lib.cmadx_mp_update(C.byref(m), C.byref(nw), pid, 1, C.byref(b), None) == _lib.EUNSUPPORTED
    b.strain_comps = 6; b.def_type = _lib.DEF_FULL_3D
    assert built_lib.cmadx_mp_update(C.byref(m), C.byref(nw), pid, 1, C.byref(b), None) == _lib.EINVAL   # no buffers
    nw.ls_max_evals = 0
    assert built_lib.cmadx_mp_update(C.byref(m), C.byref(nw), None, 0, C.byref(b), None) == _lib.EINVAL
    b.n = 0; nw.ls_max_evals = 4                                         # empty batch is a no-op
    assert built_lib.cmadx_mp_update(C.byref(m), C.byref(nw), None, 0, C.byref(b), None) == _lib.OK


def test_newton_settings_from_reference_kwargs_maps_every_line_search_field():
    """Regression: the fields are passed by keyword (a field inserted into the dataclass once
    shifted the positional line-search settings by one)."""
    from cmad_b200 import NewtonSettings
    nw = NewtonSettings.from_reference_kwargs(max_iters=7, abs_tol=1e-9, rel_tol=1e-8, line_search_settings={
        "max evals": 11, "sufficient decrease": 3e-4, "min backtrack factor": 0.25, "max backtrack factor": 0.75})
    assert (nw.mode, nw.max_iters, nw.abs_tol, nw.rel_tol) == ("traced", 7, 1e-9, 1e-8)
    assert (nw.ls_max_evals, nw.ls_sufficient_decrease, nw.ls_min_backtrack, nw.ls_max_backtrack) == (11, 3e-4, 0.25, 0.75)
    assert nw.max_ls_evals == 0
    s = nw.to_struct()
    assert (s.max_iters, s.ls_max_evals, s.ls_c1, s.ls_bmin, s.ls_bmax) == (7, 11, 3e-4, 0.25, 0.75)
    d = NewtonSettings.from_reference_kwargs()
    assert (d.ls_max_evals, d.ls_sufficient_decrease, d.ls_min_backtrack, d.ls_max_backtrack) == (4, 1e-4, 0.5, 0.9)
