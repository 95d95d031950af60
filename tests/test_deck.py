"""YAML deck front end (cmad_b200/deck.py): BASELINE.json configs[0] / [3] / [4] decks are READ, not
restated in test code.  tests/decks/*.yaml restate the reference's example decks in its own schema
(the GPU box has no reference tree); where the tree exists the reference's OWN files must parse to
the same problems.  CPU: parsing + the deck-driven quasi-static drive over the oracle assembler;
GPU: `deck.run_primal` over the CUDA kernels (the `cmad primal deck.yaml` equivalent)."""
import os
import shutil

import numpy as np
import pytest

from cmad_b200 import deck, fe_driver as drv, fe_mesh
from oracle import fe_oracle, oracle_c as oc
from tests.test_fe_driver import uniaxial_cube, uniaxial_flow_stress

HERE = os.path.dirname(os.path.abspath(__file__))
DECKS = os.path.join(HERE, "decks")
REF_EXAMPLES = "/root/reference/examples"
NAMES = ("elastic_plastic_uniaxial", "mixed_plastic", "notch_hosford")


def _same_problem(a: deck.DeckProblem, b: deck.DeckProblem):
    assert (a.name, a.block, a.mixed, a.stab_mult, a.volume_degree, a.qoi) == \
        (b.name, b.block, b.mixed, b.stab_mult, b.volume_degree, b.qoi)
    assert np.array_equal(a.conn, b.conn) and np.array_equal(a.nodes, b.nodes)
    assert a.local_newton == b.local_newton and a.nonlinear == b.nonlinear
    assert np.array_equal(a.t_schedule, b.t_schedule)
    assert a.bc_entries == b.bc_entries
    assert np.array_equal(a.parameters._flat_values, b.parameters._flat_values)
    assert np.array_equal(a.parameters.active_idx, b.parameters.active_idx)


def test_restated_decks_parse():
    p = deck.fe_problem_from_deck(os.path.join(DECKS, "elastic_plastic_uniaxial.yaml"))
    assert p.conn.shape == (512, 8) and p.nodes.shape == (729, 3) and p.block == "all" and not p.mixed
    assert list(p.parameters.active_idx) == [0, 1, 2, 3, 4]     # E, nu, D, S, Y: `{J2: {}}` holds no leaf (sorted-key flatten order)
    v = p.values
    assert (v["elastic"]["E"], v["elastic"]["nu"]) == (200000.0, 0.3)
    assert v["plastic"]["effective stress"] == {"J2": {}} and np.array_equal(v["rotation matrix"], np.eye(3))
    assert (p.local_newton.max_iters, p.local_newton.abs_tol, p.local_newton.ls_max_evals) == (20, 1e-12, 4)
    assert p.nonlinear["max iters"] == 10 and p.nonlinear["line search"]["max evals"] == 4
    assert np.allclose(p.t_schedule, np.linspace(0, 1, 6)) and p.qoi == "fe_displacement_l2"
    # the BCs the driver tests build by hand (tests/test_fe_driver.py:uniaxial_cube)
    _, _, bcs_ref, _, _ = uniaxial_cube(8, "hex8")
    bcs = p.dirichlet_bcs()
    order = np.argsort(bcs_ref.indices, kind="stable")
    assert np.array_equal(bcs.indices, bcs_ref.indices[order])
    assert np.allclose(bcs.values(0.6), bcs_ref.values(0.6)[order])
    m = deck.fe_problem_from_deck(os.path.join(DECKS, "mixed_plastic.yaml"))
    assert m.mixed and m.volume_degree == 2 and m.stab_mult == 1.0 and m.local_newton.max_iters == 100
    assert m.nonlinear["abs tol"] == 1e-9 and len(m.parameters.active_idx) == 0
    n = deck.fe_problem_from_deck(os.path.join(DECKS, "notch_hosford.yaml"))
    assert n.conn.shape == (1550, 4) and n.block == "block_1"
    assert (n.local_newton.max_iters, n.local_newton.ls_max_evals) == (500, 100)
    assert n.values["plastic"]["effective stress"] == {"hosford": {"a": 100.0}}
    assert n.material().hosford_a == 100.0 and np.allclose(n.t_schedule, [0, 1, 2, 3, 4])


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("name", NAMES)
def test_reference_decks_parse_to_the_same_problem(name):
    """The reference's own examples/<name>.yaml, read as is (its notch.exo through SciPy)."""
    _same_problem(deck.fe_problem_from_deck(os.path.join(REF_EXAMPLES, name + ".yaml")),
                  deck.fe_problem_from_deck(os.path.join(DECKS, name + ".yaml")))


def test_deck_errors():
    with pytest.raises(ValueError):
        deck.scalar_expression("__import__('os').system('true')")
    with pytest.raises(NotImplementedError):
        deck.coordinate_side_nodes(np.zeros((2, 3)), "notch_surface")
    with pytest.raises(ValueError):
        deck.split_parameters({"E": {"value": 1.0, "transform": {"exp": 2}}})
    v, a, t = deck.split_parameters({"E": {"value": 2, "active": True, "transform": {"log": 2.0}}, "Q": [[1, 0], [0, 1]]})
    assert v["E"] == 2.0 and a == {"E": True, "Q": False} and t["E"].tolist() == [2.0] and v["Q"].shape == (2, 2)


def test_deck_driven_uniaxial_drive_on_the_oracle(tmp_path):
    """configs[0] through the deck path at 2^3 cells (the mesh size is the only edit): every
    point's terminal sigma_xx is the analytical J2 + Voce flow stress."""
    src = open(os.path.join(DECKS, "elastic_plastic_uniaxial.yaml")).read().replace("cube_hex_8.exo", "cube_hex_2.exo")
    path = tmp_path / "deck.yaml"
    path.write_text(src)
    p = deck.fe_problem_from_deck(str(path))
    arr = p.arrays()
    pattern, scatter = p.pattern(arr)
    nw = p.local_newton
    prob = oc.describe(p.values, None, newton_mode="traced", strain_comps=9, max_iters=nw.max_iters,
                       abs_tol=nw.abs_tol, rel_tol=nw.rel_tol)
    rec = {}

    def assemble(U, xi_prev):
        o = fe_oracle.assemble_block(prob, arr.elem_eq.numpy(), U, xi_prev, arr.grad_N.numpy(), arr.det.numpy(),
                                     arr.quad_w.numpy())
        rec["sigma"] = o["sigma"]
        return o["R"], fe_oracle.coo_dedup_sum(o["K_elem"].reshape(-1), scatter, len(pattern.rows)), o["xi"]
    U_steps, xi, _, logs = drv.fe_quasistatic_drive(assemble, pattern, p.dirichlet_bcs(), np.zeros(arr.n_dofs),
                                                    np.zeros((arr.n_elems, arr.n_ip, 7)), p.t_schedule, p.nonlinear)
    sig_ref, alpha_ref = uniaxial_flow_stress(0.003)
    assert np.abs(rec["sigma"][:, :, 0] - sig_ref).max() < 1e-8 * sig_ref
    assert np.abs(xi[:, :, 6] - alpha_ref).max() < 1e-10
    assert all(l.iters <= 10 for l in logs)


@pytest.mark.gpu
def test_cuda_primal_from_the_configs0_deck(cuda_device):
    """`cmad primal elastic_plastic_uniaxial.yaml` on the B200 path: 8^3 hexes, 5 load steps."""
    import torch
    from cmad_b200 import fe
    p = deck.fe_problem_from_deck(os.path.join(DECKS, "elastic_plastic_uniaxial.yaml"))
    U_steps, xi, J, logs = deck.run_primal(p, cuda_device)
    sig_ref, alpha_ref = uniaxial_flow_stress(0.003)
    assert float((xi[:, :, 6] - alpha_ref).abs().max()) < 1e-10
    ux = U_steps[-1].reshape(-1, 3)[:, 0]
    assert np.abs(ux - 0.003 * p.nodes[:, 0]).max() < 1e-12
    arr = p.arrays(cuda_device)
    out = fe.fe_block_launch(p.material(), p.local_newton, arr, torch.from_numpy(U_steps[-1]).to(cuda_device), xi,
                             ("xi", "sigma"))
    assert float((out["sigma"][:, :, 0] - sig_ref).abs().max()) < 1e-7 * sig_ref
    assert J > 0 and all(l.iters <= 10 for l in logs)


@pytest.mark.gpu
def test_cuda_primal_from_the_mixed_and_notch_decks(cuda_device):
    """mixed_plastic.yaml (KA6: sigma_xx analytic, p = -sigma/3) and notch_hosford.yaml (native size,
    Hosford a = 100, 500 / 100 local settings) run from their decks."""
    m = deck.fe_problem_from_deck(os.path.join(DECKS, "mixed_plastic.yaml"))
    U_steps, xi, _, logs = deck.run_primal(m, cuda_device)
    sig_ref, alpha_ref = uniaxial_flow_stress(0.005)
    assert float((xi[:, :, 6] - alpha_ref).abs().max()) < 1e-6 * alpha_ref + 1e-9
    pr = U_steps[-1][3 * m.nodes.shape[0]:]
    assert np.abs(pr + sig_ref / 3.0).max() < 1e-5 * sig_ref
    n = deck.fe_problem_from_deck(os.path.join(DECKS, "notch_hosford.yaml"))
    U_steps, xi, _, logs = deck.run_primal(n, cuda_device)
    assert all(np.isfinite(U).all() for U in U_steps) and float(xi[:, :, 6].max()) > 0
    assert all(l.residual_norms[-1] < 1e-6 for l in logs)
