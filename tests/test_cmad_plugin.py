"""The reference-side adapter (cmad_b200/cmad_plugin.py) and the XLA FFI handlers' source.

CPU, build container only (needs /root/reference; skipped elsewhere): a REAL `FEProblem` is
built by the reference's own code (on the NumPy `jax` stand-in of tests/golden/jaxshim),
`cmad_plugin.install()` hooks `cmad.fem.assembly.assemble_element_block`, and the reference's
`assemble_global` then runs through the hook with the reference's own `geometry_cache` /
`u_gather_eq_by_block` arrays, unchanged, routed to a test backend (the CPU oracle - the CUDA
library cannot run here).  Its `(K, R, xi)` must equal what the un-hooked reference computes.
This proves the argument plumbing of the drop-in, not the kernels.

GPU: the same adapter entry point with `TorchBackend` (ctypes -> C-ABI -> CUDA), fed the
reference's arrays as stored in tests/golden/ref_fe_block.npz through duck-typed stand-ins of
`FEProblem` / `FEKernelArrays`, against the reference's own outputs in the fixture."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"

_SCRIPT = r'''
import json, os, sys, types
import numpy as np
ROOT, HERE = sys.argv[1], os.path.join(sys.argv[1], "tests", "golden")
sys.path[:0] = [os.path.join(HERE, "jaxshim"), "/root/reference", ROOT, HERE]
import make_reference_fe_block_golden as gen          # enters the reference (stubs for absent optional deps)
import cmad.fem.assembly as assembly
from cmad.fem.assembly import params_by_block_from_models
from cmad.io import registry
import cmad_b200.cmad_plugin as plugin
from cmad_b200.material import NewtonSettings
from oracle import fe_oracle, oracle_c as oc
from materials import material

class OracleBackend:
    """Consumes exactly what the FFI call would receive; computes with the CPU oracle."""
    calls = 0
    def _prob(self, newton):
        return oc.describe(material("J2"), None, newton_mode="traced", strain_comps=9, max_iters=newton.max_iters,
                           abs_tol=newton.abs_tol, rel_tol=newton.rel_tol)
    def fe_block(self, b, mat, newton):
        OracleBackend.calls += 1
        assert np.asarray(b.elem_eq).dtype == np.int32
        o = fe_oracle.assemble_block(self._prob(newton), np.asarray(b.elem_eq), np.asarray(b.U), np.asarray(b.xi_prev),
                                     np.asarray(b.grad_N), np.asarray(b.det), np.asarray(b.quad_w))
        return o["R_elem"], o["K_elem"], o["xi"]
    def fe_block_mixed(self, b, mat, newton):
        OracleBackend.calls += 1
        o = fe_oracle.assemble_block_mixed(self._prob(newton), np.asarray(b.elem_eq), np.asarray(b.elem_eq_p),
                                           np.asarray(b.U), np.asarray(b.xi_prev), np.asarray(b.grad_N), np.asarray(b.N),
                                           np.asarray(b.det), np.asarray(b.quad_w), np.asarray(b.h), b.stab_mult)
        return o["R_u"], o["R_p"], o["K_uu"], o["K_up"], o["K_pu"], o["K_pp"], o["xi"]

out = {}
for family, mixed in (("tet4", False), ("hex8", True)):
    mesh, fp = gen._problem(family, mixed, (1, 1, 1))
    ka = fp.kernel_arrays
    pb = params_by_block_from_models(fp)
    n = int(fp.dof_map.num_total_dofs)
    n_u = int(fp.dof_map.block_offsets[1]) if mixed else n
    rng = np.random.default_rng(3)
    U = np.zeros(n); U[0:n_u:3] = 0.003 * np.asarray(mesh.nodes)[:, 0]; U[:n_u] += 3e-4 * rng.standard_normal(n_u)
    if mixed:
        U[n_u:] = -50.0 + 20.0 * rng.standard_normal(n - n_u)
    n_e = mesh.connectivity.shape[0]
    n_ip = int(np.asarray(ka.geometry_cache["all"].shared.quad_w).shape[0])
    xi_prev = np.zeros((n_e, n_ip, 7))
    K0, R0, xi0 = assembly.assemble_global(fp, ka, pb, U, U, t=0.5, xi_prev_by_block={"all": xi_prev})
    assert plugin.supports(fp.models_by_block["all"], fp.gr, pb["all"])
    assert plugin.install(backend=OracleBackend())
    before = OracleBackend.calls
    K1, R1, xi1 = assembly.assemble_global(fp, ka, pb, U, U, t=0.5, xi_prev_by_block={"all": xi_prev})
    hooked = OracleBackend.calls - before
    cls = registry.resolve_model("small_elastic_plastic")
    plugin.uninstall()
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())
    out[f"{family}.{'mixed' if mixed else 'disp'}"] = {
        "hooked_calls": hooked, "rel_K": rel(K1.data, K0.data), "rel_R": rel(R1, R0), "rel_xi": rel(xi1["all"], xi0["all"]),
        "alpha_max": float(np.asarray(xi0["all"])[..., 6].max()),
        "registry_subclass": bool(issubclass(cls, plugin.B200Model) and cls.__mro__[2].__name__ == "SmallElasticPlastic")}
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "cmad")), reason="needs the reference tree (build container only)")
def test_hooked_reference_assembly_equals_unhooked_reference(tmp_path):
    script = tmp_path / "plugin_vs_reference.py"
    script.write_text(_SCRIPT)
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(next(l for l in r.stdout.splitlines() if l.startswith("RESULT "))[7:])
    assert set(res) == {"tet4.disp", "hex8.mixed"}
    for case, v in res.items():
        assert v["hooked_calls"] == 1, (case, v)            # the block went through the adapter
        assert v["alpha_max"] > 0, case                      # ... in the plastic range
        assert v["rel_K"] < 1e-10 and v["rel_R"] < 1e-10 and v["rel_xi"] < 1e-10, (case, v)
        assert v["registry_subclass"], case


def test_xla_ffi_handlers_compile_against_the_c_abi():
    """`g++ -fsyntax-only` over cmad_b200/xla/cmad_b200_xla.cc with the stand-in xla/ffi header of
    tests/xla_mock (JAX is not installable here): struct fields and entry-point signatures are
    checked against include/cmad_b200.h, every handler's parameter list against its binding."""
    from cmad_b200 import xla
    assert not xla.available() or os.path.isdir(__import__("jax").ffi.include_dir())
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "xla_mock"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"), xla.SOURCE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    src = open(xla.SOURCE).read()
    for target, symbol in xla.TARGETS.items():
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({symbol}," in src, target
    if not xla.available():
        with pytest.raises(RuntimeError):
            xla.build()


def test_supports_and_material_mapping():
    import cmad_b200.cmad_plugin as plugin
    from tests.golden.materials import material

    class SmallElasticPlastic:                       # duck-typed stand-in: only the class name and params matter
        def __init__(self, values):
            self.parameters = types.SimpleNamespace(values=values)

    class Elastic(SmallElasticPlastic):
        pass
    Elastic.__mro__  # noqa: B018
    for kind in ("J2", "hill", "hosford"):
        m = SmallElasticPlastic(material(kind))
        assert plugin.supports(m)
        mat = plugin.material_of(m, m.parameters.values)
        assert mat.yield_ == {"J2": 0, "hill": 1, "hosford": 2}[kind]
    bad = material("J2")
    bad["plastic"]["effective stress"] = {"barlat": {"a": 8.0}}
    assert not plugin.supports(SmallElasticPlastic(bad))
    assert not plugin.supports(object())
    assert not plugin.supports(SmallElasticPlastic(material("J2")), gr=object())
    nw = plugin.newton_of(object())
    assert (nw.max_iters, nw.abs_tol, nw.rel_tol, nw.ls_max_evals) == (20, 1e-12, 1e-12, 4)


# ------------------------------------------------------------------------------------- GPU
def _duck_problem(G, case):
    """Stand-ins of FEProblem / FEKernelArrays holding the REFERENCE'S arrays of the fixture."""
    g = lambda k: G[f"{case}.{k}"]
    mixed = case.endswith("mixed")
    n_f = 2 if mixed else 1
    per_elem = types.SimpleNamespace(iso_jac_det=g("iso_jac_det"), element_size=g("element_size"),
                                     field_grad_N_phys_per_block=tuple(g(f"grad_N_phys.{r}") for r in range(n_f)))
    shared = types.SimpleNamespace(quad_w=g("quad_w"), field_N_per_block=tuple(g(f"N.{r}") for r in range(n_f)))
    ka = types.SimpleNamespace(geometry_cache={"all": types.SimpleNamespace(per_elem=per_elem, shared=shared)},
                               u_gather_eq_by_block={"all": tuple(g(f"u_gather_eq.{f}") for f in range(n_f))})
    gr = types.SimpleNamespace(_mixed=mixed, _stabilization_multiplier=1.0)

    class SmallElasticPlastic:
        pass
    fp = types.SimpleNamespace(gr=gr, models_by_block={"all": SmallElasticPlastic()},
                               dof_map=types.SimpleNamespace(num_total_dofs=int(g("n_dofs"))),
                               forcing_fns_by_block_idx=None)
    return fp, ka


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["hex8.disp", "hex8.mixed", "tet4.disp", "tet4.mixed"])
def test_adapter_over_cuda_vs_reference_outputs(cuda_device, case):
    import cmad_b200.cmad_plugin as plugin
    from tests.golden.materials import material
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_fe_block.npz"))
    fp, ka = _duck_problem(G, case)
    be = plugin.TorchBackend(cuda_device)
    for s in range(2):
        U, xi_prev = G[f"{case}.asm{s}.U"], G[f"{case}.asm{s}.xi_prev"]
        R, vals, xi = plugin.assemble_element_block_b200(fp, ka, {"all": material("J2")}, "all", U, U, 0.0, xi_prev,
                                                         backend=be)
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        assert rel(R, G[f"{case}.asm{s}.R_block"]) < 1e-10
        assert rel(vals, G[f"{case}.asm{s}.vals"]) < 1e-10
        assert rel(xi, G[f"{case}.asm{s}.xi"]) < 1e-10
