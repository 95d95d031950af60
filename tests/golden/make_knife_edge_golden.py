"""How well-defined are the Newton counts / branch flags of the knife-edge material
(examples/notch_hosford.yaml: Hosford a = 100, 500 iterations, 100 line-search probes)?

Runs the REFERENCE'S OWN `make_newton_solve` (on the NumPy `jax` stand-in, as
make_reference_golden.py does) on the inputs of the `hosford_notch.notch` fixture, once more with
every strain entry moved by ONE ULP (x (1 + 2^-52) and x (1 - 2^-52)), and records the counts and
exit flags.  Where the reference's own answers change under a one-ulp change of its input, no two
implementations whose arithmetic differs at rounding level can be expected to agree; the
fixture lets the tests state that fraction instead of asserting it.

    python tests/golden/make_knife_edge_golden.py     (build container only)
Writes tests/golden/ref_knife_edge.npz."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as G      # noqa: E402  (enters the reference)
import numpy as np                     # noqa: E402
from jax import _core                  # noqa: E402


def run(case="hosford_notch.notch"):
    kind, key = case.split(".")
    TR = np.load(os.path.join(HERE, "ref_traced_newton.npz"))
    values = G.material(kind)
    P = G.parameters(values)
    model = G.SmallElasticPlastic(P)
    solve = G.make_newton_solve(model._residual, **G.NEWTON[key])
    yf = G.yield_fun_of(model, values)
    gu, xp = TR[f"{case}.grad_u"], TR[f"{case}.xi_prev"]
    out = {"iters": TR[f"{case}.iters"], "flags": TR[f"{case}.flags"]}
    pl = lambda f: bool(f > 1e-14 or abs(f) < 1e-14)      # noqa: E731
    for tag, fac in (("base", 1.0), ("up", 1.0 + 2.0 ** -52), ("down", 1.0 - 2.0 ** -52)):
        it = np.zeros(gu.shape[:2], dtype=np.int64); fl = np.zeros(gu.shape[:2], dtype=np.int64)
        for s in range(gu.shape[0]):
            for i in range(gu.shape[1]):
                U = G.mp_U_from_F(np.eye(3) + (gu[s, i] * fac).reshape(3, 3))
                x0 = [xp[s, i, :6].copy(), xp[s, i, 6:].copy()]
                xi = solve(x0, P.values, U, U)
                it[s, i] = int(_core.WHILE_LOG[-1][1][0])
                xi_np = [np.asarray(x) for x in xi]
                _, f0, _ = yf(x0, x0, P.values, U, U)
                _, f1, _ = yf(xi_np, x0, P.values, U, U)
                fl[s, i] = (1 if pl(float(f0)) else 0) | (2 if pl(float(f1)) else 0)
        out[f"iters_{tag}"], out[f"flags_{tag}"] = it, fl
        print(tag, "iters", it.ravel().tolist(), "flags", fl.ravel().tolist(), flush=True)
    assert np.array_equal(out["iters_base"], out["iters"]) and np.array_equal(out["flags_base"], out["flags"])
    np.savez_compressed(os.path.join(HERE, "ref_knife_edge.npz"), **out)


if __name__ == "__main__":
    run()
