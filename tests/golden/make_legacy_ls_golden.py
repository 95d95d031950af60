"""Golden vectors for the imperative `newton_solve` WITH its legacy line search
(`max_ls_evals > 0`, cmad/models/nonlinear_solver.py:55-81), produced by executing the reference's
own source (see make_reference_golden.py).  Histories are chosen so that the search is active: the
near-Tresca notch material (Hosford a = 100) under large two-leg strain steps, where full Newton
steps overshoot.  Stored next to the same histories solved with `max_ls_evals = 0`.

    python tests/golden/make_legacy_ls_golden.py      (build container only)
Writes tests/golden/ref_imperative_legacy_ls.npz."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as G      # noqa: E402
import numpy as np                     # noqa: E402


def run_history(kind, F, max_ls_evals, max_iters):
    values = G.material(kind)
    model = G.SmallElasticPlastic(G.parameters(values))
    N = F.shape[2] - 1
    xi = np.zeros((N + 1, 7)); sig = np.zeros((N + 1, 6)); it = np.zeros(N + 1, int); cn = np.zeros(N + 1)
    model.set_xi_to_init_vals()
    for step in range(1, N + 1):
        model.gather_global(G.mp_U_from_F(F[:, :, step]), G.mp_U_from_F(F[:, :, step - 1]))
        it[step], cn[step] = G.newton_solve(model, max_iters=max_iters, max_ls_evals=max_ls_evals)
        xi[step] = np.concatenate([np.asarray(b) for b in model.xi()])
        model.seed_none()
        model.evaluate_cauchy()
        sig[step] = G.vec6(model.Sigma())
        model.advance_xi()
    return dict(xi=xi, sigma=sig, iters=it, cnorm=cn)


def main():
    out = {}
    for name, kind, seed, scale, nsteps in (("notch_a", "hosford_notch", 5, 6.0, 10), ("notch_b", "hosford_notch", 9, 12.0, 8),
                                            ("J2", "J2", 7, 3.0, 10)):
        F = G.two_leg_F(seed, nsteps, scale=scale, diag_only=kind.startswith("hosford"))
        out[f"{name}.F"] = F
        out[f"{name}.kind"] = np.array(kind)
        for ls in (0, 3, 8):
            r = run_history(kind, F, ls, max_iters=40)
            for k, v in r.items():
                out[f"{name}.ls{ls}.{k}"] = v
            print(name, "max_ls_evals", ls, "iters", r["iters"].tolist(), "cnorm max", float(r["cnorm"].max()), flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_imperative_legacy_ls.npz"), **out)


if __name__ == "__main__":
    main()
