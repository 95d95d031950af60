#!/usr/bin/env python
"""Secondary benchmark of the material-point kernels outside the driver's bench line
(bench.py = configs[1], J2): K1 for the other yield surfaces (BASELINE.json
configs[2]: Hosford with tangent + dC/dp, 2^23 points per GPU) and K2 (adjoint /
direct calibration objective over stored histories).  Numbers quoted in DESIGN.md
and committed under profiles/.

  python benchmarks/mp_bench.py --what k1 --yield hosford:4 --log2n 23
  python benchmarks/mp_bench.py --what k2 --log2n 20 --nsteps 20
One JSON line per measurement; CUDA-event timing after warm-up.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tree(kind):
    if kind == "J2":
        es = {"J2": 0.0}
    elif kind == "hill":
        es = {"hill": dict(zip("FGHLMN", (0.45, 0.55, 0.5, 1.4, 1.5, 1.6)))}
    elif kind.startswith("hosford"):
        es = {"hosford": {"a": float(kind.split(":")[1]) if ":" in kind else 4.0}}
    elif kind.startswith("barlat"):        # Yld2004-18p, AL7079 fit of the reference (calibrations/al7079/support.py)
        c = (0.4555, 1.0274, 0.7101, 1.3755, 0.5314, 0.8817, 1.0558, 1.1133, 0.9220,
             1.2431, 1.5438, 1.2204, 0.7632, 0.5327, 0.3015, 0.9722, 0.7399, 1.0760)
        keys = [f"{p}_{ij}" for p in ("sp", "dp") for ij in ("12", "13", "21", "23", "31", "32", "44", "55", "66")]
        es = {"barlat": {**dict(zip(keys, c)), "a": float(kind.split(":")[1]) if ":" in kind else 8.0}}
    else:
        raise ValueError(kind)
    values = {"rotation matrix": np.eye(3), "elastic": {"E": 200e3, "nu": 0.3},
              "plastic": {"effective stress": es,
                          "flow stress": {"initial yield": {"Y": 200.0},
                                          "hardening": {"voce": {"S": 200.0, "D": 20.0}}}}}
    const = lambda t, c: {k: const(v, c) for k, v in t.items()} if isinstance(t, dict) else c
    active = const(values, False)
    active["elastic"] = {"E": True, "nu": True}
    active["plastic"]["flow stress"] = const(values["plastic"]["flow stress"], True)
    return values, active, const(values, None)


def timed(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    return float(np.mean(ms)), float(np.min(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="k1", choices=["k1", "k2"])
    ap.add_argument("--yield", dest="kind", default="hosford:4")
    ap.add_argument("--log2n", type=int, default=23)
    ap.add_argument("--nsteps", type=int, default=20, help="k2: load steps of the stored history")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--diag-only", action="store_true", help="strain paths with zero shear")
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--one-pass", action="store_true", help="(default) one-pass generic kernels")
    ap.add_argument("--queue", action="store_true", help="warp-level parking kernel instead of the one-pass generic kernels")
    ap.add_argument("--cta", action="store_true", help="block-level hand-off kernel instead of the one-pass generic kernels")
    ap.add_argument("--defer", type=int, default=None, help="defer_after K (None: library default)")
    ap.add_argument("--stream", action="store_true", help="lane-refill streaming kernel instead of the one-pass generic kernels")
    ap.add_argument("--tag", default=None)
    ap.add_argument("--max-iters", type=int, default=10)
    ap.add_argument("--ls-evals", type=int, default=4)
    args = ap.parse_args()

    import torch
    from cmad_b200 import NewtonSettings, Parameters, active_param_ids, material_from_values, mp, synthetic
    from cmad_b200 import _lib as L
    dev = torch.device("cuda:0")
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    n = 1 << args.log2n
    values, active, tr = tree(args.kind)
    P = Parameters(values, active, tr)
    mat = material_from_values(values)
    pid = active_param_ids(P)
    d, d2, a = (torch.from_numpy(x).to(dev) for x in synthetic.path_params(22, 0, n, diag_only=args.diag_only))
    base = {"yield": args.kind, "n_points": n, "steps": args.steps, "warmup": args.warmup, "hbm_peak_gbs": hbm,
            "fp64_peak_tflops": mp.fp64_peak_tflops()}

    if args.what == "k1":
        nw = NewtonSettings(force_generic=args.generic, one_pass=args.one_pass, stream=args.stream, queue=args.queue, cta=args.cta, defer_after=args.defer, max_iters=args.max_iters,
                            ls_max_evals=args.ls_evals)
        outs = ("xi", "sigma", "dsig_deps", "dC_dp", "iters", "flags")
        xi = torch.zeros((7, n), dtype=torch.float64, device=dev)
        for t in (20, 40):                                   # carry the state into the plastic range
            o = mp.mp_update(mat, nw, pid, xi, synthetic.strain_at_step(d, d2, a, t), outputs=("xi",))
            xi = o["xi"]
        e = synthetic.strain_at_step(d, d2, a, 60)
        out = mp.allocate_outputs(mat, n, len(pid), outs, dev)
        ms, ms_min = timed(lambda: mp.mp_update(mat, nw, pid, xi, e, outputs=outs, out=out), args.steps, args.warmup)
        b = 784
        print(json.dumps({**base, "kernel": "K1 mp_update (xi, sigma, tangent, dC/dp)",
                          "solver": ("generic" if (args.generic or args.kind != "J2") else "j2-radial")
                          + ("" if (args.kind == "J2" and not args.generic) else (" cta" if args.cta else " queue" if args.queue else (" lane-refill" if args.stream else " one-pass"))),
                          "newton": [args.max_iters, args.ls_evals], "tag": args.tag,
                          "checksum": [float(out[k].double().sum()) for k in outs],
                          "ms_per_step": ms, "ms_min": ms_min, "updates_per_s": n / ms * 1e3,
                          "alg_bytes_per_update": b, "achieved_gbs": n * b / ms / 1e6, "frac_hbm": n * b / ms / 1e6 / hbm,
                          "plastic_fraction": float(((out["flags"] & 2) != 0).double().mean()),
                          "mean_newton_iters": float(out["iters"].double().mean())}))
        return

    # ---- K2: forward history + adjoint / direct objective ---------------------------------
    lib = L.lib()
    N = args.nsteps
    ts = [round(100 * (k + 1) / N) for k in range(N)]
    strain = torch.zeros((N + 1, 6, n), dtype=torch.float64, device=dev)
    for k, t in enumerate(ts):
        strain[k + 1] = synthetic.strain_at_step(d, d2, a, t)
    data = torch.zeros((N + 1, 9, n), dtype=torch.float64, device=dev)
    data[:, 0] = 200.0                                       # some calibration data
    xi_hist = torch.zeros((N + 1, 7, n), dtype=torch.float64, device=dev)
    na = len(pid)
    result = torch.zeros((1 + na,), dtype=torch.float64, device=dev)
    ws = torch.empty((max(int(lib.cmadx_mp_objective_workspace_bytes(C.c_int64(n), C.c_int32(na))) // 8, 1),),
                     dtype=torch.float64, device=dev)
    h = L.MpHistory()
    h.n, h.ld, h.nsteps, h.strain_comps = n, n, N, 6
    h.strain, h.data, h.xi_hist = strain.data_ptr(), data.data_ptr(), xi_hist.data_ptr()
    for k in range(9):
        h.weight[k] = 1.0
    h.result, h.workspace = result.data_ptr(), ws.data_ptr()
    nw = NewtonSettings(mode="imperative").to_struct()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    pidp = pid.ctypes.data_as(C.POINTER(C.c_int32))
    fwd = lambda: L.check(lib.cmadx_mp_forward_history(C.byref(mat), C.byref(nw), C.byref(h), stream), "fwd")
    adj = lambda: L.check(lib.cmadx_mp_objective_adjoint(C.byref(mat), pidp, na, C.byref(h), stream), "adj")
    dirs = lambda: L.check(lib.cmadx_mp_objective_direct(C.byref(mat), pidp, na, C.byref(h), stream), "dir")
    base.update({"history_steps": N, "n_active": na})
    ms, mn = timed(fwd, args.steps, args.warmup)
    print(json.dumps({**base, "kernel": "K1 x N forward history (xi only)", "ms_per_step": ms, "ms_min": mn,
                      "point_steps_per_s": n * N / ms * 1e3, "alg_bytes_per_point_step": 56 + 48 + 56,
                      "frac_hbm": n * N * 160 / ms / 1e6 / hbm}))
    for name, fn in (("K2 adjoint objective (J, dJ/dp)", adj), ("K2 direct objective (J, dJ/dp)", dirs)):
        ms, mn = timed(fn, args.steps, args.warmup)
        b = 56 + 48 + 72                                     # xi_t (+ xi_{t-1} shared with the next step), strain, data
        print(json.dumps({**base, "kernel": name, "ms_per_step": ms, "ms_min": mn,
                          "point_steps_per_s": n * N / ms * 1e3, "alg_bytes_per_point_step": b,
                          "achieved_gbs": n * N * b / ms / 1e6, "frac_hbm": n * N * b / ms / 1e6 / hbm,
                          "J": float(result[0])}))
    # ---- K2-H: adjoint pass + direct-adjoint Hessian pass (hyper-dual forward mode) ---------
    result_h = torch.zeros((1 + na + na * na,), dtype=torch.float64, device=dev)
    wsh = torch.empty((max(int(lib.cmadx_mp_hessian_workspace_bytes(C.c_int64(n), C.c_int64(n), C.c_int32(N),
                                                                     C.c_int32(na))) // 8, 1),),
                      dtype=torch.float64, device=dev)
    h.result, h.workspace = result_h.data_ptr(), wsh.data_ptr()
    hes = lambda: L.check(lib.cmadx_mp_objective_hessian(C.byref(mat), pidp, na, C.byref(h), C.c_int32(0), stream), "hess")
    ms, mn = timed(hes, args.steps, args.warmup)
    print(json.dumps({**base, "kernel": "K2-H Hessian objective (J, dJ/dp, d2J/dp2): adjoint + hyper-dual pass",
                      "ms_per_step": ms, "ms_min": mn, "point_steps_per_s": n * N / ms * 1e3,
                      "hyperdual_evals_per_point_step": na * (na + 1) // 2,
                      "J": float(result_h[0]), "H_trace": float(result_h[1 + na:].reshape(na, na).diagonal().sum())}))


if __name__ == "__main__":
    main()
