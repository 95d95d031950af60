#!/usr/bin/env python
"""Secondary benchmark: FE element-block assembly (BASELINE.json configs[3]-style,
SURVEY.md 8d "Config 4").  Not the driver's bench line (that is bench.py /
configs[1]); this script produces the K3/K4/K5 roofline numbers quoted in
DESIGN.md and committed under profiles/.

One "step" = one ``assemble_element_block`` equivalent (R_e, K_e `vals`, xi) over
the whole block = one K3 launch (+ the bail-list launch for J2).  Timed with CUDA
events on the launching stream after warm-up; inputs are far larger than L2.

  python benchmarks/fe_bench.py --family tet4 --div 119      # 10.1 M tets
  python benchmarks/fe_bench.py --family hex8 --div 128      # 2.1 M hexes, 16.8 M IPs
Prints one JSON line per case.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# SURVEY.md 8(d) algorithmic bytes per element (index arrays not counted)
ALG_BYTES = {"tet4": {"K3": 96 + 96 + 8 + 56 + 1152 + 96 + 56, "K4": 96 + 96 + 8 + 56 + 96 + 56},
             "hex8": {"K3": 192 + 1536 + 64 + 448 + 4608 + 192 + 448, "K4": 192 + 1536 + 64 + 448 + 192 + 448}}


def materials(kind):
    const = lambda t, c: {k: const(v, c) for k, v in t.items()} if isinstance(t, dict) else c
    if kind == "J2":
        v = {"rotation matrix": np.eye(3), "elastic": {"E": 200e3, "nu": 0.3},
             "plastic": {"effective stress": {"J2": 0.0},
                         "flow stress": {"initial yield": {"Y": 200.0},
                                         "hardening": {"voce": {"S": 200.0, "D": 20.0}}}}}
    elif kind.startswith("hosford"):
        a = float(kind.split(":")[1]) if ":" in kind else 4.0
        v = {"rotation matrix": np.eye(3), "elastic": {"E": 200e3, "nu": 0.3},
             "plastic": {"effective stress": {"hosford": {"a": a}},
                         "flow stress": {"initial yield": {"Y": 200.0},
                                         "hardening": {"voce": {"S": 200.0, "D": 20.0}}}}}
    else:
        raise ValueError(kind)
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="tet4", choices=["tet4", "hex8"])
    ap.add_argument("--div", type=int, default=119, help="hex cells per axis (tet4: 6 tets per cell)")
    ap.add_argument("--yield", dest="kind", default="J2")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--load-steps", type=int, default=2, help="load steps taken before timing (state carried)")
    ap.add_argument("--variants", default="K3,K4,K5")
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--volume-degree", type=int, default=None,
                    help="quadrature override (tet4: 2 -> 4 points; hex8: 4 -> 27 points): the any-rule kernel")
    args = ap.parse_args()

    import torch
    from cmad_b200 import fe, fe_mesh, material_from_values
    from cmad_b200 import mp as mpmod
    dev = torch.device("cuda:0")
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    t0 = time.time()
    nodes, conn = fe_mesh.structured_hex_mesh((args.div,) * 3)
    if args.family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    arr = fe_mesh.block_arrays(nodes, conn, device=dev, volume_degree=args.volume_degree)
    n_e, n_ip = arr.n_elems, arr.n_ip
    h = 1.0 / args.div
    values = materials(args.kind)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(force_generic=args.generic)
    xi = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=dev)
    # uniaxial ramp 3x yield strain per unit t + nodal noise giving ~1x yield-strain perturbations
    Us = [torch.from_numpy(fe_mesh.synthetic_displacement(nodes, float(t), seed=42 + t, ramp=0.003,
                                                         noise=1e-3 * h)).to(dev)
          for t in range(1, args.load_steps + 2)]
    out = {"xi": torch.empty_like(xi), "R_elem": torch.empty((n_e, arr.n_basis * 3), dtype=torch.float64, device=dev),
           "K_elem": torch.empty((n_e, arr.n_basis * 3, arr.n_basis * 3), dtype=torch.float64, device=dev),
           "iters": torch.empty((n_e, n_ip), dtype=torch.int32, device=dev),
           "flags": torch.empty((n_e, n_ip), dtype=torch.int32, device=dev)}
    for t in range(args.load_steps):
        fe.fe_block_launch(mat, nw, arr, Us[t], xi, ("xi",), {"xi": out["xi"]})
        xi, out["xi"] = out["xi"], xi
    U = Us[args.load_steps]
    torch.cuda.synchronize()
    setup_s = time.time() - t0

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ev[0].record()
        for k in range(args.steps):
            fn()
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
        return float(np.mean(ms)), float(np.min(ms))

    # algorithmic bytes of the block's actual rule (the table above holds the default rules)
    n_b = arr.n_basis
    rd = 24 * n_b + n_ip * n_b * 24 + n_ip * 8 + n_ip * 56
    ALG_BYTES[args.family] = {"K3": rd + (3 * n_b) ** 2 * 8 + 24 * n_b + n_ip * 56, "K4": rd + 24 * n_b + n_ip * 56}
    base = {"family": args.family, "n_ip": n_ip, "n_elems": n_e, "n_ips": n_e * n_ip, "yield": args.kind,
            "solver": "generic" if args.generic or args.kind != "J2" else "j2-radial",
            "steps": args.steps, "warmup": args.warmup, "setup_s": round(setup_s, 1), "hbm_peak_gbs": hbm}
    variants = args.variants.split(",")
    if "K3" in variants:
        o3 = {k: out[k] for k in ("xi", "R_elem", "K_elem", "iters", "flags")}
        l0 = mpmod.launch_count()
        ms, ms_min = timed(lambda: fe.fe_block_launch(mat, nw, arr, U, xi, tuple(o3), o3))
        launches = (mpmod.launch_count() - l0) // (args.steps + args.warmup)
        fl = out["flags"]
        it = out["iters"]
        b = ALG_BYTES[args.family]["K3"]
        print(json.dumps({**base, "kernel": "K3 fe_block (R_e, K_e, xi)", "ms_per_step": ms, "ms_min": ms_min,
                          "elements_per_s": n_e / ms * 1e3, "ip_updates_per_s": n_e * n_ip / ms * 1e3,
                          "alg_bytes_per_elem": b, "achieved_gbs": n_e * b / ms / 1e6,
                          "frac_hbm": n_e * b / ms / 1e6 / hbm, "launches_per_step": launches,
                          "plastic_fraction": float(((fl & 2) != 0).double().mean()),
                          "mean_newton_iters": float(it.double().mean()),
                          "bailed_elements": mpmod.debug_bail_count() if base["solver"] == "j2-radial" else 0}))
    if "MIX" in variants:
        # mixed u-p formulation: K3 with the momentum stress dev(cauchy) - p I + the pressure-block
        # kernel; algorithmic bytes = K3's + p_e, h, R_p and the (u,p), (p,u), (p,p) streams
        n_b = arr.n_basis
        arr_m = fe_mesh.block_arrays(nodes, conn, device=dev, mixed=True, volume_degree=args.volume_degree)
        rng = np.random.default_rng(5)
        Um = torch.cat([U, torch.from_numpy(-60.0 + 5.0 * rng.standard_normal(nodes.shape[0])).to(dev)])
        l0 = mpmod.launch_count()
        ms, ms_min = timed(lambda: fe.assemble_element_block_mixed(mat, nw, arr_m, Um, xi))
        launches = (mpmod.launch_count() - l0) // (args.steps + args.warmup)
        b = ALG_BYTES[args.family]["K3"] + 8 * (n_b + 1 + n_b + 2 * 3 * n_b * n_b + n_b * n_b)
        print(json.dumps({**base, "kernel": "K3-mixed fe_block u-p (R_u, R_p, K_uu, K_up, K_pu, K_pp, xi) + atomic R",
                          "ms_per_step": ms, "ms_min": ms_min, "elements_per_s": n_e / ms * 1e3,
                          "alg_bytes_per_elem": b, "achieved_gbs": n_e * b / ms / 1e6,
                          "frac_hbm": n_e * b / ms / 1e6 / hbm, "launches_per_step": launches}))
        del arr_m, Um
    if "K4" in variants:
        o4 = {k: out[k] for k in ("xi", "R_elem")}
        ms, ms_min = timed(lambda: fe.fe_block_launch(mat, nw, arr, U, xi, tuple(o4), o4))
        b = ALG_BYTES[args.family]["K4"]
        print(json.dumps({**base, "kernel": "K4 fe_block residual-only (R_e, xi)", "ms_per_step": ms, "ms_min": ms_min,
                          "elements_per_s": n_e / ms * 1e3, "alg_bytes_per_elem": b,
                          "achieved_gbs": n_e * b / ms / 1e6, "frac_hbm": n_e * b / ms / 1e6 / hbm}))
    if "K5" in variants:
        eq = arr.elem_eq.cpu().numpy()
        t1 = time.time()
        r_plan = fe.SegmentPlan(eq.reshape(-1), arr.n_dofs, device=dev)
        plan_s = time.time() - t1
        Rg = torch.empty(arr.n_dofs, dtype=torch.float64, device=dev)
        Rflat = out["R_elem"].reshape(-1)
        ms, ms_min = timed(lambda: r_plan.sum(Rflat, out=Rg))
        n_items = Rflat.numel()
        b = n_items * (8 + 4) + arr.n_dofs * 16
        print(json.dumps({**base, "kernel": "K5 deterministic R scatter (segment sum)", "ms_per_step": ms,
                          "ms_min": ms_min, "items": n_items, "segments": arr.n_dofs, "plan_build_s": round(plan_s, 2),
                          "alg_bytes": b, "achieved_gbs": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / hbm}))
        # atomic alternative fused in K3's epilogue (non-deterministic), for comparison
        oa = {"xi": out["xi"], "R_global": Rg}
        def atom():
            Rg.zero_()
            fe.fe_block_launch(mat, nw, arr, U, xi, ("xi", "R_global"), oa)
        ms, ms_min = timed(atom)
        print(json.dumps({**base, "kernel": "K4 + atomic R scatter (memset + fused atomics)", "ms_per_step": ms,
                          "ms_min": ms_min, "elements_per_s": n_e / ms * 1e3}))
        if args.family == "tet4" and n_e <= 3_000_000 or args.family == "hex8" and n_e <= 300_000:
            t1 = time.time()
            ur, uc, scatter = fe_mesh.coo_dedup(eq)
            k_plan = fe.SegmentPlan(scatter, len(ur), device=dev)
            plan_s = time.time() - t1
            Kd = torch.empty(len(ur), dtype=torch.float64, device=dev)
            vals = out["K_elem"].reshape(-1)
            ms, ms_min = timed(lambda: k_plan.sum(vals, out=Kd))
            b = vals.numel() * (8 + 4) + len(ur) * 16
            print(json.dumps({**base, "kernel": "K5 deterministic COO dedup (segment sum)", "ms_per_step": ms,
                              "ms_min": ms_min, "items": vals.numel(), "segments": len(ur),
                              "plan_build_s": round(plan_s, 2), "alg_bytes": b, "achieved_gbs": b / ms / 1e6,
                              "frac_hbm": b / ms / 1e6 / hbm}))


if __name__ == "__main__":
    main()
