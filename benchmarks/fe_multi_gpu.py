#!/usr/bin/env python
"""Multi-GPU FE step (BASELINE.json configs[3]/[4]-style): element partition, one rank per
GPU.  Every rank assembles its element range (K3: R_e, K_e, xi; deterministic R scatter),
all-reduces the global residual R (the one exchange step of the assembly), then runs the
reverse-mode companion (K6 VJP) against a nodal adjoint vector and all-reduces the
parameter gradient - the communication pattern of an adjoint calibration step.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 benchmarks/fe_multi_gpu.py --family tet4 --div 119
Weak scaling: every rank owns its own `div`^3-cell block of a mesh N times as long (the
dof vector is the whole mesh's).  Device-side timing (CUDA events), max over ranks.
Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="tet4", choices=["tet4", "hex8"])
    ap.add_argument("--div", type=int, default=96)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--exchange", default="interface", choices=["interface", "allreduce"],
                    help="halo exchange of the interface dofs only, or all-reduce of the whole residual")
    ap.add_argument("--mixed", action="store_true",
                    help="mixed u-p formulation (examples/mixed_plastic.yaml-style, BASELINE configs[4]): "
                         "K3-mixed + pressure block + K5 over [R_u | R_p] + exchange + K6-mixed VJP")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from cmad_b200 import Parameters, active_param_ids, fe, fe_mesh, material_from_values
    from benchmarks.fe_bench import materials

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cmad_b200.comm import WORLD
    grp = WORLD if world > 1 else None       # the elements below are this rank's partition

    d = args.div
    nodes, conn = fe_mesh.structured_hex_mesh((d * world, d, d), lengths=(float(world), 1.0, 1.0))
    if args.family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    n_total = conn.shape[0]
    per = n_total // world                         # element e = i*ny*nz + ...: contiguous slabs in x
    lo, hi = rank * per, (rank + 1) * per
    # the reference's deck driver forces volume degree >= 2 on the mixed formulation
    # (cmad/cli/common.py:379-391): tet4 x 4 points (hex8 x 8 is its default anyway)
    arr = fe_mesh.block_arrays(nodes, conn[lo:hi], device=dev, mixed=args.mixed,
                               volume_degree=2 if args.mixed else None)
    arr.n_dofs = nodes.shape[0] * (4 if args.mixed else 3)
    values = materials("J2")
    const = lambda t, c: {k: const(v, c) for k, v in t.items()} if isinstance(t, dict) else c
    active = const(values, False)
    active["elastic"] = {"E": True, "nu": True}
    active["plastic"]["flow stress"] = const(values["plastic"]["flow stress"], True)
    P = Parameters(values, active, const(values, None))
    pid = active_param_ids(P)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings()
    h = 1.0 / d
    Uh = np.zeros(arr.n_dofs)
    Uh[:nodes.shape[0] * 3] = fe_mesh.synthetic_displacement(nodes, 2.0, seed=44, ramp=0.003, noise=1e-3 * h)
    if args.mixed:
        Uh[nodes.shape[0] * 3:] = 30.0 * np.random.default_rng(6).standard_normal(nodes.shape[0])
    U = torch.from_numpy(Uh).to(dev)
    lam = torch.from_numpy(np.random.default_rng(5).standard_normal(arr.n_dofs)).to(dev)
    n_e, n_ip = arr.n_elems, arr.n_ip
    xi0 = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=dev)
    r_plan = fe.mixed_r_plan(arr, device=dev) if args.mixed else \
        fe.SegmentPlan(arr.elem_eq.cpu().numpy().reshape(-1), arr.n_dofs, device=dev)
    out = {"xi": torch.empty_like(xi0),
           "R_elem": torch.empty((n_e, arr.n_basis * 3), dtype=torch.float64, device=dev),
           "K_elem": torch.empty((n_e, arr.n_basis * 3, arr.n_basis * 3), dtype=torch.float64, device=dev)}
    R = torch.empty(arr.n_dofs, dtype=torch.float64, device=dev)
    halo = fe.InterfaceExchange(arr.elem_eq, arr.n_dofs, extra_eq=arr.elem_eq_p if args.mixed else None, group=grp) \
        if args.exchange == "interface" else None
    stab = 1.0 if args.mixed else None

    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + args.warmup)]
          for k in ("t0", "asm", "red", "vjp", "end")}

    def step(k):
        ev["t0"][k].record()
        if args.mixed:
            Rm, _, xis = fe.assemble_element_block_mixed(mat, nw, arr, U, xi0, stab_mult=stab, r_plan=r_plan)
            R.copy_(Rm); out["xi"] = xis
        else:
            fe.fe_block_launch(mat, nw, arr, U, xi0, ("xi", "R_elem", "K_elem"), out)
            r_plan.sum(out["R_elem"].reshape(-1), out=R)
        ev["asm"][k].record()
        if halo is not None:
            halo.reduce(R)                                              # NCCL all-reduce of the interface dofs
        else:
            fe.reduce_residual(R, grp)                                    # NCCL all-reduce of the whole R
        ev["red"][k].record()
        pbar, _ = fe.fe_block_vjp(mat, arr, U, xi0, out["xi"], pid, lam, None, stab_mult=stab, group=grp)   # incl. all-reduce of pbar
        ev["end"][k].record()
        return pbar

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(args.warmup, args.warmup + args.steps):
        pbar = step(k)
    torch.cuda.synchronize()
    ks = range(args.warmup, args.warmup + args.steps)
    t = torch.tensor([np.mean([ev["t0"][k].elapsed_time(ev["end"][k]) for k in ks]),
                      np.mean([ev["t0"][k].elapsed_time(ev["asm"][k]) for k in ks]),
                      np.mean([ev["asm"][k].elapsed_time(ev["red"][k]) for k in ks]),
                      np.mean([ev["red"][k].elapsed_time(ev["end"][k]) for k in ks])],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        tot, asm, red, vjp = (float(x) for x in t)
        print(json.dumps({"family": args.family, "mixed": bool(args.mixed), "n_ip": n_ip, "n_gpus": world, "elements_per_gpu": n_e, "elements_total": n_total,
                          "n_dofs": arr.n_dofs, "scaling": "weak", "steps": args.steps, "warmup": args.warmup,
                          "ms_step": tot, "ms_assemble_K3_K5": asm, "ms_allreduce_R": red,
                          "ms_vjp_plus_allreduce_grad": vjp, "R_bytes": arr.n_dofs * 8,
                          "exchange": args.exchange,
                          "exchange_bytes": (halo.n_interface if halo is not None else arr.n_dofs) * 8,
                          "elements_per_s_total": n_total / tot * 1e3,
                          "grad": [float(x) for x in pbar.cpu()]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
