#!/usr/bin/env python
"""Multi-GPU material-point benchmarks (one rank per GPU, torchrun):

  config3 - BASELINE.json configs[2]: Hosford plasticity with tangent + dC/dp, 2^23 points per
            GPU (2^26 over 8 GPUs), weak scaling, no data-path collective;
  calib   - BASELINE.json configs[4]-style adjoint calibration gradient: every rank owns
            2^21 points x 20 load steps, forward history (K1 x N) + adjoint objective (K2) +
            ONE NCCL all-reduce of [J, grad] (1 + n_active doubles) + transform_grad.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29512 benchmarks/mp_multi_gpu.py
Device-side timing (CUDA events), max over ranks; one JSON line per case on rank 0.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=23)
    ap.add_argument("--calib-log2n", type=int, default=21)
    ap.add_argument("--nsteps", type=int, default=20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from benchmarks.mp_bench import tree
    from cmad_b200 import NewtonSettings, Parameters, active_param_ids, material_from_values, mp, synthetic
    from cmad_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- config3: Hosford K1, 2^23 points per GPU ------------------------------------------
    for kind in ("hosford:4", "hosford:100"):
        n = 1 << args.log2n
        values, active, tr = tree(kind)
        P = Parameters(values, active, tr)
        mat, pid = material_from_values(values), active_param_ids(P)
        d, d2, a = (torch.from_numpy(x).to(dev) for x in synthetic.path_params(22, rank * n, n, diag_only=True))
        nw = NewtonSettings()
        outs = ("xi", "sigma", "dsig_deps", "dC_dp", "iters", "flags")
        xi = torch.zeros((7, n), dtype=torch.float64, device=dev)
        for t in (20, 40):
            xi = mp.mp_update(mat, nw, pid, xi, synthetic.strain_at_step(d, d2, a, t), outputs=("xi",))["xi"]
        e = synthetic.strain_at_step(d, d2, a, 60)
        out = mp.allocate_outputs(mat, n, len(pid), outs, dev)
        ms = timed(lambda: mp.mp_update(mat, nw, pid, xi, e, outputs=outs, out=out))
        if rank == 0:
            print(json.dumps({"case": "config3 K1 " + kind, "n_gpus": world, "points_per_gpu": n, "scaling": "weak",
                              "ms_per_step": ms, "updates_per_s_total": world * n / ms * 1e3,
                              "frac_hbm_per_gpu": n * 784 / ms / 1e6 / hbm}), flush=True)
        del out, xi, e, d, d2, a

    # ---- calib: forward history + adjoint objective + all-reduce ---------------------------
    from cmad_b200.objectives import BatchedMPObjective
    lib = L.lib()
    n, N = 1 << args.calib_log2n, args.nsteps
    values, active, tr = tree("J2")
    P = Parameters(values, active, tr)
    mat, pid = material_from_values(values), active_param_ids(P)
    na = len(pid)
    d, d2, a = (torch.from_numpy(x).to(dev) for x in synthetic.path_params(22, rank * n, n))
    strain = torch.zeros((N + 1, 6, n), dtype=torch.float64, device=dev)
    for k in range(N):
        strain[k + 1] = synthetic.strain_at_step(d, d2, a, round(100 * (k + 1) / N))
    data = torch.zeros((N + 1, 9, n), dtype=torch.float64, device=dev); data[:, 0] = 200.0
    xi_hist = torch.zeros((N + 1, 7, n), dtype=torch.float64, device=dev)
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

    def local():
        L.check(lib.cmadx_mp_forward_history(C.byref(mat), C.byref(nw), C.byref(h), stream), "fwd")
        L.check(lib.cmadx_mp_objective_adjoint(C.byref(mat), pidp, na, C.byref(h), stream), "adj")
        return result.clone()

    from cmad_b200.comm import WORLD
    obj = BatchedMPObjective(P, local, group=WORLD if world > 1 else None)   # points are sharded by rank
    x0 = P.flat_active_values(True)
    res = {}

    def step():
        res["r"] = obj.evaluate(x0)          # K1 x N + K2 + all-reduce + D2H of 1 + n_active doubles

    ms = timed(step)
    ar = torch.zeros(1 + na, dtype=torch.float64, device=dev)
    ms_ar = timed(lambda: dist.all_reduce(ar)) if world > 1 else 0.0
    if rank == 0:
        print(json.dumps({"case": "calib J2 adjoint (K1 x N + K2 + allreduce)", "n_gpus": world, "points_per_gpu": n,
                          "history_steps": N, "scaling": "weak", "ms_per_objective": ms,
                          "ms_allreduce_alone": ms_ar, "allreduce_bytes": (1 + na) * 8,
                          "point_steps_per_s_total": world * n * N / ms * 1e3, "J": res["r"].J,
                          "grad": [float(g) for g in res["r"].grad]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
