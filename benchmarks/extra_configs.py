"""Driver-visible measurements of BASELINE.json configs[2..4] and of the path's collectives.

``bench.py`` (the driver's contract, headline = configs[1]) calls :func:`run_all` after its own
timed region and attaches the result as ``extra.configs`` to its ONE JSON line.  Every entry is
measured like the headline: W >= 3 warm-up steps, a barrier + ``torch.cuda.synchronize()`` on
both sides, CUDA events on the launching stream, the MAX over ranks, SM clock / throttle reasons
sampled during the timed region, algorithmic bytes against the measured HBM peak
and FP64 flops (counted by ncu on the same workload, file named in the entry) against the
measured DFMA peak.

  hosford_a4 / hosford_a100   configs[2]: Hosford K1 (xi, cauchy, tangent, dC/dp), 2^23 points per GPU
  fe_k3_tet4 / fe_k3_hex8     configs[3]: one assemble_element_block (R_e, K_e, xi), 10.1 M tets /
                              216^3 hexes (N = 1 only: no collective in it)
  fe_adjoint_mixed_*          configs[4]: one adjoint-calibration step on a mixed u-p mesh,
                              elements partitioned over the ranks: K3-mixed + pressure block + K5 +
                              interface exchange of R (NCCL) + K6-mixed VJP + all-reduce of the
                              parameter gradient (NCCL); exchange and all-reduce timed separately
  mp_objective                the MP calibration objective: fused forward history + K2 adjoint +
                              ONE all-reduce of (J, grad) = 48 bytes (NCCL)
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# FP64 flops per unit (2 x DFMA + DMUL + DADD thread instructions, predicated on) from ncu
# captures of these exact workloads; the source file is named next to each number.
FP64_FLOPS = {
    # (flops per unit, source): ncu counts of the timed region of these exact workloads
    # (benchmarks/count_flops.sh -> profiles/r2_flops_<name>.csv, reduced by profiles/tools/ncu_region_flops.py)
}
_flops_file = os.path.join(ROOT, "profiles", "r2_fp64_flops.json")
if os.path.exists(_flops_file):
    import json as _json
    FP64_FLOPS.update({k: (v["flops_per_unit"], v["source"]) for k, v in _json.load(open(_flops_file)).items()})


def _const(t, c):
    return {k: _const(v, c) for k, v in t.items()} if isinstance(t, dict) else c


def _tree(es, elastic, Y, S, D):
    values = {"rotation matrix": np.eye(3), "elastic": dict(elastic),
              "plastic": {"effective stress": es,
                          "flow stress": {"initial yield": {"Y": Y}, "hardening": {"voce": {"S": S, "D": D}}}}}
    active = _const(values, False)
    active["elastic"] = {k: True for k in elastic}
    active["plastic"]["flow stress"] = _const(values["plastic"]["flow stress"], True)
    return values, active, _const(values, None)


class Ctx:
    def __init__(self, dev, rank, world, local, hbm_peak, fp64_peak, sampler_cls, steps, warmup):
        self.dev, self.rank, self.world, self.local = dev, rank, world, local
        self.hbm, self.fp64 = hbm_peak, fp64_peak
        self.sampler_cls = sampler_cls
        self.steps, self.warmup = steps, max(warmup, 3)

    def timed(self, fn, marks=()):
        """W warm-ups, then K timed calls between barrier + synchronize; CUDA events; max over
        ranks.  ``fn(rec)`` may call ``rec(name)`` to split a step: returns (ms_per_step,
        {name: ms of the segment ending at that mark}, clocks)."""
        import torch
        import torch.distributed as dist
        K, W = self.steps, self.warmup
        noop = lambda name: None
        for _ in range(W):
            fn(noop)
        torch.cuda.synchronize(self.dev)
        names = list(marks)
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        evm = {m: [torch.cuda.Event(enable_timing=True) for _ in range(K)] for m in names}
        sampler = self.sampler_cls(self.local)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)
        sampler.start()
        torch.cuda.nvtx.range_push("cmadx_timed")      # lets `ncu --nvtx --nvtx-include "cmadx_timed/"` count this region only
        ev0[0].record()
        for k in range(K):
            fn(lambda name, k=k: evm[name][k].record())
            ev0[k + 1].record()
        torch.cuda.synchronize(self.dev)
        torch.cuda.nvtx.range_pop()
        if self.world > 1:
            dist.barrier()
        clocks = sampler.stop()
        seg = {}
        for k in range(K):
            prev = ev0[k]
            for m in names:
                seg[m] = seg.get(m, 0.0) + prev.elapsed_time(evm[m][k]) / K
                prev = evm[m][k]
        vals = [ev0[0].elapsed_time(ev0[K]) / K] + [seg[m] for m in names]
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t]
        return vals[0], dict(zip(names, vals[1:])), clocks

    def fracs(self, units_per_gpu, ms, alg_bytes, key):
        out = {"alg_bytes_per_unit": alg_bytes,
               "hbm_gbs": units_per_gpu * alg_bytes / ms / 1e6,
               "frac_hbm": units_per_gpu * alg_bytes / ms / 1e6 / self.hbm}
        if key in FP64_FLOPS:
            fl, src = FP64_FLOPS[key]
            out.update({"fp64_flops_per_unit": fl, "fp64_flops_source": src,
                        "fp64_tflops": units_per_gpu * fl / ms / 1e9,
                        "frac_fp64": units_per_gpu * fl / ms / 1e9 / self.fp64})
            out["binding_roofline"] = "hbm" if out["frac_hbm"] >= out["frac_fp64"] else "fp64"
            out["frac_binding"] = max(out["frac_hbm"], out["frac_fp64"])
        return out


# ------------------------------------------------------------------------------ configs[2]
def hosford(ctx: Ctx, a: float, log2n: int = 23):
    """Hosford K1 with tangent + dC/dp (SURVEY 8d "Config 3").  a = 4: the reference's test exponent
    with the J2AnalyticalProblem constants and make_newton_solve defaults; a = 100: the constants
    and local-solver settings of examples/notch_hosford.yaml:30-42 (500 iterations, 100 probes)."""
    import torch
    from cmad_b200 import NewtonSettings, Parameters, active_param_ids, material_from_values, mp, synthetic
    n = 1 << log2n
    if a == 100:
        values, active, tr = _tree({"hosford": {"a": 100.0}}, {"E": 1000.0, "nu": 0.25}, 2.0, 10.0, 2.0)
        nw = NewtonSettings(max_iters=500, abs_tol=1e-12, rel_tol=1e-12, ls_max_evals=100)
        ey, newton = 2e-3, "notch_hosford.yaml: 500 iters, 1e-12, line search 100 evals"
    else:
        values, active, tr = _tree({"hosford": {"a": float(a)}}, {"E": 200e3, "nu": 0.3}, 200.0, 200.0, 20.0)
        nw = NewtonSettings()
        ey, newton = 1e-3, "make_newton_solve defaults: 10 iters, 1e-14, line search 4 evals"
    P = Parameters(values, active, tr)
    mat, pid = material_from_values(values), active_param_ids(P)
    d, d2, amp = (torch.from_numpy(x).to(ctx.dev)
                  for x in synthetic.path_params(22, ctx.rank * n, n, yield_strain=ey))
    outs = ("xi", "sigma", "dsig_deps", "dC_dp", "iters", "flags")
    xi = torch.zeros((7, n), dtype=torch.float64, device=ctx.dev)
    for t in (20, 40):                                   # carry the state into the plastic range
        xi = mp.mp_update(mat, nw, pid, xi, synthetic.strain_at_step(d, d2, amp, t), outputs=("xi",))["xi"]
    e = synthetic.strain_at_step(d, d2, amp, 60)         # second (non-proportional) leg
    del d, d2, amp
    out = mp.allocate_outputs(mat, n, len(pid), outs, ctx.dev)
    l0 = mp.launch_count()
    ms, _, clocks = ctx.timed(lambda rec: mp.mp_update(mat, nw, pid, xi, e, outputs=outs, out=out))
    launches = (mp.launch_count() - l0) // (ctx.steps + ctx.warmup)
    key = f"hosford_a{int(a)}"
    res = {"name": key, "config": "configs[2]: Hosford plasticity with tangent + dC/dp sensitivities",
           "points_per_gpu": n, "n_gpus": ctx.world, "hosford_a": a, "newton": newton, "scaling": "weak",
           "collective": "none (points are independent)",
           "kernel": "mp_update_kernel<HOSFORD, reduced 4x4> (one thread per point; a > 8: two-pass deferral)",
           "ms_per_step": ms, "value": ctx.world * n / ms * 1e3, "unit": "updates/s",
           "launches_per_step": int(launches), "clocks": clocks,
           "plastic_fraction": float(((out["flags"] & 2) != 0).double().mean()),
           "mean_newton_iters": float(out["iters"].double().mean()),
           "max_newton_iters": int(out["iters"].max())}
    res.update(ctx.fracs(n, ms, 784, key))
    return res


# ------------------------------------------------------------------------------ configs[3]
ALG_FE = {"tet4": 96 + 96 + 8 + 56 + 1152 + 96 + 56, "hex8": 192 + 1536 + 64 + 448 + 4608 + 192 + 448}


def fe_block(ctx: Ctx, family: str, div: int):
    """One assemble_element_block equivalent (K3: R_e, K_e, xi) on a structured mesh with the
    uniaxial ramp + nodal noise of SURVEY 8d "Config 4" (J2, third load step, state carried)."""
    import torch
    from cmad_b200 import fe, fe_mesh, material_from_values, mp
    t0 = time.time()
    nodes, conn = fe_mesh.structured_hex_mesh((div,) * 3)
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    arr = fe_mesh.block_arrays(nodes, conn, device=ctx.dev)
    n_e, n_ip, n_b = arr.n_elems, arr.n_ip, arr.n_basis
    values, _, _ = _tree({"J2": 0.0}, {"E": 200e3, "nu": 0.3}, 200.0, 200.0, 20.0)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings()
    h = 1.0 / div
    xi = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=ctx.dev)
    out = {"xi": torch.empty_like(xi),
           "R_elem": torch.empty((n_e, n_b * 3), dtype=torch.float64, device=ctx.dev),
           "K_elem": torch.empty((n_e, n_b * 3, n_b * 3), dtype=torch.float64, device=ctx.dev),
           "iters": torch.empty((n_e, n_ip), dtype=torch.int32, device=ctx.dev),
           "flags": torch.empty((n_e, n_ip), dtype=torch.int32, device=ctx.dev)}
    for t in (1, 2):
        U = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, float(t), seed=42 + t, ramp=0.003,
                                                            noise=1e-3 * h)).to(ctx.dev)
        fe.fe_block_launch(mat, nw, arr, U, xi, ("xi",), {"xi": out["xi"]})
        xi, out["xi"] = out["xi"], xi
    U = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 3.0, seed=45, ramp=0.003, noise=1e-3 * h)).to(ctx.dev)
    del nodes, conn
    torch.cuda.synchronize(ctx.dev)
    setup_s = time.time() - t0
    keys = ("xi", "R_elem", "K_elem", "iters", "flags")
    l0 = mp.launch_count()
    ms, _, clocks = ctx.timed(lambda rec: fe.fe_block_launch(mat, nw, arr, U, xi, keys, out))
    launches = (mp.launch_count() - l0) // (ctx.steps + ctx.warmup)
    key = f"fe_k3_{family}"
    res = {"name": key, "config": "configs[3]: global_residuals assembly + local solves at all quadrature points",
           "family": family, "elements": n_e, "integration_points": n_e * n_ip, "n_gpus": 1,
           "kernel": f"fe_{family}_kernel (K3: R_e, K_e, xi; J2 radial return + hand-back list)",
           "ms_per_step": ms, "value": n_e / ms * 1e3, "unit": "elements/s",
           "ip_updates_per_s": n_e * n_ip / ms * 1e3, "launches_per_step": int(launches), "clocks": clocks,
           "setup_s": round(setup_s, 1),
           "plastic_fraction": float(((out["flags"] & 2) != 0).double().mean())}
    res.update(ctx.fracs(n_e, ms, ALG_FE[family], key))
    return res


# ------------------------------------------------------------------------------ configs[4]
def fe_adjoint_mixed(ctx: Ctx, family: str, div: int):
    """One adjoint-calibration step on a mixed u-p mesh (mixed_plastic.yaml-style), elements
    partitioned over the ranks in slabs (weak scaling: div^3 cells per rank).  The reference's
    deck driver forces the degree-2 volume rule on mixed decks (cmad/cli/common.py:379-391):
    tet4 x 4 points, hex8 x 8."""
    import torch
    from cmad_b200 import Parameters, active_param_ids, fe, fe_mesh, material_from_values, mp
    from cmad_b200.comm import WORLD, all_reduce_sum
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    grp = WORLD if world > 1 else None
    t0 = time.time()
    nodes, conn = fe_mesh.structured_hex_mesh((div * world, div, div), lengths=(float(world), 1.0, 1.0))
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    n_total = conn.shape[0]
    per = n_total // world
    lo, hi = rank * per, (rank + 1) * per
    arr = fe_mesh.block_arrays(nodes, conn[lo:hi], device=dev, mixed=True, volume_degree=2)
    n_nodes = nodes.shape[0]
    arr.n_dofs = n_nodes * 4
    values, active, tr = _tree({"J2": 0.0}, {"E": 200e3, "nu": 0.3}, 200.0, 200.0, 20.0)
    P = Parameters(values, active, tr)
    pid = active_param_ids(P)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings()
    h = 1.0 / div
    Uh = np.zeros(arr.n_dofs)
    Uh[:n_nodes * 3] = fe_mesh.synthetic_displacement(nodes, 2.0, seed=44, ramp=0.003, noise=1e-3 * h)
    Uh[n_nodes * 3:] = 30.0 * np.random.default_rng(6).standard_normal(n_nodes)
    U = torch.from_numpy(Uh).to(dev)
    lam = torch.from_numpy(np.random.default_rng(5).standard_normal(arr.n_dofs)).to(dev)
    del nodes, conn, Uh
    n_e, n_ip = arr.n_elems, arr.n_ip
    xi0 = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=dev)
    r_plan = fe.mixed_r_plan(arr, device=dev)
    halo = fe.InterfaceExchange(arr.elem_eq, arr.n_dofs, extra_eq=arr.elem_eq_p, group=grp)
    R = torch.empty(arr.n_dofs, dtype=torch.float64, device=dev)
    state = {}
    torch.cuda.synchronize(dev)
    setup_s = time.time() - t0

    def step(rec):
        Rm, _, xis = fe.assemble_element_block_mixed(mat, nw, arr, U, xi0, stab_mult=1.0, r_plan=r_plan)
        R.copy_(Rm)
        rec("assemble")
        halo.reduce(R)                                   # interface exchange of the residual (NCCL)
        rec("exchange_R")
        pbar, _ = fe.fe_block_vjp(mat, arr, U, xi0, xis, pid, lam, None, stab_mult=1.0, group=None)
        rec("vjp")
        all_reduce_sum(pbar, grp)                     # gradient all-reduce (NCCL)
        rec("allreduce_grad")
        state["pbar"] = pbar

    l0 = mp.launch_count()
    ms, seg, clocks = ctx.timed(step, marks=("assemble", "exchange_R", "vjp", "allreduce_grad"))
    launches = (mp.launch_count() - l0) // (ctx.steps + ctx.warmup)
    n_b = arr.n_basis
    alg = (ALG_FE[family] if family == "hex8" else 96 + 4 * (96 + 8 + 56) + 1152 + 96 + 4 * 56) \
        + 8 * (n_b + 1 + n_b + 2 * 3 * n_b * n_b + n_b * n_b)
    res = {"name": f"fe_adjoint_mixed_{family}",
           "config": "configs[4]: adjoint calibration gradient over a mixed_plastic.yaml-style mesh "
                     "with NCCL exchange + gradient all-reduce",
           "family": family, "n_ip": n_ip, "elements_per_gpu": n_e, "elements_total": n_total, "n_gpus": world,
           "scaling": "weak", "n_dofs": arr.n_dofs,
           "ms_per_step": ms, "value": n_total / ms * 1e3, "unit": "elements/s (adjoint step)",
           "ms_assemble_K3mixed_pressure_K5": seg["assemble"], "ms_exchange_R": seg["exchange_R"],
           "ms_vjp_K6": seg["vjp"], "ms_allreduce_grad": seg["allreduce_grad"],
           "collectives": {"exchange_R": {"kind": "all_reduce(sum) of the packed interface dofs, NCCL",
                                          "bytes": int(halo.n_interface) * 8, "ms": seg["exchange_R"]},
                           "allreduce_grad": {"kind": "all_reduce(sum), NCCL", "bytes": len(pid) * 8,
                                              "ms": seg["allreduce_grad"]}},
           "launches_per_step": int(launches), "clocks": clocks, "setup_s": round(setup_s, 1),
           "grad": [float(x) for x in state["pbar"].cpu()]}
    res.update(ctx.fracs(n_e, seg["assemble"], alg, f"fe_mixed_{family}"))
    key = f"fe_adjoint_mixed_{family}"
    if key in FP64_FLOPS:                                # FP64 rate of the WHOLE adjoint step (all its kernels)
        fl, src = FP64_FLOPS[key]
        res.update({"fp64_flops_per_unit_step": fl, "fp64_flops_source": src,
                    "fp64_tflops_step": n_e * fl / ms / 1e9, "frac_fp64_step": n_e * fl / ms / 1e9 / ctx.fp64})
    return res


# ------------------------------------------------------------------------------ MP objective
def mp_objective(ctx: Ctx, log2n: int = 21, nsteps: int = 20):
    """J and dJ/dp of the calibration objective over 2^21 points x 20 load steps per GPU: fused
    forward history (K1-hist) + adjoint recurrence (K2) + ONE all-reduce of 1 + P_a doubles."""
    import torch
    import torch.distributed as dist
    from cmad_b200 import NewtonSettings, Parameters, active_param_ids, material_from_values, mp, synthetic
    from cmad_b200 import _lib as L
    from cmad_b200.comm import WORLD
    from cmad_b200.objectives import BatchedMPObjective
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    lib = L.lib()
    n, N = 1 << log2n, nsteps
    values, active, tr = _tree({"J2": 0.0}, {"E": 200e3, "nu": 0.3}, 200.0, 200.0, 20.0)
    P = Parameters(values, active, tr)
    mat, pid = material_from_values(values), active_param_ids(P)
    na = len(pid)
    d, d2, a = (torch.from_numpy(x).to(dev) for x in synthetic.path_params(22, rank * n, n))
    strain = torch.zeros((N + 1, 6, n), dtype=torch.float64, device=dev)
    for k in range(N):
        strain[k + 1] = synthetic.strain_at_step(d, d2, a, round(100 * (k + 1) / N))
    del d, d2, a
    data = torch.zeros((N + 1, 9, n), dtype=torch.float64, device=dev)
    data[:, 0] = 200.0
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
    marks = {}

    def local():
        L.check(lib.cmadx_mp_forward_history(C.byref(mat), C.byref(nw), C.byref(h), stream), "fwd")
        marks["rec"]("forward_history")
        L.check(lib.cmadx_mp_objective_adjoint(C.byref(mat), pidp, na, C.byref(h), stream), "adj")
        marks["rec"]("adjoint_K2")
        return result                                  # all-reduced in place by the objective

    obj = BatchedMPObjective(P, local, group=WORLD if world > 1 else None)
    x0 = P.flat_active_values(True)
    res = {}

    def step(rec):
        marks["rec"] = rec
        res["r"] = obj.evaluate(x0)                    # + all-reduce + D2H of 1 + P_a doubles + transform_grad
        rec("allreduce_and_readback")

    l0 = mp.launch_count()
    ms, seg, clocks = ctx.timed(step, marks=("forward_history", "adjoint_K2", "allreduce_and_readback"))
    launches = (mp.launch_count() - l0) // (ctx.steps + ctx.warmup)
    ar_ms = None
    if world > 1:                                      # the collective alone, device-timed
        buf = torch.zeros(1 + na, dtype=torch.float64, device=dev)
        ar_ms, _, _ = ctx.timed(lambda rec: dist.all_reduce(buf))
    out = {"name": "mp_objective", "config": "configs[4] (material-point form): adjoint calibration objective "
                                              "J, dJ/dp with one NCCL all-reduce of (J, grad)",
           "points_per_gpu": n, "history_steps": N, "n_gpus": world, "scaling": "weak",
           "ms_per_step": ms, "value": world * n * N / ms * 1e3, "unit": "point-steps/s (objective + gradient)",
           "ms_forward_history": seg["forward_history"], "ms_adjoint_K2": seg["adjoint_K2"],
           "ms_allreduce_and_readback": seg["allreduce_and_readback"],
           "collectives": {"allreduce_J_grad": {"kind": "all_reduce(sum), NCCL", "bytes": (1 + na) * 8,
                                                "ms_alone": ar_ms}},
           "launches_per_step": int(launches), "clocks": clocks,
           "J": res["r"].J, "grad": [float(g) for g in res["r"].grad]}
    out.update(ctx.fracs(n * N, seg["forward_history"] + seg["adjoint_K2"], 108 + 176, "mp_objective"))
    # end to end through HOST buffers (cmadx_mp_objective_host): pinned strain / data histories in,
    # 8 (1 + P_a) bytes out per chunk - the calibration use case where little comes back over PCIe
    try:
        ne = 1 << 19
        sh = torch.empty((N + 1, 6, ne), dtype=torch.float64, pin_memory=True)
        dh = torch.empty((N + 1, 9, ne), dtype=torch.float64, pin_memory=True)
        sh.copy_(strain[:, :, :ne]); dh.copy_(data[:, :, :ne])
        nws = NewtonSettings(mode="imperative")
        dev_idx = dev.index if dev.index is not None else torch.cuda.current_device()
        run = lambda: mp.mp_objective_host(mat, nws, pid, sh, dh, np.ones((3, 3)), "adjoint", device=dev_idx)
        run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            r_host = run()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ms_e2e = float(dt.item()) * 1e3
        out["e2e"] = {"value": world * ne * N / ms_e2e * 1e3, "unit": "point-steps/s (objective + gradient)",
                      "api": "cmadx_mp_objective_host (pinned host histories in, (J, grad) out)",
                      "points_per_gpu": ne, "ms_per_call": ms_e2e,
                      "h2d_bytes_per_call": int(sh.numel() + dh.numel()) * 8, "d2h_bytes_per_call": (1 + na) * 8 * ((ne + (1 << 18) - 1) >> 18),
                      "pcie_gbs": (sh.numel() + dh.numel()) * 8 / ms_e2e / 1e6, "J_sample": float(r_host[0])}
        del sh, dh
    except Exception as exc:                              # the device-resident line stands on its own
        out["e2e"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    return out


def run_all(ctx: Ctx, which=None, log=None):
    """Every extra configuration, each guarded: a failure is reported in its entry, not raised."""
    import torch
    import torch.distributed as dist
    plan = [("hosford_a4", lambda: hosford(ctx, 4.0)), ("hosford_a100", lambda: hosford(ctx, 100.0))]
    if ctx.world == 1:
        plan += [("fe_k3_tet4", lambda: fe_block(ctx, "tet4", 119)), ("fe_k3_hex8", lambda: fe_block(ctx, "hex8", 216))]
    plan += [("fe_adjoint_mixed_tet4", lambda: fe_adjoint_mixed(ctx, "tet4", 80)),
             ("fe_adjoint_mixed_hex8", lambda: fe_adjoint_mixed(ctx, "hex8", 80)),
             ("mp_objective", lambda: mp_objective(ctx))]
    out = []
    for name, fn in plan:
        if which and name not in which:
            continue
        t0 = time.time()
        try:
            r = fn()
        except Exception as exc:                          # keep the headline line alive
            r = {"name": name, "error": f"{type(exc).__name__}: {exc}"[:300]}
            if ctx.world > 1:
                raise                                     # ranks must not diverge around collectives
        r["wall_s"] = round(time.time() - t0, 1)
        out.append(r)
        if log:
            log(f"[extra] {name}: {r.get('ms_per_step')} ms/step, {r.get('value')} {r.get('unit')}, wall {r['wall_s']} s")
        torch.cuda.empty_cache()
        if ctx.world > 1:
            dist.barrier()
    return out
