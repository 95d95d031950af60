"""Reduce the ncu CSVs written by benchmarks/count_flops.sh (one timed step of an extra config,
NVTX-filtered) to FP64 flops per unit: sum over the region's kernels of 2 x DFMA + DMUL + DADD
predicated-on thread instructions, divided by the units one step processes (read from the bench
JSON line of the same run).  usage: ncu_region_flops.py DIR  -> writes profiles/r2_fp64_flops.json"""
import csv, json, os, sys
d = sys.argv[1]
UNITS = {"hosford_a4": "points_per_gpu", "hosford_a100": "points_per_gpu", "fe_k3_tet4": "elements", "fe_k3_hex8": "elements",
         "fe_adjoint_mixed_tet4": "elements_per_gpu", "fe_adjoint_mixed_hex8": "elements_per_gpu", "mp_objective": None}
out = {}
for name, ukey in UNITS.items():
    p = os.path.join(d, f"r2_flops_{name}.csv")
    if not os.path.exists(p):
        continue
    rows = [r for r in csv.reader(open(p, errors="replace")) if len(r) > 3]
    try:
        h = next(r for r in rows if "Metric Name" in r)
    except StopIteration:
        continue
    iN, iV, iK = h.index("Metric Name"), h.index("Metric Value"), h.index("Kernel Name")
    tot = {"dfma": 0.0, "dmul": 0.0, "dadd": 0.0}
    kernels = {}
    for r in rows:
        if r is h or len(r) <= iV:
            continue
        for k in tot:
            if f"op_{k}_pred_on" in r[iN]:
                v = float(r[iV].replace(",", ""))
                tot[k] += v
                kernels[r[iK][:60]] = kernels.get(r[iK][:60], 0.0) + v * (2 if k == "dfma" else 1)
    line = json.loads(open(os.path.join(d, f"r2_flops_{name}.json")).read().strip().splitlines()[-1])
    cfg = next(c for c in line["extra"]["configs"] if c["name"] == name)
    units = cfg[ukey] if ukey else cfg["points_per_gpu"] * cfg["history_steps"]
    flops = 2 * tot["dfma"] + tot["dmul"] + tot["dadd"]
    out[name] = {"flops_per_unit": flops / units, "units": units, "flops": flops,
                 "source": f"profiles/r2_flops_{name}.csv (ncu, timed region of one bench step)",
                 "by_kernel": {k: v / units for k, v in sorted(kernels.items(), key=lambda kv: -kv[1])[:6]}}
    print(name, round(flops / units, 1), "flops/unit")
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "r2_fp64_flops.json"), "w"), indent=1)
