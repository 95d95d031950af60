"""FP64 flops per kernel launch from an `ncu --set full` report: 2 x DFMA + DMUL + DADD
predicated-on thread instructions (per_cycle_elapsed sums x elapsed SMSP cycles).
usage: python profiles/tools/ncu_flops.py REP [units_per_launch]"""
import csv, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
col = lambda r, k: float(r[h.index(k)].replace(",", ""))
for r in rows[2:]:
    cyc = col(r, "smsp__cycles_elapsed.avg")
    f = {k: col(r, f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc for k in ("dfma", "dmul", "dadd")}
    flops = 2 * f["dfma"] + f["dmul"] + f["dadd"]
    t = col(r, "gpu__time_duration.sum")
    unit_t = rows[1][h.index("gpu__time_duration.sum")]
    secs = t * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(unit_t, 1e-6)
    print(r[h.index("Kernel Name")][:90])
    print(f"  dfma {f['dfma']:.4g} dmul {f['dmul']:.4g} dadd {f['dadd']:.4g} thread-inst -> {flops:.4g} FP64 flops, "
          f"{flops / secs / 1e12:.2f} TFLOP/s over {t} {unit_t}" + (f", {flops / units:.1f} flops/unit" if units else ""))
