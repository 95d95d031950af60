"""Top source lines of an `ncu -i REP --page source --csv --print-source cuda,sass` dump by stall
samples, with the dominant stall reasons.  usage: ncu_source_lines.py DUMP.csv [top]"""
import csv, sys, collections
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = {}
tot = collections.Counter()
hdr = None; cur = None; fn = None
for r in csv.reader(open(path, errors="replace")):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1][:60]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].strip().isdigit(): continue
    d = dict(zip(hdr, r))
    def iv(k):
        try: return int(d.get(k, "0") or 0)
        except ValueError: return 0
    key = (cur, int(r[0]))
    a = agg.setdefault(key, collections.Counter())
    a["smp"] += iv("# Samples"); a["ins"] += iv("Instructions Executed"); a["tins"] += iv("Thread Instructions Executed")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k: a[k] += iv(k)
    a["src"] = d.get("Source", "")[:90]
    a["lsec"] += iv("L2 Theoretical Sectors Local")
for a in agg.values():
    for k, v in a.items():
        if k != "src": tot[k] += v
print("total samples", tot["smp"], "warp instr", tot["ins"], "avg lanes", round(tot["tins"] / max(tot["ins"], 1), 2), "local sectors", tot["lsec"])
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot['smp']:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if k.startswith("stall_") and v > 0.01 * tot["smp"]))
byfile = collections.Counter()
for (f, ln), a in agg.items(): byfile[f] += a["smp"]
print("by file:", dict(byfile.most_common(6)))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_") and v), key=lambda kv: -kv[1])[:3]
    print(f"{100*a['smp']/tot['smp']:5.1f}% smp {100*a['ins']/tot['ins']:5.1f}% ins lanes {a['tins']/max(a['ins'],1):4.1f} {f}:{ln:<4d} {' '.join(f'{k}:{v}' for k, v in st):40s} | {a['src'].strip()[:70]}")
