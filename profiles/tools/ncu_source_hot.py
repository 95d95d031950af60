"""Aggregate `ncu --page source --csv --print-source cuda,sass` by source file:line for one
kernel: stall samples and executed instructions per line, top N."""
import csv, sys, collections
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# the dump is a sequence of sections: ("File Path", f), ("Function Name", fn), header, rows...
agg = collections.OrderedDict()
cur_file, hdr = None, None
tot_s = tot_i = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or cur_file is None:
        continue
    d = dict(zip(hdr, r))
    try:
        s = int(d.get("# Samples", "0") or 0); i = int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        continue
    key = (cur_file, d["Line No"])
    a = agg.setdefault(key, [0, 0, d.get("Source", "")[:110]])
    a[0] += s; a[1] += i
    tot_s += s; tot_i += i
print(f"total samples {tot_s}  instructions {tot_i}")
for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% ins  {f}:{ln:>4s}  {src.strip()}")
