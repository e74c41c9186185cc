"""Top warp-stall sample sites of one kernel from an ncu source-page CSV.
usage: ncu -i rep.ncu-rep --page source --csv [--launch-skip n --launch-count 1] > x.csv; python tools/ncu_top_stalls.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
print(rows[start - 1][1][:150] if start else "")
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[start + 1:] if len(r) == len(hdr) and r[0] != "Address"]
samp = "# Samples"
tot = sum(int(r[ci[samp]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
for r in sorted(body, key=lambda r: -int(r[ci[samp]] or 0))[:n]:
    s = int(r[ci[samp]] or 0)
    top = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}%  {r[ci['Address']][-5:]}  {r[ci['Source']][:90]:90s} {top}")
