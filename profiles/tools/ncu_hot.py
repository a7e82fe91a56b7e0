"""Helper (dev container): summarise an ncu --page source --csv dump: top instructions by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("total samples", tot)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    n = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f"{100 * n / tot:5.1f}%  {r[ix['Source']][:90]:90s} {st}")
