"""Development aid: top stall sites of an `ncu --page source --csv` export (per SASS instruction: samples and dominant stall reasons)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[col["# Samples"]])
    except ValueError: continue
    tot += n
    st = sorted(((int(r[col[c]] or 0), c) for c in stall_cols), reverse=True)[:3]
    data.append((n, r[col["Address"]], r[col["Source"]], r[col["Instructions Executed"]], st))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", tot)
# totals per stall reason
agg = {c: 0 for c in stall_cols}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    for c in stall_cols:
        try: agg[c] += int(r[col[c]] or 0)
        except ValueError: pass
print("by reason:", ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / max(tot, 1)) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, addr, src, ex, st in sorted(data, reverse=True)[:topn]:
    print("%6d %5.1f%%  %s  exec=%s  %-60s %s" % (n, 100.0 * n / max(tot, 1), addr[-5:], ex, src[:60], " ".join("%s:%d" % (c[6:], v) for v, c in st if v)))
