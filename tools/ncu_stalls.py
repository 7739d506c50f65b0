"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        n = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    st = {}
    for s in stalls:
        v = int(r[col[s]] or 0)
        tot[s] += v
        if v:
            st[s] = v
    data.append((n, r[col["Source"]].strip(), st, r[col["Instructions Executed"]]))
T = sum(tot.values()) or 1
print("total stall samples", T, " instructions", len(data))
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
    print("  %-24s %8d %5.1f%%" % (s, v, 100 * v / T))
print("--- hottest instructions (samples, executed, SASS, top stalls)")
for n, src, st, ie in sorted(data, key=lambda x: -x[0])[:top]:
    print("%6d %9s  %-64s %s" % (n, ie, src[:64], dict(sorted(st.items(), key=lambda x: -x[1])[:3])))
