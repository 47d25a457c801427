"""Developer tool: summarise an `ncu --page source --csv` export (stall reasons, hottest SASS)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break  # next kernel instance
    if len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print(f"{s:25s} {v:8d} {100 * v / max(tot, 1):.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    st = {s: int(r[idx[s]] or 0) for s in stalls}
    main = max(st, key=st.get)
    print(r[idx["# Samples"]].rjust(7), r[idx["Instructions Executed"]].rjust(10), main.ljust(18), r[idx["Source"]][:100])
