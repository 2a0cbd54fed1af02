"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` output."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[col["# Samples"]]) for r in data)
print("kernel:", rows[0][1] if rows[0] else "?", " total samples:", tot)
agg = {n: sum(int(r[col[n]]) for r in data) for n in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:top]:
    st = {n[6:]: int(r[col[n]]) for n in stall_cols if int(r[col[n]])}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{int(r[col["# Samples"]]):7d} {100.0 * int(r[col["# Samples"]]) / max(tot, 1):5.1f}%  {r[col["Source"]].strip()[:70]:70s} {main}')
