"""tools/ncu_src_summary.py <source-page.csv> [top] -- SASS listing with sample counts, from
`ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break  # only the first kernel block
    if len(r) == len(hdr) and r[0] != "Address":
        data.append(r)
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
mode = sys.argv[2] if len(sys.argv) > 2 else "all"
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(data), "warp-inst executed", sum(int(r[idx["Instructions Executed"]] or 0) for r in data))
for k, r in enumerate(data):
    s = int(r[idx["# Samples"]] or 0)
    ex = int(r[idx["Instructions Executed"]] or 0)
    if mode == "all" or s > tot * 0.004:
        st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        wf = r[idx["L1 Wavefronts Shared"]]; wfi = r[idx["L1 Wavefronts Shared Ideal"]]
        print(f"{k:4d} {s:6d} {100.0*s/tot:5.1f}% ex={ex:8d} {r[idx['Source']].strip()[:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]} wf={wf}/{wfi}")
