"""tools/ncu_phase_summary.py <source-page.csv> <log2n> -- per-phase (barrier-delimited) instruction, sample and
stall-reason totals from `ncu -i X.ncu-rep --page source --csv --launch-count 1`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    if len(r) == len(hdr) and r[0] != "Address": data.append(r)
wk = (1 << int(sys.argv[2])) / 32
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
I = lambda r, c: int(r[idx[c]] or 0)
tot_s = sum(I(r, '# Samples') for r in data)
seg, cur = [], []
for k, r in enumerate(data):
    cur.append((k, r))
    src = r[idx['Source']]
    if 'BAR.SYNC' in src or 'EXIT' in src:
        seg.append(cur); cur = []
seg.append(cur)
print(f"total: {sum(I(r,'Instructions Executed') for r in data)/wk:.2f} warp-instr per 32 elements, {tot_s} samples")
for s in seg:
    ex = sum(I(r, 'Instructions Executed') for _, r in s); sm = sum(I(r, '# Samples') for _, r in s)
    if ex == 0 or not s: continue
    wf = sum(I(r, 'L1 Wavefronts Shared') for _, r in s)
    st = {}
    for _, r in s:
        for c in stall: st[c[6:]] = st.get(c[6:], 0) + I(r, c)
    top = sorted(((v, k) for k, v in st.items()), reverse=True)[:5]
    n = sum(st.values()) or 1
    print(f"idx {s[0][0]:5d}-{s[-1][0]:5d} instr/32el={ex/wk:6.2f} samples={100*sm/tot_s:5.1f}% smem_wf/32el={wf/wk:5.2f} | " + " ".join(f"{k}:{100*v/n:.0f}%" for v, k in top))
