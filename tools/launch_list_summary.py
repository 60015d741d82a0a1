"""tools/launch_list_summary.py <launches.csv> -- markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per kernel launches, total and average device time, and this library's kernels' shares of one sort."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[idx["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[idx["Kernel Name"]]
    val = float(r[idx["Metric Value"]].replace(",", "")); unit = r[idx["Metric Unit"]]
    us = val / 1000.0 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
ours = {k: v for k, v in agg.items() if re.search(r"onesweep|digit_histogram|digit_start|scan_|copy_back|partition|dist_plan|msd_|mid_sort|small_sort|exchange_|fill_kernel|copy_u32", k)}
tot = sum(v[1] for v in ours.values())
print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
for k, (c, us) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:90]}` | {c} | {us/1000:.3f} | {us/c:.1f} | {100*us/tot:.1f}% |")
print("\nOther kernels in the process (torch: data generation and verification, outside the timed region):\n")
print("| kernel | launches | total ms |\n|---|---:|---:|")
for k, (c, us) in sorted(((k, v) for k, v in agg.items() if k not in ours), key=lambda kv: -kv[1][1])[:12]:
    print(f"| `{k[:70]}` | {c} | {us/1000:.3f} |")
