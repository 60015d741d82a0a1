"""tools/ncu_kernel_summary.py <report.ncu-rep> -- one JSON object per profiled launch with the metrics
the roofline section of DESIGN.md quotes (run here, on the CPU box: `ncu -i` needs no GPU)."""
import csv, io, json, subprocess, sys
WANT = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs", "smsp__inst_executed.sum": "warp_insts", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct", "lts__t_bytes.sum": "l2_bytes", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__occupancy_limit_registers": "occ_limit_regs", "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "smsp__pcsamp_warps_issue_stalled_barrier": "stall_barrier", "sm__cycles_elapsed.max": "cycles"}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void <unnamed>::", "")}
    for m, k in WANT.items():
        if m in idx:
            v = r[idx[m]].replace(",", "")
            try:
                v = float(v)
            except ValueError:
                pass
            d[k] = v
            d[k + "_unit"] = units[idx[m]]
    print(json.dumps(d))
