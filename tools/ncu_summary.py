"""One-page text summary of an .ncu-rep (first N kernels): durations, DRAM bytes, pipe utilisation, occupancy, stall samples.
usage: python tools/ncu_summary.py <rep> [max_kernels]"""
import csv, subprocess, sys
rep = sys.argv[1]; limit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for r in rows[2:2 + limit]:
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"{w},{units[i]},{r[i]}")
    tot = sum(float(r[hdr.index(s)] or 0) for s in stalls) or 1.0
    top = sorted(((float(r[hdr.index(s)] or 0), s) for s in stalls), reverse=True)[:8]
    for v, s in top:
        print(f"{s},pct_of_samples,{100 * v / tot:.1f}")
    print()
