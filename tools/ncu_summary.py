"""Summarise an .ncu-rep: python tools/ncu_summary.py file.ncu-rep [out.md]   (reads via `ncu -i ... --page raw --csv`)"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
lines = []
names = [r[h.index("Kernel Name")][:70] for r in rows[2:]]
lines.append("| metric | " + " | ".join(f"launch {i+1}" for i in range(len(rows) - 2)) + " | unit |")
lines.append("|---|" + "---:|" * (len(rows) - 2) + "---|")
for i, n in enumerate(h):
    if n in want:
        lines.append(f"| {n} | " + " | ".join(r[i] for r in rows[2:]) + f" | {u[i]} |")
st = [(i, n) for i, n in enumerate(h) if "pcsamp_warps_issue_stalled" in n and "not_issued" not in n]
st.sort(key=lambda t: -float(rows[2][t[0]] or 0))
for i, n in st[:8]:
    lines.append(f"| {n} | " + " | ".join(r[i] for r in rows[2:]) + f" | {u[i]} |")
txt = "kernels: " + "; ".join(names) + "\n\n" + "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "a").write(txt)
print(txt)
