#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the markdown summary kept under profiles/.
usage: ncu_summary.py REPORT.ncu-rep OUT.md "title" "command" ["reading ..."]"""
import csv, io, subprocess, sys

rep, out, title, cmd = sys.argv[1:5]
reading = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
keep = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
keep += [h for h in H if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
with open(out, "w") as f:
    f.write(f"# {title}\n\nCommand: `{cmd}` (the same command exited 0 without ncu first).\n")
    f.write(f"Source report: {rep} (scratch, not committed); read with `ncu -i ... --page raw --csv`.\n\n")
    f.write("| metric | unit | " + " | ".join(f"launch {k+1}" for k in range(len(rows) - 2)) + " |\n")
    f.write("|---|---|" + "---|" * (len(rows) - 2) + "\n")
    name = rows[2][H.index("Kernel Name")]
    for h in keep:
        if h in H:
            i = H.index(h)
            vals = [r[i] for r in rows[2:]]
            try:
                if "stalled" in h and all(float(v) < 0.02 for v in vals):
                    continue
            except ValueError:
                pass
            f.write(f"| {h} | {U[i]} | " + " | ".join(vals) + " |\n")
    f.write(f"\nKernel: `{name}`\n\n{reading}\n")
print(open(out).read())
