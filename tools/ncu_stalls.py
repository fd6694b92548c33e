#!/usr/bin/env python3
"""Print stall-reason ratios and pipe utilisation from an ncu report (raw page)."""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
d = dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1]))
st = sorted(((float(v), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for k, v in d.items()
             if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k and v not in ("", "n/a")), reverse=True)
print("stall cycles per issued instruction:", ", ".join(f"{n} {v:.2f}" for v, n in st[:9]))
keys = ["sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_lsu",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_uniform", "sm__inst_executed_pipe_cbu", "sm__inst_executed_pipe_adu"]
print("pipe % of peak:", ", ".join(f"{k.split('pipe_')[1]} {float(d[k + '.avg.pct_of_peak_sustained_active']):.1f}" for k in keys
                                 if k + ".avg.pct_of_peak_sustained_active" in d))
for k in ["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__time_duration.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]:
    if k in d:
        print(f"{k}: {d[k]}")
