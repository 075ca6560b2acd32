"""Key counters of one kernel from an .ncu-rep as JSON (what profiles/ncu_*_summary.json hold).
   python tools/ncu_summary.py report.ncu-rep kernel_regex > profiles/ncu_<round>_<kernel>_summary.json"""
import csv
import json
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, v = rows[0], rows[1], rows[-1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
]
res = {"kernel": v[h.index("Kernel Name")] if "Kernel Name" in h else kern, "report": rep}
for k in KEYS:
    if k in h:
        i = h.index(k)
        res[k] = {"value": float(v[i].replace(",", "")), "unit": units[i]}
stalls = {}
for i, n in enumerate(h):
    if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
        x = float(v[i].replace(",", ""))
        if x >= 0.3:
            stalls[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = x
res["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
print(json.dumps(res, indent=1))
