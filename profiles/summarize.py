#!/usr/bin/env python
"""Turns gpurun_out/prof_eval_<tag>.ncu-rep (+ launch list, bench lines) into the tracked
summaries under profiles/:  python profiles/summarize.py <tag> [round-prefix]"""
import csv
import io
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else "r01"
kind = sys.argv[3] if len(sys.argv) > 3 else "eval"      # "eval" | "grad": which kernel's capture
rep = f"gpurun_out/prof_{kind}_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second",
]
out = {"source": rep, "kernel": d.get("Kernel Name", ("", ""))[1]}
out.update({k: {"unit": d[k][0], "value": d[k][1]} for k in keys if k in d})
keys_extra = ["l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
              "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]
out.update({k: {"unit": d[k][0], "value": d[k][1]} for k in keys_extra if k in d})
json.dump(out, open(f"profiles/{pre}_{kind}_kernel_ncu_metrics.json", "w"), indent=1)
if kind != "eval":
    for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"):
        print(k, d[k])
    sys.exit(0)
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
rd = float(d["dram__bytes_read.sum"][1]) * scale[d["dram__bytes_read.sum"][0]]
wr = float(d["dram__bytes_write.sum"][1]) * scale[d["dram__bytes_write.sum"][0]]
json.dump({"kernel": "dex::eval_kernel<float, 2, true, false, false, 256, true>",
           "source": f"ncu --set full, {rep} (one launch of bench.py configs[1])",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr},
          open("profiles/eval_kernel_traffic.json", "w"), indent=1)
for src, dst in ((f"gpurun_out/launches_{tag}.csv", f"profiles/{pre}_launches.csv"),
                 (f"gpurun_out/bench_{tag}.json", f"profiles/{pre}_bench_n1.json"),
                 (f"gpurun_out/bench_ref_{tag}.json", f"profiles/{pre}_bench_reference_arm.json"),
                 (f"gpurun_out/configs_{tag}.jsonl", f"profiles/{pre}_configs.jsonl")):
    try:
        shutil.copy(src, dst)
    except FileNotFoundError:
        print("missing", src)
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"):
    print(k, d[k])
print("dram bytes", rd + wr)
