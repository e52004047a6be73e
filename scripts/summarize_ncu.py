#!/usr/bin/env python3
"""Reads an .ncu-rep (ncu --set full --import-source on) here on the CPU box and writes
  profiles/<tag>_ncu_full_summary.md   key metrics per profiled kernel
  /tmp/<tag>_<kernel>.sass.txt         per-instruction executed counts / active lanes / stall samples
Usage: scripts/summarize_ncu.py gpurun_out/X.ncu-rep <tag> [symbol_steps_per_launch]"""
import csv, io, json, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
steps = float(sys.argv[3]) if len(sys.argv) > 3 else 65536 / 32 * 65537     # warp-steps of the headline batch
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size",
        "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
out = ["# %s -- `ncu --set full --clock-control none`, B200\n" % tag,
       "Times under ncu are not bench values; one capture per kernel (the timed launches of `bench.py --steps 1 --warmup 3`).\n"]
traffic = {}
for r in data:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    out.append("\n## %s\n\n| metric | unit | value |\n|---|---|---|" % name)
    for k in KEYS:
        if k in d:
            out.append("| %s | %s | %s |" % (k, units[hdr.index(k)], d[k]))
    def f(k):
        return float(d[k].replace(",", ""))
    def scaled(k):
        u = units[hdr.index(k)]
        mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[u]
        return f(k) * mul
    inst = f("smsp__inst_executed.sum")
    out.append("| warp-instructions per symbol step (32 streams) | inst | %.1f |" % (inst / steps))
    short = name.split("<")[0].replace("void ", "").replace("rdx::", "").strip()
    traffic[short] = {"dram_bytes_read": scaled("dram__bytes_read.sum"), "dram_bytes_write": scaled("dram__bytes_write.sum"),
                      "traffic": scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum"),
                      "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "warp_inst_executed": inst, "warp_inst_per_symbol_step": round(inst / steps, 1),
                      "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                      "avg_active_threads_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
                      "kernel_name": name}
open("profiles/%s_ncu_full_summary.md" % tag, "w").write("\n".join(out) + "\n")
json.dump(traffic, open("/tmp/%s_traffic.json" % tag, "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, cur = None, None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        kern = r[1].split("<")[0].replace("void ", "").replace("rdx::", "").strip()
        cur = open("/tmp/%s_%s.sass.txt" % (tag, kern), "w")
        i = 0
        continue
    if r and r[0] == "Address":
        continue
    if cur and len(r) > 8:
        cur.write("%5d %12d %5.1f %7d  %s\n" % (i, int(r[5]), float(r[8]), int(r[4]), r[1]))
        i += 1
print("\n".join(out[:3]))
for k, v in traffic.items():
    print(k, "inst/step", v["warp_inst_per_symbol_step"], "issue%", v["issue_active_pct"], "lanes", v["avg_active_threads_per_inst"], "traffic GB", v["traffic"] / 1e9)
