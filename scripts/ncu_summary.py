"""Summarise an ncu report (read here, no GPU needed) into profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_refl_toa  [kernel-key]

Writes <out>.summary.json (selected metrics per captured launch) and, when kernel-key is
given, updates profiles/traffic.json with the mean DRAM bytes per launch under
"<kernel-key>_dram_bytes_per_launch" (bench.py reads it for roofline.traffic).
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    v = r[i]
                d[k] = {"value": v, "unit": units[i]}
        launches.append(d)
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out + ".summary.json", "w") as f:
        json.dump({"report": os.path.basename(rep), "launches": launches}, f, indent=1)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = []
    for d in launches:
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in d:
                b += d[k]["value"] * scale.get(d[k]["unit"], 1)
        tot.append(b)
        print(d["kernel"][:60], "time", d.get("gpu__time_duration.sum"), "dram bytes %.4g" % b,
              "fp64 pipe %", d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", {}).get("value"),
              "regs", d.get("launch__registers_per_thread", {}).get("value"))
    if key and tot:
        tp = os.path.join(os.path.dirname(out) or ".", "traffic.json")
        cur = json.load(open(tp)) if os.path.isfile(tp) else {}
        cur[key + "_dram_bytes_per_launch"] = sum(tot) / len(tot)
        cur[key + "_source"] = os.path.basename(out) + ".summary.json"
        json.dump(cur, open(tp, "w"), indent=1)


if __name__ == "__main__":
    main()
