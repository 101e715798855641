#!/usr/bin/env python
"""Summarise an ncu --set full report (read here, no GPU): per kernel the duration, registers,
occupancy, issue / FP64-pipe utilisation, executed warp instructions, DRAM traffic and the top
stall reasons.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.md"""
import csv
import io
import subprocess
import sys


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def g(d, k, default="n/a"):
        return d[idx[k]] if k in idx else default

    stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and
                  h.endswith("per_issue_active.ratio")]
    print(f"# ncu --set full summary of `{path}`\n")
    print("Per-launch values under the profiler (cold cache, serialised, ~40 replays): use the "
          "shares and utilisations, not the absolute times.\n")
    for d in data:
        name = g(d, "Kernel Name")
        fp64 = float(g(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "nan"))
        issue = float(g(d, "smsp__issue_active.avg.pct_of_peak_sustained_active", "nan"))
        inst = float(g(d, "smsp__inst_executed.sum", "nan"))
        rd = float(g(d, "dram__bytes_read.sum", "nan"))
        wr = float(g(d, "dram__bytes_write.sum", "nan"))
        ru, wu = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
        print(f"## `{name}`\n")
        print(f"| metric | value |\n|---|---|")
        print(f"| grid x block | {g(d, 'launch__grid_size')} x {g(d, 'launch__block_size')} |")
        print(f"| gpu__time_duration.sum | {g(d, 'gpu__time_duration.sum')} "
              f"{units[idx['gpu__time_duration.sum']]} |")
        print(f"| registers / thread | {g(d, 'launch__registers_per_thread')} |")
        print(f"| occupancy limit (registers, CTAs/SM) | {g(d, 'launch__occupancy_limit_registers')} |")
        print(f"| warps active (% of peak) | {g(d, 'sm__warps_active.avg.pct_of_peak_sustained_active')} |")
        print(f"| issue slots active (%) | {issue:.1f} |")
        print(f"| FP64 pipe active (% of peak, sm__inst_executed_pipe_fp64) | {fp64:.1f} |")
        print(f"| FP64 share of issued instructions | {fp64 / 2 / issue * 100:.1f} % |")
        print(f"| warp instructions executed | {inst:.4g} |")
        print(f"| eligible warps / cycle / SMSP | {g(d, 'smsp__warps_eligible.avg.per_cycle_active')} |")
        print(f"| dram__bytes_read.sum | {rd:.4g} {ru} |")
        print(f"| dram__bytes_write.sum | {wr:.4g} {wu} |")
        print(f"| sm__throughput (% of peak) | {g(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed')} |")
        vals = sorted(((float(d[idx[k]]), k) for k in stall_keys), reverse=True)[:6]
        stalls = ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                           for v, k in vals)
        print(f"| top stall reasons (warps per issue) | {stalls} |\n")


if __name__ == "__main__":
    main(sys.argv[1])
