#!/usr/bin/env python
"""profiles/latest_traffic.json from an `ncu --set full` report of tools/profile_kernels.py (read
here, no GPU): what bench.py prints as roofline.traffic / frac_executed.  Per element-wise kernel
the FP64-pipe and total warp instructions per warp of 32 evaluations and the DRAM bytes of the
launch; for the table build the sums over its four chained kernels.
Usage: tools/make_traffic_json.py gpurun_out/prof.ncu-rep <commit> > profiles/latest_traffic.json"""
import csv
import io
import json
import subprocess
import sys

N22, N24, NODES = 1 << 22, 1 << 24, 10000 * 1002
PROC = {"0": "bremsstrahlung", "1": "pair_production", "2": "photonuclear", "3": "ionisation"}
MASK = {"1": "bremsstrahlung", "2": "pair_production", "4": "photonuclear", "8": "ionisation"}


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * scale.get(unit, 1)


def main(path, commit):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"source": f"ncu --set full of tools/profile_kernels.py ({path})", "commit": commit,
           "note": "instr per eval = warp instructions per warp of 32 evaluations (full warps); "
                   "table_build_* = sums over the chained kernels of one config-4 build (row "
                   "parameters, terms kernels, summation); "
                   "fp64 instructions counted as sm__inst_executed_pipe_fp64.sum x 32 lanes"}
    table = {"fp64": 0.0, "inst": 0.0, "dram": 0.0, "ms": {}}
    for d in data:
        name = d[idx["Kernel Name"]]
        fp64 = float(d[idx["sm__inst_executed_pipe_fp64.sum"]])
        inst = float(d[idx["smsp__inst_executed.sum"]])
        dram = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
            to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        ms = float(d[idx["gpu__time_duration.sum"]])
        if units[idx["gpu__time_duration.sum"]] in ("us", "usecond"):
            ms /= 1e3
        elif units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond"):
            ms /= 1e6
        if "vmap_kernel<" in name:
            p = name.split("vmap_kernel<")[1].split(",")[0].strip().replace("(int)", "")
            pr = PROC[p]
            n = N24 if pr in ("bremsstrahlung", "ionisation") else N22
            out[f"{pr}_fp64_instr_per_eval_executed"] = fp64 / (n / 32)
            out[f"{pr}_instr_per_eval_executed"] = inst / (n / 32)
            out[f"{pr}_dram_bytes_per_launch"] = dram
            out[f"{pr}_fp64_pipe_pct"] = float(
                d[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]])
        elif "table_terms_kernel<" in name or "table_sum_kernel" in name or "table_rowpar" in name:
            if "table_terms_kernel<" in name:
                m = name.split("table_terms_kernel<")[1].split(">")[0].strip().replace("(int)", "")
                key = {"0": "bremsstrahlung", "1": "pair_production", "2": "photonuclear",
                       "3": "ionisation", "4": "bremsstrahlung+ionisation",
                       "6": "photonuclear+bremsstrahlung+ionisation"}.get(m, m)
            else:
                key = "summation" if "table_sum_kernel" in name else "row_parameters"
            table["fp64"] += fp64 * 32
            table["inst"] += inst * 32
            table["dram"] += dram
            table["ms"][key] = ms
            out[f"table_{key}_fp64_pipe_pct"] = float(
                d[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]])
            out[f"table_{key}_dram_bytes"] = dram
            out[f"table_{key}_fp64_instr_executed"] = fp64 * 32
        elif "table_kernel<" in name:
            m = name.split("table_kernel<")[1].split(",")[0].strip().replace("(unsigned int)", "")
            table["fp64"] += fp64 * 32
            table["inst"] += inst * 32
            table["dram"] += dram
            table["ms"][MASK.get(m, m)] = ms
            out[f"table_{MASK.get(m, m)}_fp64_pipe_pct"] = float(
                d[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]])
            out[f"table_{MASK.get(m, m)}_dram_bytes"] = dram
    out["table_build_fp64_instr_executed"] = table["fp64"]
    out["table_build_instr_executed"] = table["inst"]
    out["table_build_dram_bytes"] = table["dram"]
    out["table_build_kernel_ms_under_ncu"] = table["ms"]
    out["table_build_algorithmic_bytes"] = 8 * 10000 + 64 * 10000
    out["table_build_workspace_bytes"] = 2 * 16 * 4 * NODES     # node terms written and read back
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "unknown")
