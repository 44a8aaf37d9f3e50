#!/usr/bin/env python
"""Digest of an Nsight Compute report: the numbers this repository argues with, as JSON.

    python scripts/ncu_digest.py gpurun_out/prof_stiffness_r1n.ncu-rep [-k stiffness_line] [--stalls]

Runs `ncu -i <report> --page raw --csv` (works without a GPU) and keeps, per profiled launch:
duration, DRAM bytes and throughput, L1TEX data-pipe wavefronts (shared / global split), L2 and SM
throughput, pipe utilisation (FP64, LSU, issue slots), occupancy, registers, and -- with --stalls --
the warp stall reasons sorted by weight.  The summaries under profiles/*_ncu_full_*.json are this
script's output.
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "local_load_bytes", "local_store_bytes",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]
STALL = re.compile(r"smsp__average_warps?_?(?:latency_)?issue_stalled_(\w+?)(?:_per_warp_active)?\.(?:pct|ratio)$"
                   r"|smsp__pcsamp_warps_issue_stalled_(\w+)$")


def load(report):
    res = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True)
    if res.returncode != 0:
        raise SystemExit(f"ncu -i {report} failed:\n{res.stderr[-2000:]}")
    text = res.stdout[res.stdout.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(text)))
    return rows[0], rows[1], rows[2:]


def short(name):
    """Column names carry a section prefix (e.g. SM_A.TriageCompute.): keep the metric itself."""
    m = re.search(r"((?:gpu|dram|l1tex|lts|sm|smsp|launch|gr|gpc|fbpa|local|derived)__?[\w.]*)$", name)
    return m.group(1) if m else name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("-k", "--kernel", default="", help="regex on the kernel name")
    ap.add_argument("--stalls", action="store_true", help="add the warp stall reasons, heaviest first")
    ap.add_argument("--all", action="store_true", help="every metric of the report (large)")
    args = ap.parse_args()
    out = []
    for rep in args.reports:
        head, units, rows = load(rep)
        names = [short(h) for h in head]
        kcol = head.index("Kernel Name")
        for r in rows:
            if args.kernel and not re.search(args.kernel, r[kcol]):
                continue
            rec = {"report": rep, "kernel": r[kcol]}
            stalls = {}
            for n, u, v in zip(names, units, r):
                if v == "":
                    continue
                if args.all or n in KEEP:
                    rec.setdefault(n, f"{v} {u}".strip())
                m = STALL.search(n)
                if args.stalls and m:
                    try:
                        stalls[n] = float(v.replace(",", ""))
                    except ValueError:
                        pass
            if stalls:
                rec["stalls_heaviest_first"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:14])
            out.append(rec)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
