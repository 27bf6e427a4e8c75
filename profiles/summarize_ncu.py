#!/usr/bin/env python3
"""Turns an Nsight Compute report (gpurun_out/*.ncu-rep, scratch) into the small text summary committed under
profiles/: duration, DRAM bytes, pipe utilisation, stall reasons, and the hottest source lines.

    python profiles/summarize_ncu.py gpurun_out/r1_c2_fast_v8.ncu-rep profiles/r1_c2_encode_v8.ncu.txt [units_per_launch]

NCU_LAUNCH_INDEX=n in the environment picks the n-th captured launch of a report that holds several (default 0).
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def ncu(args):
    import os
    index = os.environ.get("NCU_LAUNCH_INDEX")
    pick = ["--launch-skip", index, "--launch-count", "1"] if index is not None else []
    return subprocess.run(["ncu"] + pick + args, capture_output=True, text=True).stdout


def main():
    report, out_path = sys.argv[1], sys.argv[2]
    lines = []
    raw = list(csv.reader(io.StringIO(ncu(["-i", report, "--page", "raw", "--csv"]))))
    header, units = raw[0], raw[1]
    for row in raw[2:]:
        lines.append("kernel: " + row[header.index("Kernel Name")])
        for m in METRICS:
            if m in header:
                lines.append(f"  {m:85s} {row[header.index(m)]:>18s} {units[header.index(m)]}")
        lines.append("  stall reasons (average warps stalled per issue-active cycle, > 0.1):")
        for i, h in enumerate(header):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(row[i])
                except ValueError:
                    continue
                if v > 0.1:
                    lines.append(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:6.2f}")
        break  # first captured launch
    src = list(csv.reader(io.StringIO(ncu(["-i", report, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    executed, samples = collections.Counter(), collections.Counter()
    current, hdr = None, None
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            current = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and len(r) > hdr.index("Instructions Executed"):
            try:
                n = int(r[hdr.index("Instructions Executed")])
                s = int(r[hdr.index("# Samples")] or 0)
            except ValueError:
                continue
            key = (current, int(r[0]), r[1].strip()[:110])
            executed[key] += n
            samples[key] += s
    total, total_samples = sum(executed.values()), max(sum(samples.values()), 1)
    lines.append(f"warp instructions executed (source-attributed): {total}")
    by_file = collections.Counter()
    for (f, _, _), n in executed.items():
        by_file[f] += n
    lines.append("  by file: " + ", ".join(f"{f} {100 * n / total:.1f}%" for f, n in by_file.most_common()))
    lines.append("  hottest source lines (share of executed warp instructions, share of stall samples):")
    for key, n in executed.most_common(25):
        lines.append(f"    {100 * n / total:5.1f}%  {100 * samples[key] / total_samples:5.1f}%  {key[0]}:{key[1]}  {key[2]}")
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
