#!/usr/bin/env python3
"""profiles/r2_other_paths.jsonl (the output of measure_generic_paths.py on a B200) -> profiles/r2_other_paths.md."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PEAK = 6567.7  # MEASURED_PEAKS.json hbm_gbs

source = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "r2_other_paths.jsonl")
rows = [json.loads(line) for line in open(source) if line.strip().startswith("{")]
out = ["# Paths outside BASELINE.json's bench lines (round 2; one B200, device-resident)\n\n",
       "`python profiles/measure_generic_paths.py` -- every launch converts a stack of 8K frames tall enough that input + output exceed the\n"
       "126 MB L2 several times over (7680 x 17280 for the 8-bit paths, 7680 x 8640 otherwise; three or four such sets rotate), so GB/s are HBM\n"
       "figures (round 1 measured single 8K frames, several of which fit in L2).  HBM peak = measured 6567.7 GB/s.  Every result is\n"
       "bit-identical to the CPU checker (tests/).  Raw lines: `r2_other_paths.jsonl`; why each kernel is where it is: `r2_other_kernels_ncu.md`.\n"
       "Launches of 50-300 us: the 3-4 us a launch needs to fill and drain the machine are inside these figures.\n\n",
       "| configuration (SURVEY section 8 row) | Gpx/s | TB/s of algorithmic traffic | of HBM peak |\n|---|---|---|---|\n"]
for r in rows:
    out.append(f'| {r["case"]} | {r["gpx_s"]:.0f} | {r["gb_s"] / 1000:.2f} | {100 * r["gb_s"] / PEAK:.0f} % |\n')
out.append("""
New in round 2: tuned kernels for premultiplied alpha on the integer hosts (was the 1-thread-per-pixel generic kernel), for float reads of
planar RGB / monochrome images (per-code EOTF table; was ~45 Gpx/s with six exact powf per pixel), for the reference's interleaved RGB
layout (was 100 Gpx/s) and for Gray(+A) float hosts (was the generic kernel, 95 Gpx/s with the exact powf); the float RGB kernels use the
compact step table, packed FP32 and a table image staged by the copy engine; the streaming kernels walk (row, column) without a division
per step and keep four 8-byte groups in flight on 8-bit images.  Tried and dropped (measured slower): a copy of the 8-bit host table per
shared-memory bank, two 16-byte groups in flight, 20 / 24 warps for the RGBA32f kernel.  Three rows (8-bit monochrome decode, Gray8
encodes) were re-measured on their own after the last changes to their kernels (`note` in the raw lines).  The float decode of a YCbCr
image (config 3 and its PQ sibling) has bench lines of its own: `r2_bench_c3_rowpair_final.json`, `r2_bench_c3pq_rowpair_final.json`.
""")
with open(os.path.join(HERE, "r2_other_paths.md"), "w") as f:
    f.write("".join(out))
print(f"{len(rows)} rows")
