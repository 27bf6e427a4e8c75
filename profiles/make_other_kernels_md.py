#!/usr/bin/env python3
"""gpurun_out/r2_other_kernels_raw.csv (ncu --page raw --csv of the capture below) -> profiles/r2_other_kernels_ncu.md."""
import csv
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
source = sys.argv[1]
r = list(csv.reader(open(source)))
h, rows = r[0], r[2:]
ci = h.index
stall = [(i, n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")]
out = ["# The other tuned kernels under ncu (round 2, kernels as committed)\n\n",
       "`AVIFGPU_MEASURE_ONE_LAUNCH=1 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats\n"
       "--section Occupancy --section SchedulerStats --section ComputeWorkloadAnalysis --clock-control none -k regex:... python profiles/measure_generic_paths.py`:\n"
       "the second launch of every case of `measure_generic_paths.py`, first occurrence of each kernel instantiation.  Single cold launches under\n"
       "the profiler: the DRAM rates are a little below the timed figures of `r2_other_paths.md`; what the table is for is the *reason* a kernel\n"
       "is where it is.  DRAM % is ncu's, of the nominal 8.19 TB/s (the measured 6.57 TB/s copy bandwidth = 80 %).\n\n",
       "| kernel | us | DRAM TB/s (% nominal) | issue slots busy | warps resident | regs | LSU wavefronts | dominant stalls (warps per issue) |\n|---|---|---|---|---|---|---|---|\n"]
seen = set()
for row in rows:
    name = re.sub(r"avifgpu::\(anonymous namespace\)::|void |unnamed>::|fastenc::", "", row[ci("Kernel Name")])
    name = re.sub(r"\(.*", "", name)
    if name in seen:
        continue
    seen.add(name)
    f = lambda n: float(row[ci(n)].replace(",", ""))  # noqa: E731
    st = sorted(((float(row[i].replace(",", "")), n.split("stalled_")[1].split("_per")[0]) for i, n in stall if row[i] not in ("", "n/a")), reverse=True)
    st = [x for x in st if x[1] != "selected"][:2]
    out.append(f'| `{name}` | {f("gpu__time_duration.sum"):.1f} | {f("dram__bytes.sum.per_second"):.2f} ({f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):.0f} %) | '
               f'{f("smsp__issue_active.avg.pct_of_peak_sustained_active"):.0f} % | {f("sm__warps_active.avg.pct_of_peak_sustained_active"):.0f} % | {row[ci("launch__registers_per_thread")]} | '
               f'{f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"):.0f} % | {", ".join(f"{b} {a:.1f}" for a, b in st)} |\n')
out.append("""
Reading:
* the integer conversion kernels on 8-bit data (`DecodeYccToRgbIntKernel<unsigned char, ...>`, `EncodeRgbIntPlanarKernel<unsigned char, ...>`) use
  70-90 % of the issue slots at 1.5-4.5 bytes per pixel: instruction-bound (a byte is as many instructions as a 16-bit sample), not
  memory-bound -- hence the packed-FP32 matrix, premultiplication and quantisers of this round;
* `EncodeRgbIntPlanarKernel<unsigned char, unsigned short, ...>` (8-bit host into a deeper image) sits on the shared-memory pipe: three
  table look-ups per pixel.  A copy of the table per bank and computing the entry instead were both measured slower;
* the 8-byte-per-thread streaming kernels (`StreamDecodeKernel<unsigned char, ...>`, `EncodeGrayIntKernel<unsigned char, ...>`) wait on
  `long_scoreboard` with a third of the issue slots used: four groups in flight per thread on 8-bit images (16-bit images lose from the
  same change: registers);
* `EncodeGrayF32Kernel<*, 1>`: the look-ups are batched ahead of the flagged samples; what is left is the look-up itself;
* `TableDecodeF32Kernel` is latency-bound with three 16-byte loads and six 16-byte stores per thread at 32-64 resident warps.
""")
with open(os.path.join(HERE, "r2_other_kernels_ncu.md"), "w") as f:
    f.write("".join(out))
print(len(seen), "kernels")
