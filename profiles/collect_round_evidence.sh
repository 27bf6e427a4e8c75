#!/bin/bash
# The commands behind the *_final* files of this directory (run from the repo root on a B200 box; outputs land in
# gpurun_out/ and are then summarised / copied here, see README.md).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c3 c4 c5; do python bench.py --workload $w --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_${w}_final.json; cut -c1-120 gpurun_out/bench_${w}_final.json; done
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c2_reference.json; cut -c1-200 gpurun_out/bench_c2_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launch_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbF32Flat -c 1 -o gpurun_out/r1_c2_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:DecodeYccToRgbF32 -c 1 -o gpurun_out/r1_c3_final -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbIntPlanar -c 1 -o gpurun_out/r1_c4_final -f python bench.py --workload c4 --steps 1 --warmup 1 > gpurun_out/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:EncodeGray16Lut -c 1 -o gpurun_out/r1_c5_final -f python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
echo done
