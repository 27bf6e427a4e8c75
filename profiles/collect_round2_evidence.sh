#!/bin/bash
# The commands behind the r2_* files of this directory (run from the repo root on a B200 box; outputs land in gpurun_out/ and are
# then summarised / copied here, see README.md).  One GPU unless the file name says otherwise.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2.err | tail -1 > gpurun_out/r2_bench_c2_full_line.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_bench_c2_reference_arm.json
python profiles/measure_generic_paths.py > gpurun_out/r2_other_paths.jsonl 2>/dev/null
# launch list and full captures (see also scripts/r2_gpu15.sh for the sanitizer runs)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-shuttle > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbF32Flat -c 1 -o gpurun_out/r2_c2_d -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads --no-shuttle > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:DecodeYccToRgbF32 -c 1 -o gpurun_out/r2_c3_b -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-shuttle > /dev/null 2>&1
# multi-GPU lines (gpurun --gpus N): for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 | tail -1 > gpurun_out/r2_bench_n$n.json; done
echo done
