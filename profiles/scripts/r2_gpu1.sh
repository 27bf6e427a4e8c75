set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_c2_a.json; cut -c1-400 gpurun_out/r2_bench_c2_a.json
python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_c3_a.json; cut -c1-300 gpurun_out/r2_bench_c3_a.json
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbF32Flat -c 1 -o gpurun_out/r2_c2_a -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2_a.log 2>&1
echo done
