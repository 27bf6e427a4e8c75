set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2_a.err | tail -1 > gpurun_out/r2_bench_c2_a.json; cut -c1-1500 gpurun_out/r2_bench_c2_a.json; tail -5 gpurun_out/r2_bench_c2_a.err
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbF32Flat -c 1 -o gpurun_out/r2_c2_a -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads --no-shuttle > gpurun_out/ncu_c2_a.log 2>&1
echo done
