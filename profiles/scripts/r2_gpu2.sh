set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
AVIFGPU_LIBRARY=$PWD/profiles/scratch_exp/libavifgpu_lds32_experiment.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>&1 | tail -1 | cut -c1-900 > gpurun_out/r2_exp_lds32.json; cat gpurun_out/r2_exp_lds32.json
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2_b.err | tail -1 > gpurun_out/r2_bench_c2_b.json; cut -c1-3000 gpurun_out/r2_bench_c2_b.json; tail -5 gpurun_out/r2_bench_c2_b.err
echo done
