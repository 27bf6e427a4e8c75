set -x
AVIFGPU_LIBRARY=$PWD/profiles/scratch_exp/libavifgpu_lds32_experiment.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>&1 | tail -1 | cut -c1-1100 > gpurun_out/r2_exp_lds32.json; cat gpurun_out/r2_exp_lds32.json
python -m pytest tests/test_gpu_async_sharded.py tests/test_gpu_host_shuttle.py -m gpu -q 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 --no-other-workloads 2>gpurun_out/r2_bench_c2_c.err | tail -1 > gpurun_out/r2_bench_c2_c.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_c.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"])
print(json.dumps(d.get("e2e_shuttle"), indent=1))
PY
echo done
