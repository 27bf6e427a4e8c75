set -x
python -m pytest tests/test_gpu_async_sharded.py tests/test_gpu_host_shuttle.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-other-workloads --no-cpu-baseline 2>gpurun_out/r2_bench_c2_f.err | tail -1 > gpurun_out/r2_bench_c2_f.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_f.json"))
print("c2", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k,v in d.get("e2e_shuttle",{}).get("one_gpu",{}).items(): print(k, v.get("gpx_s"), v.get("seconds_per_image"), v.get("host_seconds_per_image"))
PY
python bench.py --workload c4 --steps 10 --warmup 3 --no-other-workloads --no-cpu-baseline 2>gpurun_out/r2_bench_c4_f.err | tail -1 > gpurun_out/r2_bench_c4_f.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c4_f.json"))
print("c4", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k,v in d.get("e2e_shuttle",{}).get("one_gpu",{}).items(): print(k, v.get("gpx_s"), v.get("seconds_per_image"), v.get("host_seconds_per_image"))
PY
python profiles/measure_generic_paths.py > gpurun_out/r2_other_paths.jsonl 2>gpurun_out/r2_other_paths.err; cat gpurun_out/r2_other_paths.jsonl | cut -c1-230; tail -3 gpurun_out/r2_other_paths.err
echo done
