set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2_e.err | tail -1 > gpurun_out/r2_bench_c2_e.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_e.json"))
print("c2", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
for k,v in d.get("e2e_shuttle",{}).get("one_gpu",{}).items(): print(k, v.get("gpx_s"), v.get("seconds_per_image"), v.get("host_seconds_per_image"))
for k,v in d.get("other_workloads",{}).items(): print(k, v["value"], v["roofline_frac"])
PY
python profiles/measure_generic_paths.py > gpurun_out/r2_other_paths.log 2>&1; tail -30 gpurun_out/r2_other_paths.log
echo done
