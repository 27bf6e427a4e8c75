set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2_e.err | tail -1 > gpurun_out/r2_bench_c2_e.json; tail -3 gpurun_out/r2_bench_c2_e.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_e.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k,v in d["e2e_shuttle"].items():
    print(k, {kk:(vv.get("gpx_s"), vv.get("best_gpx_s")) if isinstance(vv,dict) else vv for kk,vv in v.items()})
print({k:(v["value"], v.get("roofline_frac")) for k,v in d["other_workloads"].items()})
PY
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_bench_c2_reference_arm.json; cut -c1-300 gpurun_out/r2_bench_c2_reference_arm.json
echo done
