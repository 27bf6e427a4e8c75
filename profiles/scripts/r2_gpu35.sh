set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2.err | tail -1 > gpurun_out/r2_bench_c2_full_line.json; tail -2 gpurun_out/r2_bench_c2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_full_line.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k,v in d["e2e_shuttle"].items(): print(k, {kk:(vv.get("gpx_s"), vv.get("best_gpx_s")) for kk,vv in v.items()})
PY
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c4 value', d['value'], 'e2e', d['e2e']['value'], {k:{kk:vv.get('gpx_s') for kk,vv in v.items()} for k,v in d['e2e_shuttle'].items()})"
echo done
