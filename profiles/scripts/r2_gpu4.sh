set -x
for v in lds32_experiment lds64_control; do
AVIFGPU_LIBRARY=$PWD/profiles/scratch_exp/libavifgpu_$v.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['roofline']['mean_launch_ms'])"
done
python bench.py --steps 30 --warmup 5 --no-other-workloads --no-cpu-baseline 2>gpurun_out/r2_bench_c2_d.err | tail -1 > gpurun_out/r2_bench_c2_d.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_d.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"])
for k,v in d.get("e2e_shuttle",{}).get("one_gpu",{}).items(): print(k, v)
PY
ncu --set full --clock-control none --import-source on -k regex:DecodeYccToRgbF32 -c 1 -o gpurun_out/r2_c3_a -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads --no-shuttle > gpurun_out/ncu_c3_a.log 2>&1
echo done
