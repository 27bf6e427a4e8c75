set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2.err | tail -1 > gpurun_out/r2_bench_c2_full_line.json; tail -2 gpurun_out/r2_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_bench_c2_reference_arm.json
python profiles/measure_generic_paths.py > gpurun_out/r2_other_paths.jsonl 2>/dev/null; wc -l gpurun_out/r2_other_paths.jsonl
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_full_line.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
for k,v in d["e2e_shuttle"].items(): print(k, {kk:(vv.get("gpx_s"), vv.get("best_gpx_s")) for kk,vv in v.items()})
print({k:(v["value"], v.get("roofline_frac")) for k,v in d["other_workloads"].items()})
PY
timeout 600 compute-sanitizer --tool memcheck --kernel-regex kns=EncodeRgbIntPlanar --kernel-regex kns=DecodeYccToRgbInt --kernel-regex kns=StreamDecode --kernel-regex kns=EncodeGrayF32 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fastpath.py -m gpu -q -x -k "not every_" 2>&1 | tail -4
echo memcheck rc=$?
echo done
