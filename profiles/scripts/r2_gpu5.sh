set -x
python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import sys
sys.path.insert(0, "avif-format_b200/python"); sys.path.insert(0, "tests")
import avifgpu, cases
from avifgpu import abi
with avifgpu.Context(0) as gpu:
    desc = abi.EncodeDesc(64, 64, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_420, abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, cases.NCLX_2020_PQ())
    print(gpu.prepare_encode(desc).as_dict())
PY
for w in 0 1; do
AVIFGPU_WIDE_TABLE_ENTRIES=$w python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('wide=$w', d['value'], d['roofline']['frac'], d['roofline']['mean_launch_ms'])"
done
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbF32Flat -c 1 -o gpurun_out/r2_c2_b -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads --no-shuttle > gpurun_out/ncu_c2_b.log 2>&1
python bench.py --steps 30 --warmup 5 --no-other-workloads --no-cpu-baseline 2>gpurun_out/r2_bench_c2_d.err | tail -1 > gpurun_out/r2_bench_c2_d.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_c2_d.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"])
for k,v in d.get("e2e_shuttle",{}).get("one_gpu",{}).items(): print(k, v)
PY
ncu --set full --clock-control none --import-source on -k regex:DecodeYccToRgbF32 -c 1 -o gpurun_out/r2_c3_a -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads --no-shuttle > gpurun_out/ncu_c3_a.log 2>&1
echo done
