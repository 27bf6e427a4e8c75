set -x
python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_host_shuttle.py tests/test_icc_matrix.py -m gpu -x -q 2>&1 | tail -5
for w in 0 1; do
AVIFGPU_WIDE_TABLE_ENTRIES=$w python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('wide=$w', d['value'], d['roofline']['frac'], d['roofline']['mean_launch_ms'])"
done
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbF32Flat -c 1 -o gpurun_out/r2_c2_c -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads --no-shuttle > gpurun_out/ncu_c2_c.log 2>&1
echo done
