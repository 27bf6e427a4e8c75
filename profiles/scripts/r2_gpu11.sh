set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python profiles/measure_generic_paths.py > gpurun_out/r2_other_paths_c.jsonl 2>gpurun_out/r2_other_paths_c.err; tail -3 gpurun_out/r2_other_paths_c.err; cat gpurun_out/r2_other_paths_c.jsonl | cut -c1-200
for s in 1 0 1 0; do AVIFGPU_FLAT_SCATTER=$s python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('scatter', $s, d['value'], d['roofline']['frac'], d['roofline']['min_launch_ms'])"; done
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2_d.err | tail -1 > gpurun_out/r2_bench_c2_d.json; cut -c1-300 gpurun_out/r2_bench_c2_d.json
AVIFGPU_MEASURE_ONE_LAUNCH=1 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats --section ComputeWorkloadAnalysis --clock-control none -k regex:'EncodeRgbIntPlanar|DecodeYccToRgbInt|StreamDecode|EncodeGrayInt|TableDecode|EncodeRgbaF32|EncodeGrayF32' -c 80 -o /tmp/r2_other_kernels -f python profiles/measure_generic_paths.py > gpurun_out/ncu_other.log 2>&1
tail -3 gpurun_out/ncu_other.log
ncu -i /tmp/r2_other_kernels.ncu-rep --page raw --csv > gpurun_out/r2_other_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/r2_other_kernels.ncu-rep
echo done
