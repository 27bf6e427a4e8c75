set -x
# evidence for the code as committed: launch list of the default bench command, ncu --set full of the config-2 / config-3 / gray kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-shuttle > gpurun_out/launch_c2.log 2>&1
tail -2 gpurun_out/launch_c2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:EncodeRgbF32Flat -c 1 -o gpurun_out/r2_c2_d -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads --no-shuttle > gpurun_out/ncu_c2_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:DecodeYccToRgbF32 -c 1 -o gpurun_out/r2_c3_b -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-shuttle > gpurun_out/ncu_c3_b.log 2>&1
ls -la gpurun_out/
# sanitizer over the kernels touched in round 2
timeout 900 compute-sanitizer --tool memcheck --kernel-regex kns=EncodeRgbF32Flat --kernel-regex kns=EncodeRgbaF32Flat --kernel-regex kns=EncodeGrayF32 --kernel-regex kns=TableDecodeF32 --kernel-regex kns=StreamDecode --kernel-regex kns=EncodeRgbIntPlanar --kernel-regex kns=EncodeGrayInt --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fastpath.py -m gpu -q -x 2>&1 | tail -6
echo memcheck rc=$?
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=EncodeRgbF32Flat --kernel-regex kns=EncodeRgbaF32Flat --kernel-regex kns=EncodeGrayF32 --error-exitcode 9 python -m pytest tests/test_gpu_fastpath.py -m gpu -q -x 2>&1 | tail -6
echo racecheck rc=$?
echo done
