set -x
AVIFGPU_MEASURE_ONE_LAUNCH=1 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats --section ComputeWorkloadAnalysis --clock-control none -k regex:'EncodeRgbIntPlanar|DecodeYccToRgbInt|StreamDecode|EncodeGrayInt|TableDecode|EncodeRgbaF32|EncodeGrayF32' -c 80 -o /tmp/r2_other_kernels -f python profiles/measure_generic_paths.py > gpurun_out/ncu_other.log 2>&1
tail -2 gpurun_out/ncu_other.log
ncu -i /tmp/r2_other_kernels.ncu-rep --page raw --csv > gpurun_out/r2_other_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/r2_other_kernels_raw.csv
echo done
