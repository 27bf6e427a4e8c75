set -x
# one ncu --set full capture per "other path" kernel: every second launch of each case (the first is the warm-up)
AVIFGPU_MEASURE_ONE_LAUNCH=1 ncu --set full --clock-control none -k regex:'EncodeRgbIntPlanar|DecodeYccToRgbInt|StreamDecode|EncodeGrayInt|TableDecode|EncodeRgbaF32|EncodeGrayF32|EncodeRgbF32Flat|EncodeRgbF32Clip' -c 120 -o gpurun_out/r2_other_kernels -f python profiles/measure_generic_paths.py > gpurun_out/ncu_other.log 2>&1
tail -5 gpurun_out/ncu_other.log
ls -la gpurun_out/
echo done
