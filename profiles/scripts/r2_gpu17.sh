set -x
for lib in "" $PWD/profiles/scratch_exp/libavifgpu_rgba20.so $PWD/profiles/scratch_exp/libavifgpu_rgba24.so "" $PWD/profiles/scratch_exp/libavifgpu_rgba20.so $PWD/profiles/scratch_exp/libavifgpu_rgba24.so; do
echo "lib=$lib"; AVIFGPU_LIBRARY=$lib AVIFGPU_MEASURE_ONLY="RGBA32f -> 12-bit PQ 4:2:0 + A, step" python profiles/measure_generic_paths.py 2>/dev/null | cut -c1-200; done
AVIFGPU_MEASURE_ONLY="10-bit" python profiles/measure_generic_paths.py 2>/dev/null | cut -c1-200
echo done
