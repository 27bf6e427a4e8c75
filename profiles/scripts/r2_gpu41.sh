# New float decode kernel (branch-free powf, exponent-folded log2 table, row-pair units, plane walks): parity first, then
# c3 / c3pq at 3 and 2 CTAs per SM.
set -x
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
B="timeout 120 python bench.py --no-other-workloads --no-shuttle --no-cpu-baseline --steps 30 --warmup 5"
for wl in c3 c3pq; do $B --workload $wl 2>gpurun_out/r2_41_$wl.err | tee gpurun_out/r2_41_${wl}_b3.json | cut -c1-330; done
cp avif-format_b200/build/variants/libavifgpu_b2.so avif-format_b200/lib/libavifgpu.so
for wl in c3 c3pq; do $B --workload $wl 2>>gpurun_out/r2_41_$wl.err | tee gpurun_out/r2_41_${wl}_b2.json | cut -c1-330; done
tail -3 gpurun_out/r2_41_c3.err gpurun_out/r2_41_c3pq.err
echo done
