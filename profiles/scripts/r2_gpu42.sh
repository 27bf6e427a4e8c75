# Final check of the float decode kernel at two CTAs per SM: the decode-side parity tests (incl. both whole-domain sweeps),
# c3 / c3pq bench lines, then one ncu --set full capture of the HLG and the PQ launch.
set -x
timeout 150 python -m pytest tests -m gpu -x -q -k "dec_ycc32 or ycc_to_rgb32 or hlg_ootf_exponents or production_decode or production_pq or config3" 2>&1 | tail -6
B="timeout 60 python bench.py --no-other-workloads --no-shuttle --no-cpu-baseline --steps 30 --warmup 5"
for wl in c3 c3pq; do $B --workload $wl 2>gpurun_out/r2_42_$wl.err | tee gpurun_out/r2_42_${wl}.json | cut -c1-330; done
timeout 90 ncu --set full --clock-control none --import-source on -k regex:DecodeYccToRgbF32 --launch-skip 1 --launch-count 3 -o gpurun_out/r2_c3_c -f python profiles/scripts/decode_two_launches.py 2>&1 | tail -3
ls -la gpurun_out/r2_c3_c.ncu-rep
echo done
