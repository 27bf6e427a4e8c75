set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -3
python profiles/measure_generic_paths.py > gpurun_out/r2_other_paths_b.jsonl 2>gpurun_out/r2_other_paths_b.err; tail -3 gpurun_out/r2_other_paths_b.err; tail -8 gpurun_out/r2_other_paths_b.jsonl
python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench_c2_c.err | tail -1 > gpurun_out/r2_bench_c2_c.json; cut -c1-400 gpurun_out/r2_bench_c2_c.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
if [ -f profiles/scratch_exp/libavifgpu_timeline.so ]; then AVIFGPU_LIBRARY=$PWD/profiles/scratch_exp/libavifgpu_timeline.so python profiles/scripts/exp_timeline.py > gpurun_out/exp_timeline.jsonl 2>gpurun_out/exp_timeline.err; tail -3 gpurun_out/exp_timeline.err; cat gpurun_out/exp_timeline.jsonl; fi
if [ -f profiles/scripts/r2_gpu10.sh ]; then bash profiles/scripts/r2_gpu10.sh; fi
echo done
