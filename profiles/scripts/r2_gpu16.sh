set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python profiles/measure_generic_paths.py > gpurun_out/r2_other_paths_d.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r2_other_paths_d.jsonl
if [ -f profiles/scratch_exp/libavifgpu_timeline.so ]; then AVIFGPU_LIBRARY=$PWD/profiles/scratch_exp/libavifgpu_timeline.so python profiles/scripts/exp_timeline.py > gpurun_out/exp_timeline_b.jsonl 2>gpurun_out/exp_timeline_b.err; tail -3 gpurun_out/exp_timeline_b.err; cut -c1-900 gpurun_out/exp_timeline_b.jsonl; fi
echo done
