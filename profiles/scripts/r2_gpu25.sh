set -x
export AVIFGPU_LIBRARY=$PWD/profiles/scratch_exp/trace/libavifgpu.so
for v in "pinned 4320" "pageable 4320" "pageable 728" "pinned 728"; do python profiles/scripts/exp_pipeline_trace.py $v 2>&1 | tail -2; done
echo done
