set -x
export AVIFGPU_LIBRARY=$PWD/profiles/scratch_exp/pool/libavifgpu.so
python profiles/scripts/exp_pipeline_trace.py pinned 4320 2>&1 | tail -1
for nt in 0 1; do for th in 15 7 3 1; do echo "nt=$nt threads=$th"; AVIFGPU_EXP_NT=$nt AVIFGPU_EXP_THREADS=$th python profiles/scripts/exp_pipeline_trace.py pageable 4320 2>&1 | tail -1; done; done
echo done
