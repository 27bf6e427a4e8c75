set -x
python profiles/scripts/exp_pipeline.py 2>&1 | tail -12
numactl --hardware 2>/dev/null | head -5
echo done
