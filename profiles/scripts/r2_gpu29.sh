set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
AVIFGPU_MEASURE_ONLY="RGBA16" python profiles/measure_generic_paths.py 2>/dev/null | cut -c1-200
python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c4 value', d['value'], d['roofline']['frac'])"
echo done
