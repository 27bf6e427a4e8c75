set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in c5 c3 c3pq; do python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-shuttle 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['roofline']['frac'], d['roofline']['min_launch_ms'])"; done
python profiles/measure_generic_paths.py 2>/dev/null | grep -i "Gray32f\|GrayA32f\|mono" | cut -c1-200
echo done
