set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in c3 c3pq c4 c5; do python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-shuttle 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['roofline']['frac'], d['roofline']['min_launch_ms'])"; done
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-other-workloads --no-shuttle 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['value'], d['roofline']['frac'], d['roofline']['min_launch_ms'])"
python profiles/measure_generic_paths.py 2>/dev/null | grep -i "RGBA32f\|Gray32f\|GrayA32f\|PQ -> RGB32f\|HLG + OOTF" | cut -c1-200
echo done
