set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python profiles/measure_generic_paths.py 2>/dev/null | grep -i "mono\|planar RGB" | cut -c1-200
for lib in "" $PWD/profiles/scratch_exp/libavifgpu_decode4.so "" $PWD/profiles/scratch_exp/libavifgpu_decode4.so; do
for w in c3 c3pq; do AVIFGPU_LIBRARY=$lib python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-shuttle 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lib=$lib', '$w', d['value'], d['roofline']['frac'], d['roofline']['min_launch_ms'])"; done; done
echo done
