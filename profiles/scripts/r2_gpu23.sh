set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
pkg=$PWD/avif-format_b200
g++ -std=c++17 -O2 -I $pkg/host $pkg/host/tools/shuttle_bench.cpp $pkg/host/GpuRowShuttle.cpp $pkg/lib/libavifgpu.so -Wl,-rpath,$pkg/lib -lpthread -o /tmp/shuttle_now
for rep in 1 2; do for args in "c2 7680 4320 8 resident warm" "c2 7680 4320 8 resident fresh" "c3 7680 4320 8 resident warm" "c4 16384 16384 3 resident warm"; do /tmp/shuttle_now $args 0 | cut -c1-230; done; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'])
for k,v in d['e2e_shuttle'].items(): print(k, {kk:(vv.get('gpx_s'), vv.get('best_gpx_s')) for kk,vv in v.items()})"
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads --no-shuttle 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c4 value', d['value'], 'e2e', d['e2e']['value'])"
echo done
