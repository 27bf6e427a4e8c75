set -x
python -m pytest tests/test_gpu_fastpath.py -m gpu -q -k "gray_float" 2>&1 | tail -15
pkg=$PWD/avif-format_b200
g++ -std=c++17 -O2 -I $pkg/host $pkg/host/tools/shuttle_bench.cpp $pkg/host/GpuRowShuttle.cpp $PWD/profiles/scratch_exp/trace/libavifgpu.so -Wl,-rpath,$PWD/profiles/scratch_exp/trace -lpthread -o /tmp/shuttle_trace
nproc; lscpu | grep -i "model name\|numa\|socket" | head -8
for args in "c2 7680 4320 5 resident warm" "c2 7680 4320 5 resident fresh" "c3 7680 4320 5 resident warm"; do /tmp/shuttle_trace $args 0; done
echo done
