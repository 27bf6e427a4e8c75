set -x
cat /sys/kernel/mm/transparent_hugepage/enabled
python -m pytest tests/test_gpu_host_shuttle.py -m gpu -q 2>&1 | tail -3
pkg=$PWD/avif-format_b200
g++ -std=c++17 -O2 -I $pkg/host $pkg/host/tools/shuttle_bench.cpp $pkg/host/GpuRowShuttle.cpp $pkg/lib/libavifgpu.so -Wl,-rpath,$pkg/lib -lpthread -o /tmp/shuttle_now
for rep in 1 2; do for args in "c2 7680 4320 8 resident fresh" "c2 7680 4320 8 resident warm" "c4 16384 16384 3 resident fresh" "c4 16384 16384 3 resident warm"; do echo "$args"; /tmp/shuttle_now $args 0 | cut -c100-230; done; done
echo done
