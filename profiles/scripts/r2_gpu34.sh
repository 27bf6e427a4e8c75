set -x
pkg=$PWD/avif-format_b200
g++ -std=c++17 -O2 -I $pkg/host $pkg/host/tools/shuttle_bench.cpp $pkg/host/GpuRowShuttle.cpp $PWD/profiles/scratch_exp/pool/libavifgpu.so -Wl,-rpath,$PWD/profiles/scratch_exp/pool -lpthread -o /tmp/shuttle_pool
for th in 15 7 5 3 2; do for args in "c2 7680 4320 8 resident warm" "c2 7680 4320 8 resident fresh" "c3 7680 4320 8 resident warm" "c4 16384 16384 3 resident warm"; do echo "threads=$th $args"; AVIFGPU_EXP_THREADS=$th /tmp/shuttle_pool $args 0 | cut -c100-230; done; done
echo done
