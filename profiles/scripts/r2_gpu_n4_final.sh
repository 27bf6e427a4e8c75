set -x
n=4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/r2_bench_n$n.err | tail -1 > gpurun_out/r2_bench_n$n.json
tail -3 gpurun_out/r2_bench_n$n.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n$n.json"))
print("N=$n value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
t=d.get("tile"); print(t["one_gpu_ms"], t["convert_ms"], t["fused"]["convert_plus_assemble_ms"], t["fused"]["ingress_gbs"], t["gather"]["convert_plus_assemble_ms"], t["identical_to_single_gpu"], t["fused"]["identical_to_single_gpu"])
print(d["batch"]["gpx_s"], d["batch"]["per_gpu_roofline_frac"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2972$n bench.py --impl reference --gpus $n --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
echo done
