set -x
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_async_sharded.py -m gpu -q 2>&1 | tail -3
for n in 8 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/r2_bench_n$n.err | tail -1 > gpurun_out/r2_bench_n$n.json
tail -3 gpurun_out/r2_bench_n$n.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n$n.json"))
print("N=$n value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
t=d.get("tile"); print(json.dumps({k:t[k] for k in ("identical_to_single_gpu","one_gpu_ms","convert_ms","gather","fused","speedup_vs_one_gpu","owner_nvlink_ingress_bytes")}, indent=0))
print(json.dumps(d.get("batch")))
for k,v in d.get("e2e_shuttle",{}).items():
    for kk,vv in v.items(): print(k, kk, vv.get("gpx_s"), vv.get("seconds_per_image"))
PY
done
echo done
