set -x
nvidia-smi -L
python -m pytest tests/test_gpu_async_sharded.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/r2_bench_n2.err | tail -1 > gpurun_out/r2_bench_n2.json
tail -5 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n2.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("tile"), indent=1))
print(json.dumps(d.get("batch"), indent=1))
print(json.dumps(d.get("e2e_shuttle"), indent=1))
PY
echo done
