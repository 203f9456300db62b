#!/bin/bash
# multi-GPU bench lines (run with gpurun --gpus 8)
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2d_bench_${n}gpu.json 2> gpurun_out/r2d_bench_${n}gpu.err; echo "n=$n rc=$?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 10 --warmup 3 --workload final > gpurun_out/r2d_bench_final_8gpu.json 2> gpurun_out/r2d_bench_final_8gpu.err; echo "final rc=$?"
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.loads(open(f"gpurun_out/r2d_bench_{n}gpu.json").read().strip().splitlines()[-1]); print(n, d["ms_per_step"], d["kernel_ms"], d["e2e"]["ms_per_step"])
    except Exception as e: print(n, "ERR", e)
try:
    d=json.loads(open("gpurun_out/r2d_bench_final_8gpu.json").read().strip().splitlines()[-1]); print("final8", d["ms_per_step"], d["kernel_ms"])
except Exception as e: print("final ERR", e)
PY
