#!/bin/bash
# 8-GPU Venice-shape bench line with the exchange by tile ownership (run with gpurun --gpus 8)
mkdir -p gpurun_out
NLLS_B200_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_8gpu.json 2> gpurun_out/r2r_bench_8gpu.err; echo "venice8 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2r_bench_8gpu.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["kernel_ms"], d["e2e"]["ms_per_step"], d["cost_trace"][:3])
PY
grep "exchange by ownership" gpurun_out/r2r_bench_8gpu.err | head -8
