#!/bin/bash
# 8-GPU evidence (run with gpurun --gpus 8): N ranks == 1 rank at N = 8, Venice-shape and Final-shape (config C5) bench lines
mkdir -p gpurun_out
NLLS_TEST_RANKS=8 timeout 600 python -m pytest tests/test_gpu_multirank.py -q -k "nrank" > gpurun_out/r2o_pytest_8rank.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2o_pytest_8rank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_8gpu.json 2> gpurun_out/r2o_bench_8gpu.err; echo "venice8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 10 --warmup 3 --workload final --no-cpu-baseline > gpurun_out/r2o_bench_final_8gpu.json 2> gpurun_out/r2o_bench_final_8gpu.err; echo "final8 rc=$?"
tail -2 gpurun_out/r2o_pytest_8rank.log
python - <<'PY'
import json
for f in ("gpurun_out/r2o_bench_8gpu.json","gpurun_out/r2o_bench_final_8gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["kernel_ms"], d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
