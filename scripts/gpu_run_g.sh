#!/bin/bash
# 2-GPU check of the exchange by tile ownership: the multi-rank tests (N ranks == 1 rank, collective termination), the 2-GPU bench line with
# the ownership exchange and with the plain all-reduce (NLLS_B200_XG_ALLREDUCE=1)
mkdir -p gpurun_out
NLLS_B200_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_multirank.py -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
for mode in own allreduce; do
if [ $mode = allreduce ]; then export NLLS_B200_XG_ALLREDUCE=1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_2gpu_$mode.json 2> gpurun_out/r2g_bench_2gpu_$mode.err; echo "bench2 $mode rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_bench_2gpu_$mode.json").read().strip().splitlines()[-1]); print("$mode", d["ms_per_step"], d["kernel_ms"]["schur"], d["kernel_ms"]["lm_try"], d["cost_trace"][:3])
PY
done
grep "exchange by ownership" gpurun_out/r2g_bench_2gpu_own.err | head -2
