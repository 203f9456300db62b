#!/bin/bash
# 2-GPU check after moving the reduced-system planning into reduced_plan.hpp: multi-rank tests + the solve tests + the 2-GPU bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "multirank or solve or venice or lm_" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err; echo "bench2 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2g_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms']['lm_try'], d['cost_trace'][:2])"
