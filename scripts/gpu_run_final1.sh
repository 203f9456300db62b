#!/bin/bash
# Final-shape (config C5) on ONE GPU: the trajectory the 8-GPU run is compared with
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --workload final --no-cpu-baseline > gpurun_out/r2p_bench_final_1gpu.json 2> gpurun_out/r2p_bench_final_1gpu.err; echo "final1 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2p_bench_final_1gpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms'])"
NLLS_B200_S5DBG=1 NLLS_B200_VERBOSE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --camera pinhole 2>&1 >/dev/null | grep "schur5\|schur v5 plan" | tail -3
