#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
timeout 300 python scripts/time_kernels.py > gpurun_out/r2k_times.log 2>&1
NLLS_B200_S5DBG=1 timeout 300 python scripts/time_kernels.py 2>&1 | grep "schur5 cycles" | tail -1 >> gpurun_out/r2k_times.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --camera pinhole 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pinhole', d['kernel_ms'])" >> gpurun_out/r2k_times.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench.json 2>gpurun_out/r2k_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1]); print('affine', d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'])" >> gpurun_out/r2k_times.log
cat gpurun_out/r2k_times.log
