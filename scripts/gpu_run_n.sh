#!/bin/bash
# dc = 6: classic formulation (default) vs Z formulation with two producer warps (S5_ZT6=1, S5_PROD2=1: 10 consumers + 2 producers), same box
mkdir -p gpurun_out; : > gpurun_out/r2n_ab.log
NLLS_B200_LIB=build/variants/libnlls_zt_prod2.so timeout 600 python -m pytest tests -m gpu -x -q -k "solve or venice or fuzz or irregular or lm_" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest (variant) rc=$?"; tail -2 gpurun_out/r2n_pytest.log
for rep in 1 2; do
for lib in nllssolver.jl_b200/libnlls_b200.so build/variants/libnlls_zt_prod2.so; do
NLLS_B200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['ms_per_step'], d['kernel_ms']['schur'], d['kernel_ms']['lm_try'])" >> gpurun_out/r2n_ab.log
done; done
NLLS_B200_LIB=build/variants/libnlls_zt_prod2.so NLLS_B200_S5DBG=1 timeout 300 python scripts/time_kernels.py 2>&1 | grep "schur5 cycles" | tail -1 >> gpurun_out/r2n_ab.log
cat gpurun_out/r2n_ab.log
