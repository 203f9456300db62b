#!/bin/bash
# Z formulation with the transform dealt to the consumer warps one tile ahead (default) vs the classic formulation, same box
mkdir -p gpurun_out; : > gpurun_out/r2n_ab.log
timeout 900 python -m pytest tests -m gpu -x -q -k "solve or lm or venice or traj or fuzz or irregular or pinhole" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2n_pytest.log
for rep in 1 2; do
for lib in nllssolver.jl_b200/libnlls_b200.so build/variants/libnlls_bands_classic.so build/variants/libnlls_2b039eb.so; do
NLLS_B200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['ms_per_step'], d['kernel_ms']['schur'], d['kernel_ms']['lm_try'])" >> gpurun_out/r2n_ab.log
done; done
NLLS_B200_S5DBG=1 timeout 300 python scripts/time_kernels.py 2>&1 | grep "schur5 cycles" | tail -1 >> gpurun_out/r2n_ab.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --camera pinhole 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pinhole', d['kernel_ms']['schur'], d['kernel_ms']['lm_try'])" >> gpurun_out/r2n_ab.log
cat gpurun_out/r2n_ab.log
