#!/bin/bash
# in-place W -> Z transform (single-load A fragments) against the classic formulation (S5_ZT6=0), same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "solve or lm or venice or traj or fuzz or irregular or pinhole" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_pytest.log
tail -3 gpurun_out/r2l_pytest.log
echo "== Z in place (default)" > gpurun_out/r2l_times.log
timeout 300 python scripts/time_kernels.py >> gpurun_out/r2l_times.log 2>&1
echo "== classic formulation (S5_ZT6=0)" >> gpurun_out/r2l_times.log
NLLS_B200_LIB=build/variants/libnlls_zt0.so timeout 300 python scripts/time_kernels.py >> gpurun_out/r2l_times.log 2>&1
echo "== Z in place, dbg" >> gpurun_out/r2l_times.log
NLLS_B200_S5DBG=1 timeout 300 python scripts/time_kernels.py 2>&1 | grep "schur5 cycles" | tail -1 >> gpurun_out/r2l_times.log
for lib in "" build/variants/libnlls_zt0.so; do
NLLS_B200_LIB=${lib:-nllssolver.jl_b200/libnlls_b200.so} python bench.py --steps 10 --warmup 3 --no-cpu-baseline --camera pinhole 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pinhole ${lib:-default}', d['kernel_ms']['schur'], d['kernel_ms']['lm_try'])" >> gpurun_out/r2l_times.log
done
cat gpurun_out/r2l_times.log
