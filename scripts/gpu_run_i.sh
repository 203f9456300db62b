#!/bin/bash
# whole-window Schur mode (one entry per point, 7 consumer warps) against the three-band mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "solve or lm or venice or traj or fuzz or irregular" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2i_pytest.log
tail -3 gpurun_out/r2i_pytest.log
echo "== full (default)" > gpurun_out/r2i_times.log
timeout 300 python scripts/time_kernels.py >> gpurun_out/r2i_times.log 2>&1
echo "== bands" >> gpurun_out/r2i_times.log
NLLS_B200_LIB=build/variants/libnlls_bands.so timeout 300 python scripts/time_kernels.py >> gpurun_out/r2i_times.log 2>&1
echo "== full, 6 consumers" >> gpurun_out/r2i_times.log
NLLS_B200_S5_CONS=6 timeout 300 python scripts/time_kernels.py >> gpurun_out/r2i_times.log 2>&1
echo "== full, dbg" >> gpurun_out/r2i_times.log
NLLS_B200_S5DBG=1 NLLS_B200_VERBOSE=1 timeout 300 python scripts/time_kernels.py 2>&1 | grep -v "^\[nlls\] schur5 CTA" | tail -6 >> gpurun_out/r2i_times.log
cat gpurun_out/r2i_times.log
