#!/bin/bash
mkdir -p gpurun_out
./scripts/ubench/ldl_diag_test 1 > gpurun_out/r2c_diag.log 2>&1
./scripts/ubench/ldl_diag_prof 1 >> gpurun_out/r2c_diag.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "solve or lm or venice or traj" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
timeout 300 python scripts/time_kernels.py > gpurun_out/r2c_times.log 2>&1
tail -3 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_times.log gpurun_out/r2c_diag.log
