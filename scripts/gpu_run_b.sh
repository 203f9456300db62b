#!/bin/bash
mkdir -p gpurun_out
./scripts/ubench/ldl_diag_prof 1 > gpurun_out/r2b_diag_prof.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "solve or lm or venice or traj or multirank" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
for v in "NLLS_B200_PDL=0 NLLS_B200_BWD=levels" "NLLS_B200_PDL=0" "NLLS_B200_PDL=1"; do
  echo "== $v" >> gpurun_out/r2b_times.log
  env $v timeout 300 python scripts/time_kernels.py >> gpurun_out/r2b_times.log 2>&1
done
tail -3 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_times.log
