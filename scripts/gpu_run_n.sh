#!/bin/bash
# outlier kernel on the second stream (default) vs behind the Schur kernel (NLLS_B200_OUTLIER_SERIAL=1), same box
mkdir -p gpurun_out; : > gpurun_out/r2n_ab.log
timeout 300 python -m pytest tests -m gpu -x -q -k "solve or venice or irregular" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2n_pytest.log
for mode in side serial side serial; do
if [ $mode = serial ]; then export NLLS_B200_OUTLIER_SERIAL=1; else unset NLLS_B200_OUTLIER_SERIAL; fi
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['ms_per_step'], d['kernel_ms']['schur'], d['kernel_ms']['lm_try'])" >> gpurun_out/r2n_ab.log
done
cat gpurun_out/r2n_ab.log
