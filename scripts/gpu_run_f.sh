#!/bin/bash
# ND order check: solve-related GPU tests, affine bench, pinhole launch list (where does the pinhole Schur phase go?)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "solve or lm or venice or traj or fuzz or iterators or irregular" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2f_pytest.log
NLLS_B200_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_pinhole.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --camera pinhole > gpurun_out/r2f_ncu_l.log 2>&1
tail -3 gpurun_out/r2f_pytest.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["kernel_ms"])
PY
grep "reduced system" gpurun_out/r2f_bench.err | head -2
