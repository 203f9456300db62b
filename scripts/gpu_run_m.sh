#!/bin/bash
# round-2 (e) evidence run: GPU tests, bench lines (affine + pinhole), launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2q_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 3 --camera pinhole --no-cpu-baseline > gpurun_out/r2q_bench_pinhole.json 2> gpurun_out/r2q_bench_pinhole.err; echo "bench pinhole rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"schur5_kernel|lin_point_kernel|backsub_kernel|lin_cam_kernel|ldl_diag_kernel" -c 10 -o gpurun_out/r2q_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"schur5_kernel" -s 3 -c 1 -o gpurun_out/r2q_full_pinhole -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --camera pinhole > gpurun_out/r2q_ncu_fp.log 2>&1
tail -3 gpurun_out/r2q_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/r2q_bench.json","gpurun_out/r2q_bench_pinhole.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"])
PY
