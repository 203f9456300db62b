#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"schur5_kernel" -s 3 -c 1 -o gpurun_out/r2j_s6full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"
echo "== pinhole default (BR9=2)" > gpurun_out/r2j_times.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --camera pinhole 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['kernel_ms'])" >> gpurun_out/r2j_times.log
echo "== pinhole BR9=6" >> gpurun_out/r2j_times.log
NLLS_B200_LIB=build/variants/libnlls_br9_6.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --camera pinhole 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['kernel_ms'])" >> gpurun_out/r2j_times.log
NLLS_B200_LIB=build/variants/libnlls_br9_6.so timeout 600 python -m pytest tests -m gpu -x -q -k "pinhole" > gpurun_out/r2j_pytest_pin.log 2>&1; echo "pytest pinhole (BR9=6) rc=$?" >> gpurun_out/r2j_times.log
cat gpurun_out/r2j_times.log; tail -2 gpurun_out/r2j_pytest_pin.log
