#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "solve or lm or venice or traj or fuzz or irregular" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest.log
timeout 600 python scripts/time_dense.py 480 60000 > gpurun_out/r2t_dense_480.json 2> gpurun_out/r2t_dense_480.err; python -c "import json; d=json.load(open('gpurun_out/r2t_dense_480.json')); print(480, d['solve_reduced_ms'], d['reduced_solve_tflops'])"
timeout 600 python scripts/time_dense.py 1200 120000 > gpurun_out/r2t_dense_1200.json 2> gpurun_out/r2t_dense_1200.err; python -c "import json; d=json.load(open('gpurun_out/r2t_dense_1200.json')); print(1200, d['solve_reduced_ms'], d['reduced_solve_tflops'])"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('venice', d['ms_per_step'], d['kernel_ms']['reduced_solve'], d['kernel_ms']['lm_try'])"
