#!/bin/bash
# round-2 (b) evidence run: GPU tests, bench lines (affine + pinhole), launch list, ncu --set full of the top kernels, sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 3 --camera pinhole --no-cpu-baseline > gpurun_out/r2e_bench_pinhole.json 2> gpurun_out/r2e_bench_pinhole.err; echo "bench pinhole rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"schur5_kernel|lin_point_kernel|backsub_kernel|lin_cam_kernel|ldl_diag_kernel" -c 10 -o gpurun_out/r2e_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu_f.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "many_tiles" > gpurun_out/r2e_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2e_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "many_tiles" > gpurun_out/r2e_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2e_racecheck.log
tail -3 gpurun_out/r2e_pytest.log
