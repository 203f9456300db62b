#!/bin/bash
# round-2 evidence run: GPU tests, bench lines (affine + pinhole), launch list, ncu --set full of the top kernels, sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 3 --camera pinhole > gpurun_out/r2a_bench_pinhole.json 2> gpurun_out/r2a_bench_pinhole.err; echo "bench pinhole rc=$?"
./scripts/ubench/ldl_diag_prof 1 > gpurun_out/r2a_diag_prof.log 2>&1
./scripts/ubench/ldl_diag_prof 32 >> gpurun_out/r2a_diag_prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"schur5_kernel|lin_point_kernel|backsub_kernel|lin_cam_kernel" -c 8 -o gpurun_out/r2a_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_f.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "many_tiles" > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2a_memcheck.log
tail -3 gpurun_out/r2a_pytest.log
