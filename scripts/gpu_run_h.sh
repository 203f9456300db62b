#!/bin/bash
# where does the pinhole (dc = 9) Schur kernel spend its time: ncu --set full of one launch + the kernel's own cycle counters
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"schur5_kernel" -s 3 -c 1 -o gpurun_out/r2h_s9 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --camera pinhole > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
NLLS_B200_S5DBG=1 NLLS_B200_VERBOSE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --camera pinhole > gpurun_out/r2h_dbg9.json 2> gpurun_out/r2h_dbg9.err; echo "dbg9 rc=$?"
NLLS_B200_S5DBG=1 NLLS_B200_VERBOSE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_dbg6.json 2> gpurun_out/r2h_dbg6.err; echo "dbg6 rc=$?"
grep "schur5" gpurun_out/r2h_dbg9.err | tail -4; grep "schur5" gpurun_out/r2h_dbg6.err | tail -4
