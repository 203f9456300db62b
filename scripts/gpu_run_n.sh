#!/bin/bash
# camera pass at 3 CTAs per SM (80 registers, spills) vs 2 (128 registers), same box
mkdir -p gpurun_out; : > gpurun_out/r2n_ab.log
for rep in 1 2; do
for lib in nllssolver.jl_b200/libnlls_b200.so build/variants/libnlls_lincam3.so; do
NLLS_B200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['ms_per_step'], d['kernel_ms']['cost'], d['kernel_ms']['lin_cam'], d['kernel_ms']['lm_try'])" >> gpurun_out/r2n_ab.log
done; done
cat gpurun_out/r2n_ab.log
