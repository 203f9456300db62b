#!/bin/bash
# final sanity of the committed state: the whole GPU suite
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2s_pytest.log; tail -3 gpurun_out/r2s_pytest.log
