#!/bin/bash
# quick GPU check of a build between full captures: the GPU suite, per-kernel times of both variants, the batch sweep
# (gpurun --timeout 1200 -- bash tools/capture_check.sh)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/check_tests.log
for v in v3 slim; do timeout 120 python tools/ab_kernels.py $v 0 2>&1 | tail -1; done > gpurun_out/check_ab.log
timeout 120 python tools/ab_kernels.py slim fp16 CVB_COMPUTE=fp16 2>&1 | tail -1 >> gpurun_out/check_ab.log
timeout 400 python tools/batch_sweep.py v3 > gpurun_out/check_sweep.log 2>&1
cat gpurun_out/check_tests.log gpurun_out/check_ab.log; grep -E "^v3 (256|1000|4096|16384) " gpurun_out/check_sweep.log | cut -c1-160
