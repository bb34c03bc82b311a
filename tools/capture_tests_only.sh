#!/bin/bash
# the GPU suite + smoke on the current tree (gpurun --timeout 900 -- bash tools/capture_tests_only.sh)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_tests_final.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 >> gpurun_out/r02_tests_final.log
cat gpurun_out/r02_tests_final.log
