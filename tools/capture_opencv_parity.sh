#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trained_parity.py tests/test_oracle_opencv_cpu.py -q 2>&1 | tail -4 | tee gpurun_out/r02_opencv_parity.log
