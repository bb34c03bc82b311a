#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/r02s_tests.log
grep -q " failed" gpurun_out/r02s_tests.log && exit 1
bash tools/capture_round.sh
timeout 200 python tools/train_small_probe.py 2>&1 | tail -1 | tee gpurun_out/r02s_train_probe.log
