#!/bin/bash
bash tools/capture_round.sh
timeout 200 python tools/train_small_probe.py 2>&1 | tail -1 | tee gpurun_out/r02q_train_probe.log
timeout 200 python tools/train_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02q_train_probe.log
