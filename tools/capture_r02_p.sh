#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_zz_reference_graph_gpu.py tests/test_cli_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02p_tests.log
timeout 200 python tools/train_small_probe.py 2>&1 | tail -1 | tee gpurun_out/r02p_train_probe.log
CVB_TRAIN_AUX=0 timeout 200 python tools/train_small_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r02p_train_probe.log
timeout 200 python tools/train_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02p_train_probe.log
