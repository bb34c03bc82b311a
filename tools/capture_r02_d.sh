#!/bin/bash
# round 2, call 4: graph-captured training (tests + bench), DP / CLI tests, ncu source-level capture of conv2 / conv3 / FC4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_cli_gpu.py tests/test_dp_gpu.py tests/test_zz_reference_graph_gpu.py -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02d_tests.log
timeout 300 python tools/train_bench.py 2>&1 | tail -3 | tee gpurun_out/r02d_train.log
CVB_TRAIN_GRAPH=0 timeout 300 python tools/train_bench.py 2>&1 | tail -3 | tee -a gpurun_out/r02d_train.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_slab|k_fc4_tc" --launch-skip 9 -c 3 -o gpurun_out/r02d_src -f python bench.py --steps 1 --warmup 1 --sites 151552 --cpu-seconds 1 --no-extra > gpurun_out/r02d_ncu.log 2>&1
ls -la gpurun_out/r02d_src.ncu-rep; tail -2 gpurun_out/r02d_ncu.log
