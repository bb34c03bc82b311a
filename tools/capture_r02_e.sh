#!/bin/bash
# round 2, call 5: branch-free conv epilogue + constant-bank bias + split-K FC4: parity, per-kernel A/B, batch sweep, DP log
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_trained_parity.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02e_tests.log
timeout 90 python tools/ab_resident.py v3 1 2>&1 | tail -1 | tee gpurun_out/r02e_ab.log
timeout 90 python tools/ab_resident.py slim 1 2>&1 | tail -1 | tee -a gpurun_out/r02e_ab.log
timeout 400 python tools/batch_sweep.py 2>&1 | tail -24 | tee gpurun_out/r02e_sweep.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dp_check.py v3 > gpurun_out/r02e_dp_v3.log 2>&1; echo "dp v3 rc=$?"; grep "rank \|Error\|error\|dp_check" gpurun_out/r02e_dp_v3.log | head -20
