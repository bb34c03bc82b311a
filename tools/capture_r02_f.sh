#!/bin/bash
# round 2, call 6: k_fc4_both + predicated pooling loads: parity, then per-kernel A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_trained_parity.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02f_tests.log
timeout 90 python tools/ab_resident.py v3 1 2>&1 | tail -1 | tee gpurun_out/r02f_ab.log
CVB_FC4_BOTH=0 timeout 90 python tools/ab_resident.py v3 1 CVB_FC4_BOTH=0 2>&1 | tail -1 | tee -a gpurun_out/r02f_ab.log
timeout 90 python tools/ab_resident.py slim 1 2>&1 | tail -1 | tee -a gpurun_out/r02f_ab.log
