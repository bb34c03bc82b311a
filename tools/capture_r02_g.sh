#!/bin/bash
# round 2, call 7: PDL chain + single-chunk fast path + barrier-token loads: parity; conv3 epilogue variants; PDL on/off sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_trained_parity.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02g_tests.log
for v in -1 0 1; do CVB_C3_VARIANT=$v timeout 90 python tools/ab_resident.py v3 1 CVB_C3_VARIANT=$v 2>&1 | tail -1; done | tee gpurun_out/r02g_ab.log
CVB_FC4_BOTH=1 timeout 90 python tools/ab_resident.py v3 1 CVB_FC4_BOTH=1 2>&1 | tail -1 | tee -a gpurun_out/r02g_ab.log
timeout 300 python tools/batch_sweep.py v3 2>&1 | tail -11 | tee gpurun_out/r02g_sweep_pdl.log
CVB_PDL=0 timeout 300 python tools/batch_sweep.py v3 2>&1 | tail -11 | tee gpurun_out/r02g_sweep_nopdl.log
