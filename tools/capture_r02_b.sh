#!/bin/bash
# round 2, call 2: new tests first, then the whole GPU suite, the slim A/B that call 1 lost to a script bug, smoke, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trained_parity.py tests/test_dp_gpu.py tests/test_cli_gpu.py -x -q -m gpu -s 2>&1 | tail -25 | tee gpurun_out/r02b_new_tests.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02b_gpu_tests.log
for s in 0 1; do timeout 90 python tools/ab_resident.py slim $s 2>&1 | tail -1; done | tee gpurun_out/r02b_ab_slim.log
timeout 90 python tools/ab_resident.py slim 0 CVB_SLIM_FC4_TC=1 2>&1 | tail -1 | tee -a gpurun_out/r02b_ab_slim.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02b_smoke.log
timeout 300 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 1500 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
