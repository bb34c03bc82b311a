#!/bin/bash
# round 2, call 10: whole GPU suite after the clean-up, smoke, full bench line, sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02j_tests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02j_smoke.log
timeout 600 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -c 600 gpurun_out/r02j_bench.json; tail -3 gpurun_out/r02j_bench.err
timeout 400 python tools/batch_sweep.py 2>&1 | tail -22 | tee gpurun_out/r02j_sweep.log
