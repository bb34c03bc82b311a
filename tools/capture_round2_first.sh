#!/bin/bash
# First GPU call of the next round (about 6 GPU-minutes):
#   gpurun --timeout 900 -- bash tools/capture_round2_first.sh
# 1. the GPU tests added without a GPU at the end of round 1 (reference-graph fixture), then the whole GPU suite
# 2. A/B of the resident-weight conv kernels (CVB_CONV_RESIDENT), every setting under its own timeout
# 3. bench lines with the default kernels and, if step 2 was bit-identical, with the resident kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zz_reference_graph_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/r02_new_gpu_tests.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests.log
bash tools/capture_resident.sh
timeout 300 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
if grep -q '"bit_identical_to_default": true' gpurun_out/ab_resident_v3_3.json 2>/dev/null; then
  CVB_CONV_RESIDENT=3 timeout 300 python bench.py > gpurun_out/r02_bench_resident.json 2> gpurun_out/r02_bench_resident.err
  CVB_CONV_RESIDENT=7 timeout 300 python tools/train_bench.py > gpurun_out/r02_train_resident.json 2>&1
fi
tail -c 600 gpurun_out/r02_bench_default.json; echo; tail -c 600 gpurun_out/r02_bench_resident.json 2>/dev/null
