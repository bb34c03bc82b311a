#!/bin/bash
# the GPU suite (its two data-parallel tests need >= 2 GPUs), compute-sanitizer over the small cases, and the bench line at
# N GPUs as the driver launches it:   gpurun --gpus N --timeout 1500 -- bash tools/capture_final_multi.sh N
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_tests_2gpu.log
  timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -8 > gpurun_out/r02_sanitizer.log
  cat gpurun_out/r02_tests_2gpu.log gpurun_out/r02_sanitizer.log
else
  timeout 600 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -3 > gpurun_out/r02_tests_dp_n$N.log
  cat gpurun_out/r02_tests_dp_n$N.log
fi
bash tools/capture_r02_multi.sh $N
