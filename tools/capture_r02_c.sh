#!/bin/bash
# round 2, call 3: DP check with full logs (2 GPUs), fixed CLI test, satfinite / chunk-size A/B, new bench line
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dp_check.py v3 > gpurun_out/r02c_dp_v3.log 2>&1; echo "dp v3 rc=$?"; grep -v "^W\|^\*\*\*" gpurun_out/r02c_dp_v3.log | tail -30
timeout 600 python -m pytest tests/test_cli_gpu.py tests/test_dp_gpu.py -x -q -m gpu -s 2>&1 | tail -12 | tee gpurun_out/r02c_tests.log
timeout 90 python tools/ab_resident.py v3 1 2>&1 | tail -1 | tee gpurun_out/r02c_ab.log
CVB_CHUNK_PER_SM=192 timeout 90 python tools/ab_resident.py v3 1 CVB_CHUNK_PER_SM=192 2>&1 | tail -1 | tee -a gpurun_out/r02c_ab.log
timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 3000 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
