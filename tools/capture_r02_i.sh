#!/bin/bash
# round 2, call 9: fp16 tolerance, conv2 bias variant, FC4 4-CTA clusters again, batch sweep (KCH 512 split), 2-GPU bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trained_parity.py tests/test_forward_gpu.py -x -q -m gpu -s -k "fp16 or slim" 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r02i_tests.log
timeout 90 python tools/ab_resident.py v3 1 2>&1 | tail -1 | tee gpurun_out/r02i_ab.log
CVB_C2_VARIANT=1 timeout 90 python tools/ab_resident.py v3 1 CVB_C2_VARIANT=1 2>&1 | tail -1 | tee -a gpurun_out/r02i_ab.log
CVB_TC_FC4_CLUSTER=4 timeout 90 python tools/ab_resident.py v3 1 CVB_TC_FC4_CLUSTER=4 2>&1 | tail -1 | tee -a gpurun_out/r02i_ab.log
timeout 400 python tools/batch_sweep.py 2>&1 | tail -22 | tee gpurun_out/r02i_sweep.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err; tail -c 2500 gpurun_out/r02i_bench_n2.json; tail -3 gpurun_out/r02i_bench_n2.err
