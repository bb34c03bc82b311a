#!/bin/bash
# 8-GPU sanity of the driver's launch line (inference weak scaling with the count feed, slim, training DP through NCCL in the library)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo rc=$?; grep -c "" gpurun_out/r02_bench_n8.json; tail -c 1200 gpurun_out/r02_bench_n8.json; tail -4 gpurun_out/r02_bench_n8.err
