#!/bin/bash
# gpurun --gpus N --timeout 600 -- bash tools/capture_dp_breakdown.sh N
cd $GRAFT_REPO_ROOT
N=${1:-8}
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_n$N.%p.log timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/dp_breakdown.py > gpurun_out/dp_breakdown_n$N.json 2> gpurun_out/dp_breakdown_n$N.err
cat gpurun_out/dp_breakdown_n$N.json; tail -3 gpurun_out/dp_breakdown_n$N.err
f=$(ls gpurun_out/nccl_n$N.*.log | head -1); grep -c "" $f; grep -i "nvls\|algo\|Connected all\|channels\|P2P/CUMEM\|via" $f | cut -c1-200 | sort | uniq -c | sort -rn | head -25
rm -f gpurun_out/nccl_n$N.*.log
