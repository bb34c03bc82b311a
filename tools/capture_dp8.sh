cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/train_bench_dp.py 20 > gpurun_out/r1o_dp_bench8.log 2>&1
grep -h "metric\|Error\|error" gpurun_out/r1o_dp_bench8.log | tail -6
