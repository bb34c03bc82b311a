cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r1l_dp_check.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_bench_dp.py 20 > gpurun_out/r1l_dp_bench2.log 2>&1
timeout 200 python tools/train_bench_dp.py 20 10000 > gpurun_out/r1l_dp_bench1.log 2>&1
grep -h "rank\|metric\|Error\|error" gpurun_out/r1l_dp_check.log gpurun_out/r1l_dp_bench2.log gpurun_out/r1l_dp_bench1.log | tail -12
