#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_forward_gpu.py -x -q -m gpu -k "count_feed or c_abi or dropout or adam" 2>&1 | tail -6 | tee gpurun_out/r02o_tests.log
timeout 200 python tools/train_small_probe.py 2>&1 | tail -1 | tee gpurun_out/r02o_train_probe.log
timeout 120 python tools/dp_check.py 2>/dev/null | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 tools/dp_check.py v3_slim 2>&1 | grep "rank \|dp_check" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_n2.json") if l.startswith("{")][-1])
print("n2 value", d["value"], "e2e", d["e2e"]["value"], "train", d["train"]["global_batch_10000"]["value"], d["train"]["per_gpu_batch_10000"]["value"])
PY
