#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it:  gpurun --gpus N -- bash tools/capture_r02_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo rc=$? lines=$(grep -c "" gpurun_out/r02_bench_n$N.json)
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_n$N.json") if l.startswith("{")][-1])
t = d["train"]
print("n=$N value %.1fM e2e %.1fM (%.3f) slim fp16x3 %.1fM fp16 %.1fM  train global10k %.2fM per-gpu10k %.2fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["value"]/d["value"], d["slim"]["fp16x3"]["value"]/1e6, d["slim"]["fp16"]["value"]/1e6, t["global_batch_10000"]["value"]/1e6, t["per_gpu_batch_10000"]["value"]/1e6))
PY
