"""Data-parallel training throughput (BASELINE config "v3 training ... 8xB200 with NCCL grad allreduce"): one process per
GPU, parallel.DataParallelTrainer (shard the batch, one SUM all-reduce of the flat gradient buffer, identical Adam step on
every rank).  Times K steps between barriers with a device synchronize on both sides, max over ranks; rank 0 prints one
JSON line per batch size.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
       tools/train_bench_dp.py [steps] [global_batch ...]          (N = 1 works without torchrun)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairvoyante_b200 import clairvoyante_v3 as cv, param, parallel, synth   # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    batches = [int(a) for a in sys.argv[2:]] or [param.trainBatchSize, param.trainBatchSize * world]
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = cv.Clairvoyante(device=local)
    m.init(seed=0)
    tr = parallel.DataParallelTrainer(m, dist)
    for gb in batches:
        x, y = synth.make_sites(gb, 1), synth.make_labels(gb, 1)
        for i in range(3):
            tr.train(x, y, seed=i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t = time.perf_counter()
        for i in range(steps):
            loss, _ = tr.train(x, y, seed=100 + i)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], device="cuda")
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(dt.item()) / steps * 1e3
            print(json.dumps(dict(metric="training tensors/sec (fwd + bwd + Adam)", variant="v3", n_gpus=world,
                                  global_batch=gb, per_gpu_batch=gb // world, steps=steps, ms_per_step=round(ms, 3),
                                  value=round(gb / ms * 1e3), train_mode=m.trainMode, last_loss=float(loss),
                                  collective="none" if world == 1 else "NCCL all-reduce SUM of 1,631,512 fp32 per step")),
                  flush=True)
    if dist is not None:
        dist.destroy_process_group()
    m.close()


if __name__ == "__main__":
    main()
