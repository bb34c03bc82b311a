"""Where a data-parallel training step spends its time at N ranks: the local step without the collective, the full step,
and the gradient all-reduce alone (in-library ncclAllReduce of the 1.63 M-float buffer), max over ranks, per per-GPU batch.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_breakdown.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from clairvoyante_b200 import _lib, clairvoyante_v3 as cv, parallel, synth, utils_v2  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = cv.Clairvoyante(device=local)
m.init(seed=0)
tr = parallel.DataParallelTrainer(m, dist)


def timed(fn, reps):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    lo = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return float(t) / reps * 1e3, float(lo) / reps * 1e3


out = dict(n_gpus=world)
for per_gpu in (1250, 10000):
    x, y = synth.make_labeled_sites(per_gpu * world, seed=3)
    x = utils_v2.with_counts(x)
    lo, hi = parallel.shard_range(len(x), rank, world)
    xs, ys = tr._shard(x, lo, hi), y[lo:hi]
    for i in range(3):
        tr.train(x, y, seed=i)
    local_ms = timed(lambda i: m._train_step(xs, ys, apply_update=0, seed=i), 20)
    dp_ms = timed(lambda i: tr.train(x, y, seed=i), 20)
    out["per_gpu_%d" % per_gpu] = dict(local_step_no_collective_ms=local_ms, dp_step_ms=dp_ms)
ar = timed(lambda i: _lib.check(m._lib.cvb_allreduce_gradients(m._h)), 100)
out["allreduce_alone_ms"] = ar
ad = timed(lambda i: m.applyAdam(), 20)      # Adam + the loss read-back (one host synchronisation), no collective
out["adam_readback_ms"] = ad
buf = torch.zeros(1631516, device="cuda")
out["torch_allreduce_same_size_ms"] = timed(lambda i: dist.all_reduce(buf), 100)   # torch.distributed's communicator, same bytes
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
m.close()
