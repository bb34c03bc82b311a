"""2-rank NCCL check of parallel.DataParallelTrainer: DP step on a global batch == single-GPU step on the same batch
(dropout off), and sharded predict == single-GPU predict.  torchrun --nproc-per-node 2 tools/dp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from clairvoyante_b200 import clairvoyante_v3 as cv, initializers as I, parallel, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = I.init_weights("v3", 2)
x, y = synth.make_sites(10000, 1), synth.make_labels(10000, 1)
m = cv.Clairvoyante(dropoutRateFC4=0.0, device=local); m.init(seed=1); m.setWeights(W)
tr = parallel.DataParallelTrainer(m, dist)
loss, _ = tr.train(x, y, seed=5)
ref = cv.Clairvoyante(dropoutRateFC4=0.0, device=local); ref.init(seed=1); ref.setWeights(W)
lref, _ = ref.train(x, y)
wa, wb = m.getWeights(), ref.getWeights()
err = max(float(np.abs(wa[k] - wb[k]).max()) for k in wa)
def gather(loc):
    out = [None] * world; dist.all_gather_object(out, loc); return out
full = parallel.predict_sharded(m, x[:5001], rank, world, gather)
single = m.predict(x[:5001])
same = all(np.array_equal(a, b) for a, b in zip(full, single))
print("rank %d: DP loss %.4f single %.4f | max |w_dp - w_single| = %.3g | sharded predict == single: %s" % (rank, loss, lref, err, same), flush=True)
assert abs(loss - lref) <= 1e-5 * abs(lref) and err < 2e-6 and same
dist.destroy_process_group()
