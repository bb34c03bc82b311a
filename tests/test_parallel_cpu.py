"""CPU, world_size 2 over gloo: the multi-GPU host logic (site-list sharding, gather order) with a stub model."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from clairvoyante_b200 import parallel


def test_shard_ranges_cover_and_order():
    for n in (0, 1, 7, 8, 1000, 1001):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


class _Stub(object):
    def predict(self, X):
        s = X.reshape(len(X), 528).sum(1, keepdims=True).astype(np.float32)
        return s * np.ones((1, 4), np.float32), s * np.ones((1, 2), np.float32), s + np.zeros((1, 4), np.float32), s + np.zeros((1, 6), np.float32)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((37, 33, 4, 4)).astype(np.float32)

    def gather(local):
        out = [None] * world
        dist.all_gather_object(out, local)
        return out

    full = parallel.predict_sharded(_Stub(), X, rank, world, gather)
    single = _Stub().predict(X)
    ok = all(np.array_equal(a, b) for a, b in zip(full, single))
    lo, hi = parallel.shard_range(len(X), rank, world)
    local = parallel.predict_sharded(_Stub(), X, rank, world)
    ok = ok and len(local[0]) == hi - lo
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_predict_equals_single_process_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_predict_stream_keeps_depth_in_flight_and_drains_on_close():
    """cpu: host logic of ClairvoyanteBase.predictStream (model.py) over a stand-in for the submit / collect pair: results in
    submission order, never more than `depth` tickets outstanding, and none left when the consumer stops early"""
    from clairvoyante_b200.model import ClairvoyanteBase

    class Fake(object):
        predictStream = ClairvoyanteBase.predictStream

        def __init__(self):
            self.out, self.peak, self.log = [], 0, []

        def predictSubmit(self, X, want_logits=False):
            self.out.append(X)
            self.peak = max(self.peak, len(self.out))
            self.log.append(("s", X))
            return X

        def predictCollect(self, t):
            assert self.out[0] == t                     # oldest first
            self.out.pop(0)
            self.log.append(("c", t))
            return (t, t * 10)

    f = Fake()
    assert list(f.predictStream(range(7), depth=3)) == [(i, i * 10) for i in range(7)]
    assert f.peak == 3 and f.out == []
    assert f.log[:5] == [("s", 0), ("s", 1), ("s", 2), ("c", 0), ("s", 3)]     # a collect only once `depth` are in flight
    f = Fake()
    assert list(f.predictStream(range(2), depth=9)) == [(0, 0), (1, 10)] and f.peak == 2      # depth is capped at 4 slots
    f = Fake()
    assert list(f.predictStream(range(9), depth=9)) and f.peak == 4
    f = Fake()
    g = f.predictStream(range(10), depth=4)
    assert next(g) == (0, 0)
    g.close()                                           # the consumer walks away: everything submitted is still collected
    assert f.out == [] and [t for k, t in f.log if k == "s"] == [t for k, t in f.log if k == "c"]


class _TrainStub(object):
    """stand-in for the model in DataParallelTrainer's host logic: the 'gradient' of a shard is a fixed linear function of its
    rows (so shard gradients must SUM to the global one), the loss sums ride at the tail of the buffer as in the library
    (cvb_grad_buffer), and applyAdam is a plain step on the reduced buffer"""
    _lib = None
    device = 0

    def __init__(self):
        import torch
        self.g = torch.zeros(8 + 4, dtype=torch.float64)
        self.w = np.zeros(8)
        self.calls = []

    def gradTensor(self):
        return self.g

    def _train_step(self, X, Y, apply_update=1, seed=None):
        import torch
        self.calls.append((len(X), int(seed), apply_update))
        feat = X.reshape(len(X), -1)[:, :8].astype(np.float64)
        self.g[:8] = torch.from_numpy((feat * Y[:, :1]).sum(0))
        self.g[8:] = torch.from_numpy(Y[:, :4].astype(np.float64).sum(0))
        if apply_update:
            return self.applyAdam()

    def applyAdam(self):
        self.w -= 0.1 * self.g[:8].numpy()
        return np.float32(self.g[8:].sum().item()), {}


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((101, 33, 4, 4)).astype(np.float32)      # odd size: shards of 51 and 50
    Y = rng.standard_normal((101, 16)).astype(np.float32)
    single = _TrainStub()
    l1, _ = single._train_step(X, Y, 1, seed=7)
    m = _TrainStub()
    tr = parallel.DataParallelTrainer(m, dist)
    assert not tr.in_library                                         # gloo: the torch.distributed route
    l2, _ = tr.train(X, Y, seed=7)
    lo, hi = parallel.shard_range(len(X), rank, world)
    ok = (np.abs(m.w - single.w).max() < 1e-9 and abs(float(l1) - float(l2)) < 1e-3 * abs(float(l1))
          and m.calls == [(hi - lo, (7 + 0x51ED270B * rank) & 0xFFFFFFFFFFFFFFFF, 0)])      # its shard, its dropout stream, no local update
    q.put((rank, bool(ok), m.calls[0][1]))
    dist.destroy_process_group()


def test_data_parallel_trainer_host_logic_gloo():
    """cpu, world_size 2 over gloo: every rank trains on its contiguous shard with its own dropout stream, the SUM all-reduce
    of [gradients | loss sums] gives every rank the single-process step on the whole batch"""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[:2] for r in res] == [(0, True), (1, True)]
    assert res[0][2] != res[1][2]                                    # distinct dropout seeds per rank
