"""Multi-GPU use of the hot path: one process per GPU (torch.distributed), no data-path collective for
inference, one SUM all-reduce per step for training.

Inference: the reference shards the genome into chunks and runs independent processes
(callVarBamParallel.py:77-89); here the candidate list is cut into contiguous ranges in rank order, so
concatenating the ranks' outputs reproduces the single-GPU order exactly.

Training: the loss is a SUM over the batch (clairvoyante_v3.py:140-151), so the gradient of a global batch is the
sum of the shard gradients; lambda * sum(0.5 w^2) is added once, inside the optimizer step, on every rank.
"""
import ctypes

import numpy as np


def shard_range(n, rank, world):
    """contiguous [lo, hi) of rank `rank` out of `world` over n items; sizes differ by at most one"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def predict_sharded(m, X, rank, world, gather=None):
    """predict this rank's shard of X; with `gather` (a callable taking the local (n_i,16) array and returning the
    list of all ranks' arrays, e.g. built on torch.distributed.all_gather_object) returns the full result on every
    rank, otherwise only the local shard.  Returns (base, zygosity, varType, indelLength)."""
    lo, hi = shard_range(len(X), rank, world)
    outs = m.predict(X[lo:hi])
    if gather is None:
        return outs
    local = np.concatenate(outs, axis=1)
    full = np.concatenate(gather(local), axis=0)
    return full[:, 0:4].copy(), full[:, 4:6].copy(), full[:, 6:10].copy(), full[:, 10:16].copy()


class _CudaBuffer(object):
    """exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)"""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = dict(shape=(int(numel),), typestr="<f4", data=(int(ptr), False), version=2)


class DataParallelTrainer(object):
    """train(X, Y) on the global batch: every rank takes its shard, gradients + loss sums are all-reduced (SUM) over
    NVLink/NCCL, then every rank applies the identical Adam step.

    The reduction runs INSIDE the library (cvb_allreduce_init, include/cvb200.h): ncclAllReduce on the library's compute
    stream between the backward pass and the optimiser kernel, one C call per step, no host synchronisation in between.
    torch.distributed only carries the 128-byte ncclUniqueId to the other ranks when the trainer is built.  On a backend
    without NCCL (the gloo CPU tests) or with in_library=False the first version's route is used: torch.distributed
    all_reduce on the exposed gradient buffer."""

    def __init__(self, model, dist=None, in_library=True):
        self.m = model
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.in_library = False
        self.grad = None
        if dist is None or self.world == 1:
            return
        lib = model._lib
        from . import _lib
        if in_library and dist.get_backend() == "nccl":
            uid = ctypes.create_string_buffer(128)
            if self.rank == 0:
                _lib.check(lib.cvb_nccl_unique_id(uid))
            box = [uid.raw]
            dist.broadcast_object_list(box, src=0)
            _lib.check(lib.cvb_allreduce_init(model._h, box[0], self.world, self.rank))
            self.in_library = True
        else:
            import torch
            self._torch = torch
            host = getattr(model, "gradTensor", None)
            if host is not None:            # a stand-in model (the gloo CPU tests) hands over its gradient buffer as a tensor
                self.grad = host()
                return
            ptr, numel = ctypes.c_void_p(), ctypes.c_int64()
            _lib.check(lib.cvb_grad_buffer(model._h, ctypes.byref(ptr), ctypes.byref(numel)))
            self.grad = torch.as_tensor(_CudaBuffer(ptr.value, numel.value), device="cuda:%d" % model.device)

    @staticmethod
    def _shard(X, lo, hi):
        """rows [lo, hi) of a batch; a CountBatch keeps its raw counts (slices of it alone would drop them)"""
        c = getattr(X, "counts", None)
        xs = X[lo:hi]
        if c is not None and hasattr(xs, "counts"):
            xs.counts = c[lo:hi]
        return xs

    def train(self, X, Y, seed):
        lo, hi = shard_range(len(X), self.rank, self.world)
        X = self._shard(X, 0, len(X))      # (view; keeps the caller's object untouched)
        # the dropout stream is indexed by (seed, element); give every rank a distinct, reproducible stream
        seed = (seed + 0x51ED270B * self.rank) & 0xFFFFFFFFFFFFFFFF
        if self.in_library or self.world == 1:
            return self.m._train_step(self._shard(X, lo, hi), Y[lo:hi], apply_update=1, seed=seed)
        self.m._train_step(self._shard(X, lo, hi), Y[lo:hi], apply_update=0, seed=seed)
        self.dist.all_reduce(self.grad, op=self.dist.ReduceOp.SUM)
        if self.grad.is_cuda:
            # NCCL enqueues the reduction on torch's stream and returns; the optimizer kernels run on the library's
            # own (non-blocking) stream, so the reduced gradients must be complete before applyAdam reads them
            self._torch.cuda.current_stream(self.grad.device).synchronize()
        return self.m.applyAdam()
