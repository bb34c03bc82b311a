"""GPU, BASELINE.json sizes: properties that do not need an oracle pass over millions of sites (the fp64 NumPy oracle
does ~300 sites/s): determinism, shard == whole, probability simplex, agreement of the two arithmetic paths, and an
oracle spot-check on a random sample."""
import os

import numpy as np
import pytest

from clairvoyante_b200 import initializers as I, synth
from oracle import cv_oracle as O

pytestmark = pytest.mark.gpu
N = int(os.environ.get("CVB_FULLSIZE_SITES", str(1 << 20)))      # 4194304 = configs[1]; default 1 Mi keeps the suite short


def test_full_size_properties_v3():
    import torch
    from clairvoyante_b200 import clairvoyante_v3 as cv
    W = I.init_weights("v3", 0)
    pool = synth.make_sites(65536, 77)
    x = torch.from_numpy(pool).cuda().repeat((N + 65535) // 65536, 1, 1, 1)[:N].contiguous()
    m = cv.Clairvoyante(); m.setWeights(W)
    st = torch.cuda.current_stream().cuda_stream
    outs = {}
    for mode in ("fp16x3", "fp32"):
        m.setComputeMode(mode)
        o = torch.empty((N, 16), device="cuda"); l = torch.empty((N, 16), device="cuda")
        m.predictDevice(x.data_ptr(), N, o.data_ptr(), l.data_ptr(), st)
        torch.cuda.synchronize()
        o2 = torch.empty_like(o)
        m.predictDevice(x.data_ptr(), N, o2.data_ptr(), None, st)
        torch.cuda.synchronize()
        assert torch.equal(o, o2), "not deterministic (%s)" % mode
        # tiling property: the synthetic stream repeats every 65,536 sites -> so must the outputs (other chunks / CTAs / lanes)
        assert torch.equal(o[:65536], o[65536:131072])
        last = 65536 * (N // 65536 - 1)
        assert torch.equal(o[:1000], o[last:last + 1000])
        # probability simplex per softmax head, sigmoid range
        for a, b in ((4, 6), (6, 10), (10, 16)):
            assert float((o[:, a:b].sum(1) - 1).abs().max()) < 1e-5
        assert float(o.min()) >= 0.0 and float(o.max()) <= 1.0 and bool(torch.isfinite(l).all())
        outs[mode] = (o, l)
    # the tensor-core path and the all-fp32 path agree far inside the 1e-3 budget, and on every clear argmax
    d = (outs["fp16x3"][1] - outs["fp32"][1]).abs().max()
    assert float(d) < 1e-3
    for a, b in ((0, 4), (4, 6), (6, 10), (10, 16)):
        l32 = outs["fp32"][1][:, a:b]
        top = torch.topk(l32, 2, dim=1).values
        clear = (top[:, 0] - top[:, 1]) > 2e-3
        assert bool((outs["fp16x3"][1][:, a:b].argmax(1) == l32.argmax(1))[clear].all())
    # shard == whole through the host API on a 300k prefix (3 ranks' worth of contiguous ranges)
    xh = x[:300000].cpu().numpy()
    whole = np.concatenate(m.predict(xh), 1)
    parts = [np.concatenate(m.predict(xh[a:b]), 1) for a, b in ((0, 100000), (100000, 200000), (200000, 300000))]
    assert np.array_equal(np.concatenate(parts), whole)
    # oracle spot check on random sites of the big batch
    idx = np.random.default_rng(3).integers(0, N, 256)
    ref = O.forward(W, pool[idx % 65536], "v3")
    assert float(np.abs(outs["fp32"][1][idx].cpu().numpy() - ref["logits"]).max()) <= 1e-3
    m.close()
