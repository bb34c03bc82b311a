"""ms per training step through Clairvoyante.train for several batch sizes, pageable vs pinned host arrays (the per-rank work of
a data-parallel step at the reference's global batch of 10,000 on 8 GPUs is 1,250 tensors).  python tools/train_small_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clairvoyante_b200 import clairvoyante_v3 as cv, synth
m = cv.Clairvoyante(); m.init(seed=1)
out = {}
for n in (1250, 2500, 5000, 10000):
    x, y = synth.make_labeled_sites(n, 3)
    xp = torch.from_numpy(x).pin_memory().numpy(); yp = torch.from_numpy(y).pin_memory().numpy()
    from clairvoyante_b200 import utils_v2
    for name, (a, b) in (("pageable", (x, y)), ("pinned", (xp, yp)), ("counts_pageable", (utils_v2.with_counts(x), y))):
        for _ in range(4):
            m.train(a, b)
        t0 = time.perf_counter()
        for _ in range(20):
            m.train(a, b)
        out["%d/%s" % (n, name)] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
    t0 = time.perf_counter()
    for _ in range(20):
        m.getLoss(xp, yp)
    out["%d/getLoss_pinned" % n] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
print(json.dumps(out))
