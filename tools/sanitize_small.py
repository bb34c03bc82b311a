"""compute-sanitizer target: one small predict per variant / mode / feed and one training step per variant.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clairvoyante_b200 import clairvoyante_v3 as cv, clairvoyante_v3_slim as cvs, initializers as I, synth, utils_v2 as U
for variant, mod, modes in (("v3", cv, ("fp16x3", "fp32")), ("v3_slim", cvs, ("fp16x3", "fp16", "fp32"))):
    m = mod.Clairvoyante(dropoutRateFC5=0.1)
    m.setWeights(I.init_weights(variant, 0))
    x, y = synth.make_labeled_sites(333, 1)
    for mode in modes:
        m.setComputeMode(mode)
        for feed in (x, U.with_counts(x), U.pack_counts(x).astype(np.int16), x.astype(np.float16)):
            m.predict(feed)
        m.predict(synth.make_sites(1000, 2))
        xs = synth.make_sites(1500, 3)                 # four tickets in flight, collected out of order, then the generator
        t = [m.predictSubmit(U.with_counts(xs[i:i + 375])) for i in range(0, 1500, 375)]
        for i in (1, 3, 0, 2):
            m.predictCollect(t[i])
        for _ in m.predictStream([xs[:500], xs[500:501], xs[501:]], depth=3):
            pass
    m.train(x, y); m.train(x, y); m.train(x, y); m.getLoss(x, y)
    m.close()
    print(variant, "ok", flush=True)
