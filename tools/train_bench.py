"""Training throughput (forward + backward + Adam through Clairvoyante.train, host batches of param.trainBatchSize
= 10,000 tensors, reference train.py:87-96 / README.md:309-319 "tensors/s").   python tools/train_bench.py [steps]"""
import json
import os
import sys
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairvoyante_b200 import param, synth   # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = param.trainBatchSize
    x, y = synth.make_sites(n, 1), synth.make_labels(n, 1)
    out = {}
    for variant in ("v3", "v3_slim"):
        if variant == "v3":
            from clairvoyante_b200 import clairvoyante_v3 as cv
        else:
            from clairvoyante_b200 import clairvoyante_v3_slim as cv
        m = cv.Clairvoyante()
        m.init(seed=0)
        for _ in range(3):
            m.train(x, y)
        t = time.time()
        for _ in range(steps):
            loss, _ = m.train(x, y)
        dt = (time.time() - t) / steps
        t = time.time()
        for _ in range(steps):
            m.getLoss(x, y)
        dl = (time.time() - t) / steps
        out[variant] = dict(batch=n, train_mode=m.trainMode, train_ms_per_step=round(dt * 1e3, 3),
                            train_tensors_per_s=round(n / dt), getloss_tensors_per_s=round(n / dl), last_loss=float(loss))
        m.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
