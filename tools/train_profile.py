"""One warm + N timed training steps of one variant (for an ncu launch list of the training path).
   python tools/train_profile.py [variant] [steps] [batch]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairvoyante_b200 import param, synth   # noqa: E402


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "v3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n = int(sys.argv[3]) if len(sys.argv) > 3 else param.trainBatchSize
    if variant == "v3":
        from clairvoyante_b200 import clairvoyante_v3 as cv
    else:
        from clairvoyante_b200 import clairvoyante_v3_slim as cv
    x, y = synth.make_sites(n, 1), synth.make_labels(n, 1)
    m = cv.Clairvoyante()
    m.init(seed=0)
    m.train(x, y)
    t = time.time()
    for _ in range(steps):
        loss, _ = m.train(x, y)
    dt = (time.time() - t) / steps
    print("variant %s batch %d: %.3f ms/step, %.0f tensors/s, loss %.3f" % (variant, n, dt * 1e3, n / dt, loss))
    m.close()


if __name__ == "__main__":
    main()
