"""Rows/s of the training feed's block reader (utils_v2.DecompressArray, reference utils_v2.py:189-207) for the three block
containers, on synthetic count tensors.  CPU only; run on the host that will feed the GPU:
    python tests/tools/feed_blocks_bench.py [rows] > profiles/rNN_feed_blocks_bench.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from clairvoyante_b200 import param, synth, utils_v2 as U  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    bs, batch = param.bloscBlockSize, param.trainBatchSize
    X = synth.make_sites(n, 1)
    Y = synth.make_labels(n, 1).astype(np.float64)
    out = {"rows": n, "block_rows": bs, "batch_rows": batch, "cores": os.cpu_count(), "raw_MB": round(X.nbytes / 1e6, 1), "containers": {}}
    for name, pack in (("lz4_unshuffled(default)", U.pack_array), ("lz4_shuffled(reference)", U.pack_array_blosc),
                       ("zlib(old)", U.pack_array_zlib)):
        t = time.perf_counter()
        xb = [pack(X[i:i + bs]) for i in range(0, n, bs)]
        yb = [pack(Y[i:i + bs]) for i in range(0, n, bs)]
        tp = time.perf_counter() - t
        best = 0.0
        for _ in range(3):
            t = time.perf_counter()
            k = 0
            for s in range(0, n, batch):
                a, m, _e = U.DecompressArray(xb, s, batch, n)
                U.DecompressArray(yb, s, batch, n)
                k += m
            best = max(best, k / (time.perf_counter() - t))
        assert np.array_equal(a, X[n - len(a):])
        out["containers"][name] = {"pack_s": round(tp, 3), "X_MB": round(sum(len(b) for b in xb) / 1e6, 1),
                                   "decompress_rows_per_s": round(best)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
