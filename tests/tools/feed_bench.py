"""Throughput of the native text-feed parser (csrc/text_feed.cpp) on this host, beside the row-by-row Python
restatement of the reference feed (oracle/feed_oracle.py).   python tests/tools/feed_bench.py [rows]"""
import ctypes
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from clairvoyante_b200 import _lib, synth, utils_v2 as U   # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    x = synth.make_sites(2000, 1)
    raw = x.copy()
    raw[..., 1:] += raw[..., 0:1]
    rows = ["chr1 %d %s %s" % (i, "ACGT" * 8 + "A", " ".join("%0.1f" % v for v in t.reshape(-1))) for i, t in enumerate(raw)]
    buf = (("\n".join(rows) + "\n") * (n // 2000)).encode()
    n = 2000 * (n // 2000)
    lib = _lib.load()
    xx = np.zeros((n, 528), np.float32)
    meta = np.zeros((n, 10), np.int64)
    a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    out = dict(host_cores=os.cpu_count(), rows=n, bytes=len(buf), parse_only={})
    for th in (1, 2, 4, 8, 16, 32, 64):
        if th > (os.cpu_count() or 1):
            break
        best = 0.0
        for _ in range(3):
            t = time.time()
            _lib.check(lib.cvb_parse_tensor_text(buf, len(buf), 1, n, th, xx.ctypes.data, meta.ctypes.data,
                                                 ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
            best = max(best, b.value / (time.time() - t))
        out["parse_only"][str(th)] = dict(rows_per_s=round(best), mb_per_s=round(best * len(buf) / n / 1e6))
    assert np.array_equal(xx[:2000].reshape(-1, 33, 4, 4), x)
    d = tempfile.mkdtemp()
    fn = os.path.join(d, "t.txt")
    open(fn, "wb").write(buf)
    err, sys.stderr = sys.stderr, open(os.devnull, "w")
    try:
        t = time.time()
        k = sum(g[1] for g in U.GetTensor(fn, 1000))
        out["GetTensor_rows_per_s"] = round(k / (time.time() - t))
        from oracle import feed_oracle as FO
        t = time.time()
        k = 0
        for g in FO.GetTensor(fn, 1000):
            k += g[1]
            if k >= 6000:
                break
        out["python_restatement_rows_per_s"] = round(k / (time.time() - t))
    finally:
        sys.stderr = err
        os.remove(fn)
        os.rmdir(d)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
