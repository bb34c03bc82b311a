"""Synthetic pileup tensors with the statistics of dataPrepScripts/CreateTensor.py
output after utils_v2.GetTensor's channel subtraction.

Shape (N,33,4,4) float32, NHWC, element (n,h,w,c) at n*528+h*16+w*4+c
(CreateTensor.py:20-21,37-40; utils_v2.py:45).  Values are integer-valued,
|x| <= 250 (dcov cap, CreateTensor.py:296).  Deterministic in (seed, start).
"""
import numpy as np

SITE_FLOATS = 33 * 4 * 4
_BLOCK = 8192          # sites per RNG block: site i depends only on (seed, i // _BLOCK)


def _block(seed, b, n):
    rng = np.random.default_rng([seed, b])
    H = 33
    d = np.minimum(rng.poisson(40.0, (n, H)), 250).astype(np.float32)          # depth per position
    ref = rng.integers(0, 4, (n, H))
    x = np.zeros((n, H, 4, 4), np.float32)
    ii, hh = np.meshgrid(np.arange(n), np.arange(H), indexing="ij")
    x[ii, hh, ref, 0] = d                                                      # ch0: reference base counts
    # ch3: query-base counts = ch0 with 5% of positions moving part of the depth to an alt base
    x[..., 3] = x[..., 0]
    mut = rng.random((n, H)) < 0.05
    moved = np.floor(d * rng.random((n, H)) * 0.5) * mut
    alt = (ref + rng.integers(1, 4, (n, H))) % 4
    x[ii, hh, ref, 3] -= moved
    x[ii, hh, alt, 3] += moved
    # ch1: ch3 + sparse insertion counts ; ch2: ch0 + sparse deletion counts at the ref base
    ins = np.floor(d * 0.3 * rng.random((n, H))) * (rng.random((n, H)) < 0.02)
    insb = rng.integers(0, 4, (n, H))
    x[..., 1] = x[..., 3]
    x[ii, hh, insb, 1] += ins
    dele = np.floor(d * 0.3 * rng.random((n, H))) * (rng.random((n, H)) < 0.02)
    x[..., 2] = x[..., 0]
    x[ii, hh, ref, 2] += dele
    # utils_v2.py:46 -- subtract the reference channel from the other three
    x[..., 1:4] -= x[..., 0:1]
    return x


def make_sites(n, seed=0, start=0):
    """Sites [start, start+n) of the infinite seeded stream."""
    out = np.empty((n, 33, 4, 4), np.float32)
    i = 0
    while i < n:
        g = start + i
        b, off = divmod(g, _BLOCK)
        blk = _block(seed, b, _BLOCK)
        take = min(n - i, _BLOCK - off)
        out[i:i + take] = blk[off:off + take]
        i += take
    return out


def make_labels(n, seed=0):
    """(N,16) float32 labels with the encoding of utils_v2.py:78-121,141-148;
    P(non-variant) = 2/3 (PairWithNonVariants --amp 2)."""
    rng = np.random.default_rng([seed, 0x1ABE1])
    y = np.zeros((n, 16), np.float32)
    kind = rng.integers(0, 6, n)             # 0..3 non-variant, 4 het, 5 hom
    ref = rng.integers(0, 4, n)
    alt = (ref + rng.integers(1, 4, n)) % 4
    vt = rng.integers(0, 3, n)               # 0 SNP, 1 INS, 2 DEL
    vlen = rng.integers(1, 6, n)
    for i in range(n):
        if kind[i] < 4:                      # utils_v2.py:142-147
            y[i, ref[i]] = 1.0; y[i, 5] = 1.0; y[i, 6] = 1.0; y[i, 10] = 1.0
            continue
        snp = vt[i] == 0
        if kind[i] == 4:                     # 0/1 (utils_v2.py:90-96)
            y[i, ref[i]] = 0.5
            if snp:
                y[i, alt[i]] = 0.5
            y[i, 4] = 1.0
        else:                                # 1/1 (utils_v2.py:98-103)
            if snp:
                y[i, alt[i]] = 1.0
            y[i, 5] = 1.0
        y[i, 7 if snp else (8 if vt[i] == 1 else 9)] = 1.0
        L = 0 if snp else int(vlen[i])
        y[i, 15 if L > 4 else 10 + L] = 1.0
    return y
