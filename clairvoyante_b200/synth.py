"""Synthetic pileup tensors with the statistics of dataPrepScripts/CreateTensor.py
output after utils_v2.GetTensor's channel subtraction.

Shape (N,33,4,4) float32, NHWC, element (n,h,w,c) at n*528+h*16+w*4+c
(CreateTensor.py:20-21,37-40; utils_v2.py:45).  Values are integer-valued,
|x| <= 250 (dcov cap, CreateTensor.py:296).  Deterministic in (seed, start).
"""
import numpy as np

SITE_FLOATS = 33 * 4 * 4
_BLOCK = 8192          # sites per RNG block: site i depends only on (seed, i // _BLOCK)


def _block_raw(seed, b, n):
    """Raw CreateTensor counts (before utils_v2.py:46's channel subtraction): non-negative integers <= 250 + 0.3*250."""
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
    return x


def _block(seed, b, n):
    x = _block_raw(seed, b, n)
    # utils_v2.py:46 -- subtract the reference channel from the other three
    x[..., 1:4] -= x[..., 0:1]
    return x


def _stream(block_fn, n, seed, start, dtype):
    out = np.empty((n, 33, 4, 4), dtype)
    i = 0
    while i < n:
        g = start + i
        b, off = divmod(g, _BLOCK)
        blk = block_fn(seed, b, _BLOCK)
        take = min(n - i, _BLOCK - off)
        out[i:i + take] = blk[off:off + take]
        i += take
    return out


def make_sites(n, seed=0, start=0):
    """Sites [start, start+n) of the infinite seeded stream."""
    return _stream(_block, n, seed, start, np.float32)


def make_counts(n, seed=0, start=0):
    """The same sites as RAW counts, int16 (what CreateTensor.py:56 prints, before utils_v2.py:46):
    make_counts(...) with channels 1..3 minus channel 0, as float32, equals make_sites(...) exactly."""
    return _stream(_block_raw, n, seed, start, np.int16)


def make_labeled_sites(n, seed=0):
    """(x, y): sites whose CENTRE row (position 16) carries an implanted genotype and the (N,16) labels that describe it
    (encoding of utils_v2.py:78-121,141-148), so that a network can actually learn the mapping -- used to produce
    trained-like weights (tests/golden/make_trained_weights.py) and by the training bench.  2/3 non-variant, 1/6 0/1,
    1/6 1/1; variants are SNP / INS / DEL with equal odds, indel length 1..5 shown as counts on the following rows."""
    x = _stream(_block_raw, n, seed, 0, np.float32)
    rng = np.random.default_rng([seed, 0x5EED])
    y = np.zeros((n, 16), np.float32)
    c = 16
    i = np.arange(n)
    d = np.maximum(x[:, c, :, 0].sum(1), 8.0)                      # depth at the centre
    ref = x[:, c, :, 0].argmax(1)
    kind = rng.integers(0, 6, n)                                   # 0..3 non-variant, 4 het, 5 hom
    vt = rng.integers(0, 3, n)                                     # 0 SNP, 1 INS, 2 DEL
    alt = (ref + rng.integers(1, 4, n)) % 4
    vlen = rng.integers(1, 6, n)
    frac = np.where(kind == 4, 0.35 + 0.3 * rng.random(n), 0.85 + 0.15 * rng.random(n))
    moved = np.floor(d * frac).astype(np.float32)
    # clean centre: every channel shows the reference base at full depth
    x[:, c] = 0.0
    for ch in range(4):
        x[i, c, ref, ch] = d
    var = kind >= 4
    snp = var & (vt == 0)
    x[i[snp], c, ref[snp], 3] -= moved[snp]; x[i[snp], c, alt[snp], 3] += moved[snp]
    x[i[snp], c, ref[snp], 1] -= moved[snp]; x[i[snp], c, alt[snp], 1] += moved[snp]
    for k, chn in ((1, 1), (2, 2)):                                # insertions on channel 1, deletions on channel 2
        sel = var & (vt == k)
        for L in range(1, 6):
            s = sel & (vlen >= L)
            rows = c + L - 1 if k == 1 else c + L
            base = rng.integers(0, 4, n) if k == 1 else x[:, rows, :, 0].argmax(1)
            x[i[s], rows, base[s], chn] += moved[s]
    x[..., 1:4] -= x[..., 0:1]
    nv = ~var
    y[i[nv], ref[nv]] = 1.0; y[nv, 5] = 1.0; y[nv, 6] = 1.0; y[nv, 10] = 1.0           # utils_v2.py:142-147
    het, hom = kind == 4, kind == 5
    y[i[het], ref[het]] = 0.5                                                            # utils_v2.py:90-96
    hs = het & snp
    y[i[hs], alt[hs]] = 0.5
    y[het, 4] = 1.0
    ho = hom & snp                                                                       # utils_v2.py:98-103
    y[i[ho], alt[ho]] = 1.0
    y[hom, 5] = 1.0
    y[i[var], np.where(vt == 0, 7, np.where(vt == 1, 8, 9))[var]] = 1.0
    L = np.where(vt == 0, 0, vlen)
    y[i[var], np.where(L > 4, 15, 10 + L)[var]] = 1.0
    return x, y


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
