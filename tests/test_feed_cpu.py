"""CPU: the batch feed (utils_v2) against its contracts in reference utils_v2.py:23-59,62-207."""
import gzip
import os

import numpy as np
import pytest

from clairvoyante_b200 import param, synth, utils_v2 as U


def _rows(x, seqs=None, chrom="chr1", start=1000):
    """text rows in CreateTensor.py:56 format from channel-subtracted tensors (undo utils_v2.py:46 first)"""
    raw = x.copy()
    raw[..., 1:] += raw[..., 0:1]
    out = []
    for i, t in enumerate(raw):
        seq = (seqs[i] if seqs else "ACGTACGTACGTACGTACGTACGTACGTACGTA")
        out.append("%s %d %s %s" % (chrom, start + i, seq, " ".join("%0.1f" % v for v in t.reshape(-1))))
    return out


@pytest.mark.parametrize("gz", [False, True])
def test_get_tensor_batches(tmp_path, gz):
    x = synth.make_sites(23, 5)
    rows = _rows(x)
    rows.insert(5, "chr1 1 " + "A" * 16 + "N" + "A" * 16 + " " + " ".join(["0.0"] * 528))   # centre base not ACGT -> dropped
    rows.insert(9, "chr1 2 ACGT 1.0 2.0")                                                     # malformed -> reported, skipped
    fn = str(tmp_path / ("t.gz" if gz else "t.txt"))
    data = ("\n".join(rows) + "\n").encode()
    with (gzip.open(fn, "wb") if gz else open(fn, "wb")) as fh:
        fh.write(data)
    got = list(U.GetTensor(fn, 7))
    assert [g[0] for g in got] == [0, 0, 0, 1] and [g[1] for g in got] == [7, 7, 7, 2]
    xs = np.concatenate([g[2] for g in got])
    assert xs.dtype == np.float32 and xs.shape == (23, 33, 4, 4) and np.array_equal(xs, x)
    pos = sum([g[3] for g in got], [])
    assert pos[0] == "chr1:1000:ACGTACGTACGTACGTACGTACGTACGTACGTA" and len(pos) == 23
    assert got[0][2].base is not got[1][2].base       # each batch owns fresh storage (callVar.py overlaps consumers)


def test_get_tensor_empty_final_batch(tmp_path):
    x = synth.make_sites(14, 6)
    fn = str(tmp_path / "t.txt")
    open(fn, "w").write("\n".join(_rows(x)) + "\n")
    got = list(U.GetTensor(fn, 7))
    assert [(g[0], g[1]) for g in got] == [(0, 7), (0, 7), (1, 0)]       # utils_v2.py:56-59: empty tail is legal
    assert got[-1][2].shape == (0, 33, 4, 4) and got[-1][3] == []


def test_decompress_array_slicing():
    rng = np.random.default_rng(0)
    total = 1234
    data = rng.standard_normal((total, 16))
    bs = param.bloscBlockSize
    blocks = [U.pack_array(data[i:i + bs]) for i in range(0, total, bs)]
    for start, num in [(0, 500), (0, 1000), (10, 17), (499, 2), (500, 500), (990, 300), (1000, 234), (1200, 34), (1233, 1), (700, 10000)]:
        out, n, end = U.DecompressArray(blocks, start, num, total)
        exp_n = min(num, total - start)
        assert n == exp_n and end == (1 if start + num >= total else 0)
        assert np.array_equal(out, data[start:start + exp_n])
    with pytest.raises(ValueError):
        U.unpack_array(b"\x02\x01blosc-frame")


def test_get_training_array_labels(tmp_path):
    x = synth.make_sites(6, 9)
    seqs = ["ACGTACGTACGTACGT" + c + "CGTACGTACGTACGTA" for c in "ACGTAC"]
    tfn, vfn, bfn = [str(tmp_path / n) for n in ("t.txt", "v.txt", "b.bed")]
    open(tfn, "w").write("\n".join(_rows(x, seqs)) + "\n")
    # truth rows: ctg pos ref alt gt1 gt2 (dataPrepScripts/GetTruth.py:52-79)
    open(vfn, "w").write("chr1 1000 A G 0 1\nchr1 1001 C T 1 1\nchr1 1002 G GAC 0 1\nchr1 1003 TACGTAC T 1 1\nchr1 5000 A C 0 1\n")
    open(bfn, "w").write("chr1 900 1005\n")           # intervaltree [900, 1004): 1004 and 1005 fall outside
    total, xb, yb, pb = U.GetTrainingArray(tfn, vfn, bfn, shuffle=False)
    assert total == 4
    Y, _, _ = U.DecompressArray(yb, 0, total, total)
    X, _, _ = U.DecompressArray(xb, 0, total, total)
    P, _, _ = U.DecompressArray(pb, 0, total, total)
    assert [p.decode() for p in P] == ["chr1:1000", "chr1:1001", "chr1:1002", "chr1:1003"]
    assert np.array_equal(X, x[:4]) and Y.dtype == np.float64
    exp = np.zeros((4, 16))
    exp[0, [0, 2]] = 0.5; exp[0, 4] = 1; exp[0, 7] = 1; exp[0, 10] = 1            # het SNP A>G
    exp[1, 3] = 1; exp[1, 5] = 1; exp[1, 7] = 1; exp[1, 10] = 1                   # hom SNP C>T
    exp[2, 2] = 0.5; exp[2, 4] = 1; exp[2, 8] = 1; exp[2, 12] = 1                 # het insertion of 2
    exp[3, 5] = 1; exp[3, 9] = 1; exp[3, 15] = 1                                  # hom deletion of 6 (>4)
    assert np.array_equal(Y, exp)
    total2, _, yb2, _ = U.GetTrainingArray(tfn, None, None, shuffle=False)         # no truth: everything non-variant
    Y2, _, _ = U.DecompressArray(yb2, 0, total2, total2)
    assert total2 == 6 and (Y2[:, 5] == 1).all() and (Y2[:, 6] == 1).all() and (Y2[:, 10] == 1).all()
    assert [int(np.argmax(r[:4])) for r in Y2] == [0, 1, 2, 3, 0, 1]
