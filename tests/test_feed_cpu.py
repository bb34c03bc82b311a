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


def _nasty_stream(n, seed):
    """rows with everything the tokeniser must survive: tabs / runs of blanks / CRLF, lower-case and N centre bases,
    blank lines, short and over-long rows, exponent / signed / integer / non-numeric tokens, no final newline"""
    rng = np.random.RandomState(seed)
    x = synth.make_sites(n, seed)
    rows = _rows(x, chrom="chrX", start=5)
    out = []
    for i, r in enumerate(rows):
        f = r.split(" ")
        k = rng.randint(0, 12)
        if k == 0:
            f[2] = f[2].lower()
        elif k == 1:
            f[2] = f[2][:16] + "N" + f[2][17:]
        elif k == 2:
            f = f[:-3]                                  # too few fields
        elif k == 3:
            f = f + ["1.0"]                             # too many fields
        elif k == 4:
            f[3 + rng.randint(528)] = "abc"             # not a number
        elif k == 5:
            j = 3 + rng.randint(528)
            f[j] = "%.3e" % float(f[j])                 # exponent form (strtod path)
        elif k == 6:
            j = 3 + rng.randint(528)
            f[j] = "+" + f[j].lstrip("-")
            f[j + 1 if j < 530 else j] = "7"            # bare integer
        sep = ["\t", "  ", " "][rng.randint(3)]
        line = sep.join(f)
        if k == 7:
            line = "  " + line + " \r"                  # leading blanks, CRLF
        out.append(line)
        if k == 8:
            out.append("")                              # blank line
        if k == 9:
            out.append("   \t ")
    return "\n".join(out)                               # NO trailing newline


@pytest.mark.parametrize("seed,num,threads", [(1, 1000, 0), (2, 37, 1), (3, 64, 3), (4, 5000, 0)])
def test_native_parser_matches_python_restatement(tmp_path, seed, num, threads, monkeypatch):
    """csrc/text_feed.cpp (through utils_v2.GetTensor) vs the row-by-row restatement of utils_v2.py:23-59 on the same bytes;
    small read size so that lines straddle read boundaries"""
    from oracle import feed_oracle as FO
    text = _nasty_stream(700 if seed != 4 else 3000, seed)
    fn = str(tmp_path / "t.txt")
    open(fn, "w").write(text)
    monkeypatch.setattr(U, "_READ_BYTES", 50021)
    got = list(U.GetTensor(fn, num, threads))
    ref = list(FO.GetTensor(fn, num))
    assert [(g[0], g[1]) for g in got] == [(r[0], r[1]) for r in ref]
    for g, r in zip(got, ref):
        assert g[2].dtype == np.float32 and g[2].shape == r[2].shape
        assert np.array_equal(g[2], r[2])
        assert g[3] == r[3]
    assert sum(g[1] for g in got) > 300


def test_native_parser_c_abi_edges():
    """direct C-ABI calls: empty input, partial last line with and without final_chunk, max_lines cap"""
    import ctypes
    from clairvoyante_b200 import _lib
    lib = _lib.load()
    row = ("c 7 " + "ACGT" * 8 + "A " + " ".join(["2.0", "3.0", "2.0", "1.0"] * 132)).encode()
    x = np.full((4, 528), -1, np.float32); meta = np.zeros((4, 10), np.int64)
    nl, nk, nu = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()

    def call(buf, final, cap):
        _lib.check(lib.cvb_parse_tensor_text(buf, len(buf), final, cap, 1, x.ctypes.data, meta.ctypes.data,
                                             ctypes.byref(nl), ctypes.byref(nk), ctypes.byref(nu)))
        return nl.value, nk.value, nu.value
    assert call(b"", 1, 4) == (0, 0, 0)
    assert call(row, 0, 4) == (0, 0, 0)                          # incomplete line is left for the next call
    assert call(row, 1, 4) == (1, 1, len(row))
    assert np.array_equal(x[0, :4], [2.0, 1.0, 0.0, -1.0])       # channels 1..3 relative to channel 0
    assert bytes(row[meta[0, 3]:meta[0, 3] + meta[0, 4]]) == b"c" and meta[0, 8] == 33
    three = row + b"\n" + row + b"\n" + row + b"\n"
    assert call(three, 0, 2) == (2, 2, 2 * (len(row) + 1))       # stops at max_lines, reports where to resume
    assert lib.cvb_parse_tensor_text(None, 5, 0, 1, 1, None, None, ctypes.byref(nl), ctypes.byref(nk), ctypes.byref(nu)) != 0
    assert b"NULL" in lib.cvb_last_error()


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


def test_old_zlib_container_still_reads():
    x = synth.make_sites(9, 2)
    b = U.pack_array_zlib(x)
    assert b[:5] == b"CVBZ1" and np.array_equal(U.unpack_array(b), x)


def test_tensor2bin_blosc_container_equals_default(tmp_path):
    """tensor2Bin --blosc (reference-format frames + protocol-2 pickles) holds exactly what the default container holds"""
    import types
    from clairvoyante_b200 import tensor2Bin
    x = synth.make_sites(7, 5)
    seqs = ["ACGTACGTACGTACGT" + c + "CGTACGTACGTACGTA" for c in "ACGTACG"]
    tfn, vfn = str(tmp_path / "t.txt"), str(tmp_path / "v.txt")
    open(tfn, "w").write("\n".join(_rows(x, seqs)) + "\n")
    open(vfn, "w").write("chr1 1000 A G 0 1\nchr1 1003 TACGTAC T 1 1\n")
    got = {}
    for blosc in (False, True):
        fn = str(tmp_path / ("d%d.bin" % blosc))
        tensor2Bin.Run(types.SimpleNamespace(tensor_fn=tfn, var_fn=vfn, bed_fn=None, bin_fn=fn, blosc=blosc))
        total, xb, yb, pb = U.load_bin(fn)
        blocks = [U.unpack_array(b) for b in pb]
        order = np.argsort(np.concatenate(blocks))                       # GetTrainingArray shuffles
        got[blosc] = (total, U.DecompressArray(xb, 0, total, total)[0][order], U.DecompressArray(yb, 0, total, total)[0][order],
                      bytes(xb[0][:5]))
    assert got[False][0] == got[True][0] == 7
    assert np.array_equal(got[False][1], got[True][1]) and np.array_equal(got[False][2], got[True][2])
    # both are Blosc-1 frames around Python-2 pickles; only --blosc applies python-blosc's default byte shuffle
    assert got[False][3][0] == 2 and got[True][3][0] == 2 and (got[False][3][2] & 1) == 0 and (got[True][3][2] & 1) == 1
