"""CPU: the native alignment pile-up (csrc/pileup.cpp, replaces dataPrepScripts/CreateTensor.py) against the statement-by-
statement Python restatement oracle/createtensor_oracle.py on synthetic SAM text -- bit-exact counts, same set and order of
emitted centres -- plus the reference's edge cases: depth cap, MAPQ filter, minimum coverage, window at the region start,
left-edge option, soft clips / N / = / X ops, lower-case and N bases, chunked feeding."""
import gzip
import io
import os
import subprocess
import sys

import numpy as np
import pytest

from clairvoyante_b200 import CreateTensor as CT, utils_v2 as U
from oracle import createtensor_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synth_alignments(rng, ref_len=3000, n_reads=400, read_len=(60, 160), ops="MIDS=XN", lower=0.02, dup_pos=0.15):
    """random reference, position-sorted SAM rows with random CIGARs, candidate positions"""
    ref = "".join(rng.choice(list("ACGT"), size=ref_len))
    ref = "".join(c.lower() if rng.random() < lower else ("N" if rng.random() < 0.003 else c) for c in ref)
    starts = np.sort(rng.integers(0, ref_len - 200, size=n_reads))
    for i in range(1, n_reads):                       # runs of identical POS exercise the depth cap
        if rng.random() < dup_pos:
            starts[i] = starts[i - 1]
    rows = ["@HD\tVN:1.0\tSO:coordinate", "@SQ\tSN:ctg\tLN:%d" % ref_len]
    for ri, st in enumerate(starts):
        L = int(rng.integers(*read_len))
        cigar, seq, rp, left = [], [], int(st), L
        if "S" in ops and rng.random() < 0.3:
            k = int(rng.integers(1, 8)); cigar.append("%dS" % k); seq.append("".join(rng.choice(list("ACGT"), size=k)))
        while left > 0:
            r = rng.random()
            if r < 0.70 or not cigar or cigar[-1][-1] in "IDN":
                k = int(min(left, rng.integers(1, 40)))
                op = "M" if rng.random() < 0.8 else ("=" if rng.random() < 0.5 and "=" in ops else ("X" if "X" in ops else "M"))
                s = [ref[(rp + j) % ref_len].upper() if rng.random() > 0.05 else str(rng.choice(list("ACGTNa"))) for j in range(k)]
                seq.append("".join(s)); rp += k; left -= k
            elif r < 0.82 and "I" in ops:
                k = int(rng.integers(1, 6)); op = "I"; seq.append("".join(rng.choice(list("ACGT"), size=k))); left -= k
            elif r < 0.94 and "D" in ops:
                k = int(rng.integers(1, 6)); op = "D"; rp += k
            elif "N" in ops:
                k = int(rng.integers(5, 30)); op = "N"
            else:
                continue
            cigar.append("%d%s" % (k, op))
        mq = int(rng.integers(0, 61))
        rows.append("\t".join(["r%d" % ri, "0", "ctg", str(int(st) + 1), str(mq), "".join(cigar), "*", "0", "0", "".join(seq), "*"]))
    cands = sorted(set(int(c) for c in rng.integers(1, ref_len, size=ref_len // 25)))
    return ref, "\n".join(rows) + "\n", cands


def run_native(sam, ref, cands, ref_start=None, chunk=None, **opts):
    p = CT.Pileup(ref, cands, ref_start, **opts)
    b = sam.encode()
    if chunk:
        for i in range(0, len(b), chunk):
            p.feed(b[i:i + chunk])
        p.feed(b"", final=True)
    else:
        p.feed(b, final=True)
    c, x = p.take()
    st = p.stats()
    p.close()
    return c, x, st


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("opts", [dict(), dict(minMQ=20), dict(dcov=2), dict(minCoverage=4), dict(considerleftedge=False)])
def test_native_pileup_equals_restatement(seed, opts):
    rng = np.random.default_rng(seed)
    ref, sam, cands = synth_alignments(rng)
    okw = dict(min_mq=opts.get("minMQ", 0), dcov=opts.get("dcov", 250), min_coverage=opts.get("minCoverage", 0),
               consider_left_edge=opts.get("considerleftedge", True))
    want = O.create_tensors(sam, ref, cands, None, **okw)
    c, x, st = run_native(sam, ref, cands, None, **opts)
    assert [int(v) for v in c] == [w[0] for w in want]
    assert len(want) > 10
    for i, (ctr, t) in enumerate(want):
        assert np.array_equal(x[i], t), ctr
    assert st["malformed"] == 0 and st["open_centres"] == 0


def test_region_offset_and_window_at_region_start():
    """refStart set (CreateTensor.py:101-104): reference indexed relative to the fetched region; centres whose window would
    start before it are dropped (:54-55)"""
    rng = np.random.default_rng(7)
    ref, sam, cands = synth_alignments(rng, ref_len=2500, n_reads=300)
    ref_start = 11                                    # ref string starts at 1-based position 11
    sub = ref[ref_start - 1:]
    cands = sorted(set(cands + [12, 20, 27, 28, 29]))
    want = O.create_tensors(sam, sub, cands, ref_start)
    c, x, _ = run_native(sam, sub, cands, ref_start)
    assert [int(v) for v in c] == [w[0] for w in want] and all(np.array_equal(x[i], w[1]) for i, w in enumerate(want))
    assert all(ctr - (ref_start - 1) - 17 >= 0 for ctr in c)


def test_chunked_feed_and_odd_rows():
    rng = np.random.default_rng(11)
    ref, sam, cands = synth_alignments(rng, n_reads=250)
    sam = sam + "short\trow\n" + "\n" + sam.split("\n")[5] + "\n"          # malformed + blank + an out-of-order duplicate row
    want = O.create_tensors(sam, ref, cands)
    for chunk in (1, 7, 100, 4096):
        c, x, st = run_native(sam, ref, cands, chunk=chunk)
        assert [int(v) for v in c] == [w[0] for w in want]
        assert all(np.array_equal(x[i], w[1]) for i, w in enumerate(want))
        assert st["malformed"] == 1
    # no trailing newline on the last row
    c2, x2, _ = run_native(sam.rstrip("\n"), ref, cands, chunk=13)
    assert np.array_equal(c2, c) and np.array_equal(x2, x)


def test_empty_inputs():
    c, x, st = run_native("", "ACGT" * 50, [30, 60])
    assert len(c) == 0 and x.shape == (0, 33, 4, 4) and st["sam_rows"] == 0
    rng = np.random.default_rng(3)
    ref, sam, _ = synth_alignments(rng, n_reads=50)
    c, x, _ = run_native(sam, ref, [])
    assert len(c) == 0


def test_text_rows_feed_gettensor_and_direct_generator_agree(tmp_path):
    """rows printed in the reference format (:56) parse back through utils_v2.GetTensor to exactly what
    GetTensorFromAlignments yields in process (channel subtract, centre-base filter, position strings)"""
    rng = np.random.default_rng(5)
    ref, sam, cands = synth_alignments(rng, n_reads=300)
    rows = []
    for centers, X in CT.pileup_tensors(sam, ref, cands, None, batch=64):
        rows += CT.tensor_rows("ctg", centers, X, ref)
    want = O.create_tensors(sam, ref, cands)
    assert rows == [O.tensor_line("ctg", c, ref, None, t) for c, t in want]
    fn = tmp_path / "t.gz"
    with gzip.open(fn, "wt") as f:
        f.write("\n".join(rows) + "\n")
    err, sys.stderr = sys.stderr, open(os.devnull, "w")
    try:
        a = list(U.GetTensor(str(fn), 100))
    finally:
        sys.stderr = err
    b = list(CT.GetTensorFromAlignments(sam, ref, cands, "ctg", 100))
    assert len(a) == len(b) and len(a) >= 2
    for (e1, n1, x1, p1), (e2, n2, x2, p2) in zip(a, b):
        assert (e1, n1, p1) == (e2, n2, p2) and np.array_equal(x1, x2)


def test_command_line_on_sam_and_fasta(tmp_path):
    rng = np.random.default_rng(9)
    ref, sam, cands = synth_alignments(rng, n_reads=200)
    (tmp_path / "a.sam").write_text(sam)
    (tmp_path / "ref.fa").write_text(">other\nACGT\n>ctg desc\n" + "\n".join(ref[i:i + 60] for i in range(0, len(ref), 60)) + "\n")
    (tmp_path / "can.txt").write_text("".join("ctg %d\n" % c for c in cands) + "other 5\n")
    r = subprocess.run([sys.executable, "-m", "clairvoyante_b200.CreateTensor", "--bam_fn", str(tmp_path / "a.sam"), "--ref_fn",
                        str(tmp_path / "ref.fa"), "--can_fn", str(tmp_path / "can.txt"), "--ctgName", "ctg", "--samtools",
                        "/nonexistent/samtools"], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    want = O.create_tensors(sam, ref, cands)
    assert r.stdout.split("\n")[:-1] == [O.tensor_line("ctg", c, ref, None, t) for c, t in want]


def test_golden_alignment_fixture():
    """tests/golden/alignments.npz (make_golden_alignments.py): both oracles and both native stages reproduce the committed
    candidate rows and tensors"""
    from clairvoyante_b200 import ExtractVariantCandidates as EVC
    from oracle import candidates_oracle as OC
    g = np.load(os.path.join(ROOT, "tests", "golden", "alignments.npz"))
    sam, ref = str(g["sam"]), str(g["ref"])
    rows = [str(r) for r in g["candidate_rows"]]
    assert OC.make_candidates(sam, "ctg", ref, None) == rows
    c = EVC.Candidates("ctg", ref)
    c.feed(sam, final=True)
    text, pos = c.take()
    c.close()
    assert text.decode().split("\n")[:-1] == rows and len(rows) > 20
    want = O.create_tensors(sam, ref, pos.tolist())
    centers, x, _ = run_native(sam, ref, pos)
    assert np.array_equal(centers, g["centers"]) and np.array_equal(x, g["tensors"].astype(np.float32))
    assert [w[0] for w in want] == g["centers"].tolist() and all(np.array_equal(w[1], x[i]) for i, w in enumerate(want))


def test_sam_text_view_equals_samtools_view_semantics():
    """cvb_sam_view (what the command lines apply to .sam text instead of `samtools view -F 2308 file ctg:start-end`):
    header / contig / flag / region-overlap filter, any chunking, against a line-by-line Python statement of the rule"""
    import io
    import re
    rng = np.random.default_rng(11)
    ref, sam, _ = synth_alignments(rng, ref_len=3000, n_reads=500)
    rows = sam.split("\n")[:-1]
    out = []
    for i, r in enumerate(rows):
        f = r.split("\t")
        if not r.startswith("@"):
            f[1] = str([0, 16, 4, 256, 2048, 1024, 83, 2064][i % 8])
            if i % 17 == 0:
                f[2] = "ctg2"
        out.append("\t".join(f))
    out.insert(40, "")
    out.insert(90, "short\trow")
    text = "\n".join(out)                                   # (no trailing newline: the last record must still come through)

    def want(a, b):
        keep = []
        for line in text.split("\n"):
            if line.startswith("@") or not line.strip():
                continue
            f = line.split("\t")
            if len(f) < 6:
                keep.append(line)
                continue
            if int(f[1]) & 2308 or f[2] != "ctg":
                continue
            lo = int(f[3])
            hi = lo + max(sum(int(k) for k, op in re.findall(r"(\d+)([MIDNSHP=X])", f[5]) if op in "MDN=X"), 1) - 1
            if a is not None and (hi < a or lo > b):
                continue
            keep.append(line)
        return "".join(k + "\n" for k in keep).encode()

    for a, b in ((None, None), (1, 3000), (700, 1500), (2990, 5000), (5000, 6000)):
        for size in (1 << 20, 977, 64):
            v = CT._SamTextView(io.BytesIO(text.encode()), "ctg", a, b)
            got = b""
            while True:
                c = v.read(size)
                if not c:
                    break
                got += c
            assert got == want(a, b), (a, b, size)


@pytest.mark.parametrize("seed", [0, 5, 9])
@pytest.mark.parametrize("opts", [dict(), dict(dcov=2, minCoverage=3), dict(considerleftedge=False, minMQ=10)])
def test_thread_count_does_not_change_the_result(seed, opts):
    """the CIGAR walks of a feed call may run on several threads (cvb_pileup_set_threads): same tensors, same order, same
    statistics for every thread count -- on sorted input, on unsorted / damaged input, for whole and chunked feeds"""
    rng = np.random.default_rng(seed)
    ref, sam, _ = synth_alignments(rng, ref_len=6000, n_reads=1500, dup_pos=0.3)
    cands = sorted(set(int(c) for c in rng.integers(1, 6000, size=900)))          # dense: every ~7 bases
    rows = sam.split("\n")
    shuffled = rows[:2] + [rows[i] for i in rng.permutation(np.arange(2, len(rows) - 1))[:400]]
    for text, chunk in ((sam, None), (sam, 20000), ("\n".join(shuffled) + "\n", None)):
        want = None
        for threads in (1, 2, 3, 8):
            c, x, st = run_native(text, ref, cands, None, chunk, threads=threads, **opts)
            if want is None:
                want = (c, x, st)
                assert len(c) > 20
            else:
                assert np.array_equal(c, want[0]) and np.array_equal(x, want[1]) and st == want[2], (threads, chunk)


def test_sam_text_view_large_chunks_use_threads_and_agree_with_small_chunks():
    """chunks of several MB are filtered on more than one thread (cut at line starts, compacted afterwards): same bytes as
    the same text filtered in small, single-threaded chunks -- with and without a final newline, with a region"""
    import io
    rng = np.random.default_rng(12)
    ref, sam, _ = synth_alignments(rng, ref_len=3000, n_reads=800)
    rows = [r for r in sam.split("\n") if r]
    body = []
    for k in range(80):                                    # ~10 MB, flags / contigs varied
        for i, r in enumerate(rows):
            f = r.split("\t")
            if not r.startswith("@"):
                f[1] = str([0, 16, 4, 256, 0, 0, 2048, 0][(i + k) % 8])
                if (i + k) % 19 == 0:
                    f[2] = "ctg2"
            body.append("\t".join(f))
    for tail in ("\n", "", "\nr_last\t0\tctg\t100\t60\t50M\t*\t0\t0\t" + "A" * 50 + "\t*"):
        text = ("\n".join(body) + tail).encode()
        assert len(text) > 8 << 20
        for region in ((None, None), (500, 1800)):
            def filtered(size):
                v = CT._SamTextView(io.BytesIO(text), "ctg", *region)
                out = []
                while True:
                    c = v.read(size)
                    if not c:
                        return b"".join(out)
                    out.append(c)
            small, big = filtered(100000), filtered(1 << 30)
            assert small == big and len(small) > 1 << 20
