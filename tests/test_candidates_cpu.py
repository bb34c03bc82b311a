"""CPU: the native candidate extractor (csrc/candidates.cpp, replaces dataPrepScripts/ExtractVariantCandidates.py) against
the statement-by-statement restatement oracle/candidates_oracle.py on synthetic SAM text: identical output rows in identical
order, over the reference's options (region, BED, coverage / frequency thresholds, MAPQ, training subsample) and edge cases
(soft-clipped reads, insertions / deletions booked left of a read's start, lower-case reference, other contigs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from clairvoyante_b200 import ExtractVariantCandidates as EVC, CreateTensor as CT
from oracle import candidates_oracle as O

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_pileup_cpu import synth_alignments   # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_native(sam, ctg, ref, ref_start=None, chunk=None, **opts):
    c = EVC.Candidates(ctg, ref, ref_start, **opts)
    b = sam.encode()
    rows, pos = [], []
    step = chunk or len(b) or 1
    for i in range(0, len(b), step):
        c.feed(b[i:i + step])
        t, p = c.take()
        rows += t.decode().split("\n")[:-1]
        pos += p.tolist()
    c.feed(b"", final=True)
    t, p = c.take()
    rows += t.decode().split("\n")[:-1]
    pos += p.tolist()
    st = c.stats()
    c.close()
    return rows, pos, st


def with_extras(sam, rng):
    """reads of another contig, heavily soft-clipped reads, indels at the very start of a read"""
    rows = sam.rstrip("\n").split("\n")
    out = []
    for r in rows:
        out.append(r)
        if r.startswith("@"):
            continue
        f = r.split("\t")
        u = rng.random()
        if u < 0.03:
            out.append("\t".join([f[0] + "x", f[1], "other"] + f[3:]))
        elif u < 0.06:
            out.append("\t".join(f[:5] + ["60S20M", "*", "0", "0", "A" * 80, "*"]))
        elif u < 0.10:
            out.append("\t".join(f[:5] + ["3I20M2D10M", "*", "0", "0", "ACG" + "T" * 30, "*"]))
        elif u < 0.13:
            out.append("\t".join(f[:5] + ["2D25M", "*", "0", "0", "C" * 25, "*"]))
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("opts", [dict(), dict(minMQ=25), dict(minCoverage=8), dict(threshold=0.3), dict(minCoverage=0, threshold=0),
                                  dict(ctgStart=501, ctgEnd=1500), dict(bed=[(100, 400), (390, 395), (900, 901), (2000, 2600)]),
                                  dict(ctgStart=301, ctgEnd=2500, bed=[(0, 350), (1000, 1200)]),
                                  dict(minCoverage=0, threshold=0, outputProb=0.3, seed=12345)])
def test_native_candidates_equal_restatement(seed, opts):
    rng = np.random.default_rng(seed)
    ref, sam, _ = synth_alignments(rng, n_reads=500)
    sam = with_extras(sam, rng)
    okw = dict(opts)
    if "outputProb" in okw:
        okw["output_prob"] = okw.pop("outputProb")
    want = O.make_candidates(sam, "ctg", ref, None, **okw)
    rows, pos, st = run_native(sam, "ctg", ref, None, **opts)
    assert rows == want
    assert pos == [int(r.split()[1]) for r in want]
    assert len(want) > (2 if opts else 50)
    assert st["open_positions"] == 0 and st["malformed"] == 0


def test_ties_follow_the_reference_dict_order_and_lowercase_reference():
    # position 1 (0-based): A x2, C x2 -> tie: A first (dict order A C D G I N T), reference base C -> candidate because
    # the top key differs; lower-case reference base never equals an upper-case key -> every covered position is output
    ref = "ACgTACGTAC" * 5
    sam = "".join("r%d\t0\tctg\t2\t60\t3M\t*\t0\t0\t%s\t*\n" % (i, s) for i, s in enumerate(["AGT", "AGT", "CGT", "CGT"]))
    want = O.make_candidates(sam, "ctg", ref, None, minCoverage=0)
    rows, _, _ = run_native(sam, "ctg", ref, None, minCoverage=0)
    assert rows == want
    assert rows[0].startswith("ctg 2 C 4 A 2 C 2 D 0 G 0 I 0 N 0 T 0")
    assert rows[1].split()[:4] == ["ctg", "3", "g", "4"]


def test_chunked_feed_region_offset_and_empty():
    rng = np.random.default_rng(4)
    ref, sam, _ = synth_alignments(rng, n_reads=300)
    sam = with_extras(sam, rng) + "bad\trow\n"
    ref_start = 21
    sub = ref[ref_start - 1:]
    want = O.make_candidates(sam, "ctg", sub, ref_start)
    for chunk in (1, 11, 1000, None):
        rows, _, st = run_native(sam, "ctg", sub, ref_start, chunk=chunk)
        assert rows == want and st["malformed"] == 1
    rows, pos, st = run_native("", "ctg", ref)
    assert rows == [] and pos == [] and st["reads_processed"] == 0


def test_pipeline_candidates_into_pileup_and_command_line(tmp_path):
    """the two native stages chained like callVarBam.py:56-66, and the CLI on .sam + FASTA (+ BED)"""
    rng = np.random.default_rng(6)
    ref, sam, _ = synth_alignments(rng, n_reads=400)
    rows, pos, _ = run_native(sam, "ctg", ref)
    assert len(pos) > 50
    centers = np.concatenate([c for c, _ in CT.pileup_tensors(sam, ref, pos)])
    assert set(centers.tolist()) <= set(pos) and len(centers) > 50
    (tmp_path / "a.sam").write_text(sam)
    (tmp_path / "ref.fa").write_text(">ctg\n" + "\n".join(ref[i:i + 70] for i in range(0, len(ref), 70)) + "\n")
    (tmp_path / "r.bed").write_text("ctg\t200\t1500\nother\t1\t5\n")
    cmd = [sys.executable, "-m", "clairvoyante_b200.ExtractVariantCandidates", "--bam_fn", str(tmp_path / "a.sam"), "--ref_fn",
           str(tmp_path / "ref.fa"), "--ctgName", "ctg", "--samtools", "/nonexistent/samtools"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split("\n")[:-1] == rows
    r = subprocess.run(cmd + ["--bed_fn", str(tmp_path / "r.bed"), "--ctgStart", "300", "--ctgEnd", "2000"], cwd=ROOT,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    # the reference fetches [ctgStart+1-1e6, ctgEnd+1e6] -> here the whole contig from position 1
    want = O.make_candidates(sam, "ctg", ref, 1, ctgStart=301, ctgEnd=2000, bed=[(200, 1499)])
    assert r.stdout.split("\n")[:-1] == want


def test_mutated_sam_rows_never_crash_and_still_agree():
    """junk CIGAR ops, '*' CIGARs, SEQ shorter than the CIGAR claims, very long deletions, truncated rows, N/H/P ops: both native
    stages keep agreeing with their restatements (which define those cases: the regular expression skips junk, short SEQ
    reads as N, short rows are dropped)"""
    from oracle import createtensor_oracle as OT
    rng = np.random.default_rng(123)
    for it in range(40):
        ref, sam, cands = synth_alignments(rng, ref_len=int(rng.integers(260, 1200)), n_reads=int(rng.integers(1, 60)), read_len=(5, 90))
        out = []
        for r in sam.split("\n"):
            f, u = r.split("\t"), rng.random()
            if len(f) >= 10:
                if u < 0.10: f[5] = f[5] + "7Q3M9"
                elif u < 0.15: f[5] = "*"
                elif u < 0.20: f[9] = f[9][:max(1, len(f[9]) // 2)]
                elif u < 0.25: f[5] = "30000D5M"
                elif u < 0.30: f = f[:int(rng.integers(1, 10))]
                elif u < 0.33: f[5] = "5M2N5M3H2P4="
            out.append("\t".join(f))
        sam2 = "\n".join(out)
        want = OT.create_tensors(sam2, ref, cands)
        p = CT.Pileup(ref, cands)
        p.feed(sam2, final=True)
        cc, x = p.take()
        p.close()
        assert [int(v) for v in cc] == [w[0] for w in want] and all(np.array_equal(x[i], w[1]) for i, w in enumerate(want)), it
        rows, _, _ = run_native(sam2, "ctg", ref, None, minCoverage=0)
        assert rows == O.make_candidates(sam2, "ctg", ref, None, minCoverage=0), it


@pytest.mark.parametrize("seed", [1, 4, 8])
@pytest.mark.parametrize("opts", [dict(), dict(minMQ=20, threshold=0.05, minCoverage=2),
                                  dict(ctgStart=801, ctgEnd=5000, bed=[(500, 2500), (3000, 3001), (3500, 9000)])])
def test_thread_count_does_not_change_the_result(seed, opts):
    """several host threads per feed call (cvb_candidates_set_threads: rows tokenised in parallel, counting split by position):
    same rows, same order, same statistics as the serial loop -- sorted input, shuffled input, whole and chunked feeds"""
    rng = np.random.default_rng(seed)
    ref, sam, _ = synth_alignments(rng, ref_len=7000, n_reads=3000, dup_pos=0.3)
    sam = with_extras(sam, rng)
    rows = sam.split("\n")
    head = [r for r in rows if r.startswith("@")]
    body = [r for r in rows if r and not r.startswith("@")]
    shuffled = "\n".join(head + [body[i] for i in rng.permutation(len(body))[:1500]]) + "\n"
    for text, chunk in ((sam, None), (sam, 300000), (shuffled, None)):
        want = None
        for threads in (1, 2, 3, 8):
            got = run_native(text, "ctg", ref, None, chunk, threads=threads, **opts)
            if want is None:
                want = got
                assert len(got[0]) > 50
            else:
                assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2], (threads, chunk)
