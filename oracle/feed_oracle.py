"""CPU restatement of the reference's text feed (clairvoyante/utils_v2.py:20-59) -- TEST INFRASTRUCTURE ONLY.

Pure-Python tokeniser following the reference row by row: `row.split()`, `np.array(strings, float32)`, drop rows whose
centre reference base is not ACGT (:39), `X[...,1:4] -= X[...,0:1]` (:46,57), batches of `num`, final partial batch with
endFlag 1.  The native parser (clairvoyante_b200/csrc/text_feed.cpp) is checked against this on the same bytes.
Pinned against the reference itself: tests/golden/reference_run.npz holds what the reference's own utils_v2.GetTensor
yielded for 630 rows when it was executed in the build container (generator and its stand-ins:
tests/golden/make_golden_reference_run.py; check: tests/test_reference_run_cpu.py).  One deliberate difference from the reference (shared with the product):
a malformed row is reported and skipped instead of silently re-using the previous row's fields (:34-41).
"""
import io
import shlex
import subprocess
import sys

import numpy as np

H, CENTRE, WIDTH = 33, 16, 528
_ACGT = frozenset("ACGT")


def _open_text(fn):
    proc = subprocess.Popen(shlex.split("gzip -fdc %s" % fn), stdout=subprocess.PIPE, bufsize=8388608)
    return proc, io.TextIOWrapper(proc.stdout, encoding="ascii", errors="replace")


def _sub(x):
    x[..., 1:] -= x[..., 0:1]
    return x


def GetTensor(tensor_fn, num):
    """Generator over batches of `num` candidate sites parsed from `chrom pos refseq33 v0..v527` rows
    (format: dataPrepScripts/CreateTensor.py:56).  Yields (0, num, X, pos) for full batches and finally
    (1, c, X[:c], pos) with 0 <= c < num (possibly empty), X float32 (c,33,4,4), pos = 'chrom:pos:seq'."""
    proc, fo = _open_text(tensor_fn)
    width, h, centre = WIDTH, H, CENTRE
    total = 0
    rows = np.empty((num, width), dtype=np.float32)
    pos, c = [], 0
    for row in fo:
        f = row.split()
        if len(f) != width + 3:
            if f:
                print("UnpackATensorRecord Failure", row, file=sys.stderr)
            continue
        seq = f[2].upper()
        if len(seq) <= centre:
            print("UnpackATensorRecord Failure", row, file=sys.stderr)
            continue
        if seq[centre] not in _ACGT:          # TODO in the reference too: IUPAC codes (utils_v2.py:39)
            continue
        try:
            rows[c] = np.array(f[3:], dtype=np.float32)
        except ValueError:
            print("UnpackATensorRecord Failure", row, file=sys.stderr)
            continue
        pos.append(f[0] + ":" + f[1] + ":" + seq)
        c += 1
        if c == num:
            x = _sub(rows.reshape(num, h, 4, 4))
            total += c
            print("Processed %d tensors" % total, file=sys.stderr)
            yield 0, c, x, pos
            rows = np.empty((num, width), dtype=np.float32)   # fresh storage: the consumer still holds x
            pos, c = [], 0
    fo.close()
    proc.wait()
    x = _sub(rows[:c].reshape(c, h, 4, 4))
    total += c
    print("Processed %d tensors" % total, file=sys.stderr)
    yield 1, c, x, pos
