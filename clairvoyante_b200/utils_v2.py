"""Batch feed: text tensor stream -> batches, truth/labels -> training blocks, block slicing.

Python-3 counterpart of reference clairvoyante/utils_v2.py with the same public names and the same
yield / return contracts:

  GetTensor(tensor_fn, num)                       utils_v2.py:23-59   -> yields (endFlag, count, X, pos)
  GetTrainingArray(tensor_fn, var_fn, bed_fn)     utils_v2.py:62-186  -> (total, Xblocks, Yblocks, posBlocks)
  DecompressArray(blocks, start, num, maximum)    utils_v2.py:189-207 -> (array, num, endFlag)

Differences that are deliberate and documented:
  * python-blosc and intervaltree (requirements.txt:3-4) are not available; blocks are packed with
    zlib + np.save (`pack_array` / `unpack_array`) and BED membership uses sorted arrays + bisect with the
    same half-open [begin, end-1) semantics the reference builds (utils_v2.py:71-74).  A .bin written by the
    reference (blosc frames) is rejected with a clear error; reading it is a "next" row.
  * a malformed row is reported and SKIPPED; the reference prints the failure and then re-uses the previous
    row's fields (utils_v2.py:34-41), silently duplicating a record.
"""
import bisect
import gc
import io
import random
import shlex
import subprocess
import sys
import zlib

import numpy as np

from . import param

base2num = dict(zip("ACGT", (0, 1, 2, 3)))
_ACGT = frozenset("ACGT")
_MAGIC = b"CVBZ1"


def SetupEnv():
    """utils_v2.py:14-18 (blosc threads / TF log level have no counterpart here)"""
    gc.enable()


def _site_floats():
    return (2 * param.flankingBaseNum + 1) * 4 * param.matrixNum


def _open_text(fn):
    """`gzip -fdc fn` like the reference (utils_v2.py:24-28): transparently reads plain or gzip text; 'PIPE' = stdin"""
    if fn == "PIPE":
        return None, sys.stdin
    proc = subprocess.Popen(shlex.split("gzip -fdc %s" % fn), stdout=subprocess.PIPE, bufsize=8388608)
    return proc, io.TextIOWrapper(proc.stdout, encoding="ascii", errors="replace")


def _close_text(proc, fh):
    if proc is not None:
        fh.close()
        proc.wait()


def UnpackATensorRecord(a, b, c, *d):
    """utils_v2.py:20-21"""
    return a, b, c, np.array(d, dtype=np.float32)


def _subtract_reference_channel(x):
    """utils_v2.py:46 -- channels 1..3 hold counts relative to the reference channel 0"""
    x[..., 1:] -= x[..., 0:1]
    return x


def GetTensor(tensor_fn, num):
    """Generator over batches of `num` candidate sites parsed from `chrom pos refseq33 v0..v527` rows
    (format: dataPrepScripts/CreateTensor.py:56).  Yields (0, num, X, pos) for full batches and finally
    (1, c, X[:c], pos) with 0 <= c < num (possibly empty), X float32 (c,33,4,4), pos = 'chrom:pos:seq'."""
    proc, fo = _open_text(tensor_fn)
    width = _site_floats()
    h, centre = 2 * param.flankingBaseNum + 1, param.flankingBaseNum
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
            x = _subtract_reference_channel(rows.reshape(num, h, 4, param.matrixNum))
            total += c
            print("Processed %d tensors" % total, file=sys.stderr)
            yield 0, c, x, pos
            rows = np.empty((num, width), dtype=np.float32)   # fresh storage: the consumer still holds x
            pos, c = [], 0
    _close_text(proc, fo)
    x = _subtract_reference_channel(rows[:c].reshape(c, h, 4, param.matrixNum))
    total += c
    print("Processed %d tensors" % total, file=sys.stderr)
    yield 1, c, x, pos


# ------------------------------------------------------------------------------------------------
# block container (stands in for blosc.pack_array / unpack_array, utils_v2.py:174-176,198)
# ------------------------------------------------------------------------------------------------
def pack_array(a):
    buf = io.BytesIO()
    np.save(buf, np.asarray(a), allow_pickle=False)
    return _MAGIC + zlib.compress(buf.getvalue(), 1)


def unpack_array(b):
    if not isinstance(b, (bytes, bytearray)) or bytes(b[:5]) != _MAGIC:
        raise ValueError("not a clairvoyante_b200 block (a reference .bin holds python-blosc frames; "
                         "re-create it with this repo's tensor2Bin.py)")
    return np.load(io.BytesIO(zlib.decompress(bytes(b[5:]))), allow_pickle=False)


class _Regions(object):
    """per-contig interval membership with intervaltree's half-open semantics"""

    def __init__(self):
        self._raw = {}
        self._idx = {}

    def add(self, name, begin, end):
        self._raw.setdefault(name, []).append((begin, end))
        self._idx.pop(name, None)

    def __contains__(self, name):
        return name in self._raw

    def hit(self, name, p):
        if name not in self._raw:
            raise KeyError(name)
        if name not in self._idx:
            iv = sorted(self._raw[name])
            begins = [b for b, _ in iv]
            run, m = [], -1
            for _, e in iv:
                m = max(m, e)
                run.append(m)
            self._idx[name] = (begins, run)
        begins, run = self._idx[name]
        i = bisect.bisect_right(begins, p)
        return i > 0 and run[i - 1] > p


def _truth_label(ref, alt, gt1, gt2):
    """16-vector [A C G T | HET HOM | REF SNP INS DEL | len 0 1 2 3 4 >4] from one truth row (utils_v2.py:84-119)"""
    v = [0.0] * 16
    single = len(ref) == 1 and len(alt) == 1
    if gt1 == "0" and gt2 == "1":
        v[base2num[ref[0]]] = 0.5
        if single:
            v[base2num[alt[0]]] = 0.5
        v[4] = 1.0
    elif gt1 == "1" and gt2 == "1":
        if single:
            v[base2num[alt[0]]] = 1
        v[5] = 1.0
    if len(ref) > 1 and len(alt) == 1:
        v[9] = 1.0
    elif len(alt) > 1 and len(ref) == 1:
        v[8] = 1.0
    else:
        v[7] = 1.0
    d = abs(len(ref) - len(alt))
    v[15 if d > 4 else 10 + d] = 1.0
    return v


def GetTrainingArray(tensor_fn, var_fn, bed_fn, shuffle=True):
    regions = None
    if bed_fn is not None:
        regions = _Regions()
        proc, fh = _open_text(bed_fn)
        for row in fh:
            f = row.split()
            begin, end = int(f[1]), int(f[2]) - 1
            if end == begin:
                end += 1
            regions.add(f[0], begin, end)
        _close_text(proc, fh)

    Y = {}
    if var_fn is not None:
        proc, fh = _open_text(var_fn)
        for row in fh:
            f = row.split()
            if regions is not None and not regions.hit(f[0], int(f[1])):
                continue
            Y[f[0] + ":" + f[1]] = _truth_label(f[2], f[3], f[4], f[5])
        _close_text(proc, fh)

    X = {}
    h, centre = 2 * param.flankingBaseNum + 1, param.flankingBaseNum
    width = _site_floats()
    proc, fh = _open_text(tensor_fn)
    total = 0
    for row in fh:
        f = row.split()
        if len(f) != width + 3:
            continue
        chrom, coord, seq = f[0], f[1], f[2].upper()
        if regions is not None and (chrom not in regions or not regions.hit(chrom, int(coord))):
            continue
        if seq[centre] not in _ACGT:
            continue
        key = chrom + ":" + coord
        X[key] = _subtract_reference_channel(np.array(f[3:], dtype=np.float32).reshape(h, 4, param.matrixNum))
        if key not in Y:                      # non-variant default label (utils_v2.py:141-148)
            v = [0.0] * 16
            v[base2num[seq[centre]]] = 1.0
            v[5] = v[6] = v[10] = 1.0
            Y[key] = v
        total += 1
        if total % 100000 == 0:
            print("Processed %d tensors" % total, file=sys.stderr)
    _close_text(proc, fh)

    keys = sorted(X.keys())
    if shuffle:
        random.shuffle(keys)
    xb, yb, pb = [], [], []
    step = param.bloscBlockSize
    for i in range(0, len(keys), step):
        chunk = keys[i:i + step]
        xb.append(pack_array(np.array([X[k] for k in chunk], dtype=np.float32)))
        yb.append(pack_array(np.array([Y[k] for k in chunk], dtype=np.float64)))   # labels are float64 upstream (:175)
        pb.append(pack_array(np.array(chunk, dtype="S")))
    if len(keys) % step == 0:                 # the reference always appends a (possibly empty) trailing block (:181-184)
        xb.append(pack_array(np.zeros((0, h, 4, param.matrixNum), np.float32)))
        yb.append(pack_array(np.zeros((0, 16), np.float64)))
        pb.append(pack_array(np.array([], dtype="S1")))
    return len(keys), xb, yb, pb


def DecompressArray(array, start, num, maximum):
    """rows [start, start+num) across bloscBlockSize-row blocks; clamps at `maximum` (utils_v2.py:189-207)"""
    endFlag = 0
    if start + num >= maximum:
        num = maximum - start
        endFlag = 1
    bs = param.bloscBlockSize
    first, last = start // bs, (start + num - 1) // bs
    parts = [unpack_array(array[first])]
    for i in range(first + 1, last + 1):
        parts.append(unpack_array(array[i]))
    out = np.concatenate(parts)
    left = start % bs
    if left != 0 or num % bs != 0:
        out = out[left:left + num]
    return out, num, endFlag
