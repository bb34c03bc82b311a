"""CreateTensor: alignments + candidate positions -> candidate tensors, on the native pile-up (csrc/pileup.cpp).

Python-3 counterpart of reference dataPrepScripts/CreateTensor.py with the same command line (:260-312) and the same output
rows `ctg pos refseq33 v0 .. v527` (:56), so that it drops into the reference's pipelines
(`ExtractVariantCandidates | CreateTensor | callVar`, callVarBam.py:61).  The per-base Python loops of the reference
(:148-252, ~10^3 sites/s) run in C++ (`cvb_pileup_*`, include/cvb200.h).

Inputs: like the reference, `samtools faidx` / `samtools view -F 2308` are used when `samtools` is on PATH; without it (or
when `--bam_fn` names a `.sam` / `.sam.gz` text file) the SAM text and the FASTA are read directly.

In-process use, skipping the text stream altogether:
    for end, n, X, pos in GetTensorFromAlignments(sam_chunks, ref_seq, candidates, ctgName, num): m.predict(X) ...
yields exactly what utils_v2.GetTensor yields for the rows this module would have printed.
"""
import argparse
import ctypes
import gzip
import os
import shlex
import shutil
import subprocess
import sys

import numpy as np

from . import _lib, param

_F = param.flankingBaseNum
_ACGT = frozenset("ACGT")


class Pileup(object):
    """thin handle over cvb_pileup_* (one contig / region)"""

    def __init__(self, ref_seq, candidates, ref_start=None, minMQ=0, dcov=250, minCoverage=0, considerleftedge=True, threads=None):
        """threads: host threads for the CIGAR walks of a feed() call (None = CVB_HOST_THREADS or up to 16 cores);
        the result does not depend on it"""
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        ref = ref_seq.encode("ascii", "replace") if isinstance(ref_seq, str) else bytes(ref_seq)
        cand = np.ascontiguousarray(candidates, dtype=np.int64)
        _lib.check(self._lib.cvb_pileup_create(ref, len(ref), int(ref_start or 0), cand.ctypes.data, cand.size,
                                               int(minMQ), int(dcov), int(minCoverage), 1 if considerleftedge else 0,
                                               ctypes.byref(self._h)))
        if threads is None:
            threads = int(os.environ.get("CVB_HOST_THREADS", 0)) or min(16, os.cpu_count() or 1)
        _lib.check(self._lib.cvb_pileup_set_threads(self._h, int(threads)))

    def feed(self, sam_bytes, final=False):
        b = sam_bytes.encode("ascii", "replace") if isinstance(sam_bytes, str) else sam_bytes
        if isinstance(b, np.ndarray):                     # (uint8 buffer from _SamTextView: no copy)
            _lib.check(self._lib.cvb_pileup_feed(self._h, b.ctypes.data, b.size, 1 if final else 0))
            return
        _lib.check(self._lib.cvb_pileup_feed(self._h, b, len(b), 1 if final else 0))

    def ready(self):
        return int(self._lib.cvb_pileup_ready(self._h))

    def take(self, max_sites=None):
        """(centers int64 (n,), raw count tensors float32 (n,33,4,4)) of the finished sites, in position order"""
        n = self.ready() if max_sites is None else min(int(max_sites), self.ready())
        x = np.empty((n, 2 * _F + 1, 4, param.matrixNum), np.float32)
        c = np.empty((n,), np.int64)
        got = ctypes.c_int64()
        _lib.check(self._lib.cvb_pileup_take(self._h, n, x.ctypes.data, c.ctypes.data, ctypes.byref(got)))
        return c[:got.value], x[:got.value]

    def stats(self):
        s = (ctypes.c_int64 * 4)()
        _lib.check(self._lib.cvb_pileup_stats(self._h, s))
        return dict(sam_rows=int(s[0]), rows_used=int(s[1]), malformed=int(s[2]), open_centres=int(s[3]))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvb_pileup_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _chunks(sam_source, size=8 << 20):
    """bytes chunks from a file object, a bytes/str blob or an iterable of chunks"""
    if isinstance(sam_source, (bytes, str)):
        yield sam_source
    elif hasattr(sam_source, "read"):
        read = getattr(sam_source, "read_array", None) or sam_source.read     # (_SamTextView: uint8 buffers, no copy)
        while True:
            b = read(size)
            if len(b) == 0:
                break
            yield b
    else:
        for b in sam_source:
            yield b


def pileup_tensors(sam_source, ref_seq, candidates, ref_start=None, batch=4096, **opts):
    """generator of (centers, raw tensors) batches over a position-sorted SAM stream"""
    p = Pileup(ref_seq, candidates, ref_start, **opts)
    try:
        for b in _chunks(sam_source):
            p.feed(b)
            while p.ready() >= batch:
                yield p.take(batch)
        p.feed(b"", final=True)
        while p.ready():
            yield p.take(batch)
    finally:
        p.close()


def _refseq33(ref_seq, ref_start, center):
    new_ref_pos = center - (0 if ref_start is None else ref_start - 1)
    return ref_seq[new_ref_pos - (_F + 1):new_ref_pos + _F]


def tensor_text(ctgName, centers, X, ref_seq, ref_start=None):
    """the text rows of CreateTensor.py:56 for one batch, as one bytes object ("...\\n" per row), formatted natively"""
    lib = _lib.load()
    n = len(centers)
    if n == 0:
        return b""
    ref = ref_seq.encode("ascii", "replace") if isinstance(ref_seq, str) else ref_seq
    c = np.ascontiguousarray(centers, dtype=np.int64)
    x = np.ascontiguousarray(X, dtype=np.float32)
    for per_value in (7, 16):               # "250.0 " is 6 bytes; the library refuses a row that might not fit
        cap = n * (len(ctgName) + 64 + 528 * per_value) + 528 * 16
        buf = np.empty(cap, np.uint8)       # (no zero fill)
        got = lib.cvb_pileup_format_rows(ctgName.encode(), c.ctypes.data, x.ctypes.data, n, ref, len(ref), int(ref_start or 0),
                                         buf.ctypes.data, cap)
        if got >= 0:
            return buf[:got].tobytes()
    _lib.check(1)


def tensor_rows(ctgName, centers, X, ref_seq, ref_start=None):
    """the same rows as a list of str (without the newline)"""
    t = tensor_text(ctgName, centers, X, ref_seq, ref_start)
    return t.decode("ascii").split("\n")[:-1] if t else []


def GetTensorFromAlignments(sam_source, ref_seq, candidates, ctgName, num, ref_start=None, **opts):
    """Same yields as utils_v2.GetTensor(tensor_fn, num) -- (endFlag, n, X, pos) with X channel-subtracted (utils_v2.py:46),
    rows whose centre reference base is not ACGT dropped (:39), pos = 'ctg:pos:SEQ' -- without the text stream."""
    h = 2 * _F + 1
    X = np.empty((num, h, 4, param.matrixNum), np.float32)
    pos, c = [], 0
    for centers, T in pileup_tensors(sam_source, ref_seq, candidates, ref_start, batch=max(num, 1), **opts):
        seqs = [_refseq33(ref_seq, ref_start, ctr).upper() for ctr in centers.tolist()]
        keep = [i for i, q in enumerate(seqs) if len(q) > _F and q[_F] in _ACGT]
        i0 = 0
        while i0 < len(keep):                       # fill the current batch with whole slices of this pile-up batch
            k = min(num - c, len(keep) - i0)
            idx = keep[i0:i0 + k]
            X[c:c + k] = T[idx]
            pos += ["%s:%d:%s" % (ctgName, int(centers[i]), seqs[i]) for i in idx]
            c += k
            i0 += k
            if c == num:
                yield 0, c, _subtracted_with_counts(X), pos
                X = np.empty((num, h, 4, param.matrixNum), np.float32)
                pos, c = [], 0
    yield 1, c, _subtracted_with_counts(X[:c]), pos


def _subtracted_with_counts(X):
    """raw pile-up counts -> the batch GetTensor would yield (channels 1..3 relative to channel 0, utils_v2.py:46), as a
    CountBatch that keeps the raw counts as uint8 / int16 for the narrow host->device feed (cvb_predict_host_counts_*)"""
    from . import utils_v2
    narrow = os.environ.get("CVB_FEED", "counts") != "fp32"
    counts = utils_v2.pack_counts(X, subtracted=False) if narrow and len(X) else None
    X[..., 1:] -= X[..., 0:1]
    if counts is None:
        return X
    out = X.view(utils_v2.CountBatch)
    out.counts = counts
    return out


# ------------------------------------------------------------------------------------------------
# command line (reference CreateTensor.py:96-146, :260-312)
# ------------------------------------------------------------------------------------------------
def _read_fasta_contig(ref_fn, ctgName, start=None, end=None):
    """plain FASTA reader used when samtools is absent: bases [start, end] (1-based, inclusive) of contig ctgName"""
    from .utils_v2 import open_maybe_gzip
    seq, on = [], False
    with open_maybe_gzip(ref_fn) as f:
        for line in f:
            if line.startswith(">"):
                if on:
                    break
                on = line[1:].split()[0] == ctgName if line[1:].split() else False
            elif on:
                seq.append(line.strip())
    s = "".join(seq)
    if start is not None and end is not None:
        s = s[start - 1:end]
    return s


def _load_reference(args):
    """:103-128 -- returns (refSeq, refStart or None); ctgStart/ctgEnd are converted to 1-based like the reference"""
    if args.ctgStart is not None and args.ctgEnd is not None:
        args.ctgStart += 1
        ref_start = max(1, args.ctgStart - param.expandReferenceRegion)
        ref_end = args.ctgEnd + param.expandReferenceRegion
        region = "%s:%d-%d" % (args.ctgName, ref_start, ref_end)
    else:
        args.ctgStart = args.ctgEnd = None
        ref_start, ref_end, region = None, None, args.ctgName
    if shutil.which(args.samtools):
        p = subprocess.Popen(shlex.split("%s faidx %s %s" % (args.samtools, args.ref_fn, region)), stdout=subprocess.PIPE,
                             bufsize=8388608)
        rows = p.stdout.read().decode("ascii", "replace").split("\n")
        p.stdout.close()
        p.wait()
        seq = "".join(r.rstrip() for r in rows[1:]) if p.returncode == 0 else ""
    else:
        seq = _read_fasta_contig(args.ref_fn, args.ctgName, ref_start, ref_end)
    if len(seq) == 0:
        sys.exit("Failed to load reference seqeunce. Please check if the provided reference fasta %s and the ctgName %s "
                 "are correct." % (args.ref_fn, args.ctgName))
    return seq, ref_start


def _load_candidates(args):
    """GetCandidate (:61-83): positions of this contig inside [ctgStart, ctgEnd]"""
    if args.can_fn == "PIPE":
        fo = sys.stdin
    else:                      # gzip or plain by content: the reference reads with `gzip -fdc` and its own stages write
        from .utils_v2 import open_maybe_gzip    # gzip under suffix-less names (ExtractVariantCandidates.py, GetTruth.py)
        fo = open_maybe_gzip(args.can_fn)
    out = []
    for row in fo:
        row = row.split()
        if len(row) < 2 or args.ctgName != row[0]:
            continue
        pos = int(row[1])
        if args.ctgStart is not None and pos < args.ctgStart:
            continue
        if args.ctgEnd is not None and pos > args.ctgEnd:
            continue
        out.append(pos)
    if fo is not sys.stdin:
        fo.close()
    return out


class _SamTextView(object):
    """What `samtools view -F 2308 <file> ctg[:start-end]` (reference CreateTensor.py:134-136) prints, for a SAM TEXT file:
    header lines are dropped, records of other contigs, unmapped / secondary / supplementary records (flag & 2308) and
    records that do not overlap the 1-based inclusive region are skipped (native: cvb_sam_view, csrc/sam_view.cpp).
    File-like: read() hands out filtered chunks."""

    def __init__(self, fh, ctgName, start=None, end=None):
        self.fh, self.ctg, self.carry = fh, ctgName.encode(), b""
        self.start, self.end = (-1, -1) if start is None or end is None else (int(start), int(end))
        self.lib = _lib.load()

    def read(self, size=8 << 20):
        return self.read_array(size).tobytes()

    def read_array(self, size=8 << 20):
        """like read(), but hands out the uint8 buffer the filter wrote (no copy; empty at the end of the stream)"""
        while True:
            b = self.fh.read(size)
            final = not b
            data = self.carry + b if self.carry else b
            if final and not data:
                return np.empty(0, np.uint8)
            out = np.empty(len(data) + 1, np.uint8)             # (no zero fill, unlike ctypes.create_string_buffer)
            n, used = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(self.lib.cvb_sam_view(data, len(data), 1 if final else 0, self.ctg, 2308, self.start, self.end,
                                             out.ctypes.data, ctypes.byref(n), ctypes.byref(used)))
            self.carry = data[used.value:]
            if n.value or final:
                return out[:n.value]

    def close(self):
        self.fh.close()


def _open_alignments(args):
    fn = args.bam_fn
    if fn.endswith(".sam") or fn.endswith(".sam.gz"):
        from .utils_v2 import open_maybe_gzip
        fh = open_maybe_gzip(fn, "rb")
        return None, _SamTextView(fh, args.ctgName, args.ctgStart, args.ctgEnd if args.ctgStart is not None else None)
    if not shutil.which(args.samtools):
        sys.exit("samtools not found: pass a .sam / .sam.gz text file as --bam_fn or install samtools")
    region = "%s:%d-%d" % (args.ctgName, args.ctgStart, args.ctgEnd) if args.ctgStart is not None else args.ctgName
    p = subprocess.Popen(shlex.split("%s view -F 2308 %s %s" % (args.samtools, args.bam_fn, region)), stdout=subprocess.PIPE,
                         bufsize=8388608)
    return p, p.stdout


def OutputAlnTensor(args):
    ref_seq, ref_start = _load_reference(args)
    cands = _load_candidates(args)
    proc, sam = _open_alignments(args)
    if args.tensor_fn != "PIPE":
        out = gzip.open(args.tensor_fn, "wb", compresslevel=6)
    else:
        out = sys.stdout.buffer
    for centers, X in pileup_tensors(sam, ref_seq, cands, ref_start, minMQ=args.minMQ, dcov=args.dcov,
                                     minCoverage=args.minCoverage, considerleftedge=args.considerleftedge):
        out.write(tensor_text(args.ctgName, centers, X, ref_seq, ref_start))
    sam.close()
    if proc is not None:
        proc.wait()
    if out is not sys.stdout.buffer:
        out.close()
    else:
        out.flush()


def main():
    parser = argparse.ArgumentParser(
        description="Generate tensors summarizing local alignments from a BAM file and a list of candidate locations")
    parser.add_argument('--bam_fn', type=str, default="input.bam", help="Sorted bam file input (or .sam / .sam.gz text), default: %(default)s")
    parser.add_argument('--ref_fn', type=str, default="ref.fa", help="Reference fasta file input, default: %(default)s")
    parser.add_argument('--can_fn', type=str, default="PIPE",
                        help="Variant candidate list generated by ExtractVariantCandidates.py or true variant list generated by "
                             "GetTruth.py, use PIPE for standard input, default: %(default)s")
    parser.add_argument('--tensor_fn', type=str, default="PIPE", help="Tensor output, use PIPE for standard output, default: %(default)s")
    parser.add_argument('--minMQ', type=int, default=0, help="Minimum Mapping Quality, default: %(default)d")
    parser.add_argument('--ctgName', type=str, default="chr17", help="The name of sequence to be processed, default: %(default)s")
    parser.add_argument('--ctgStart', type=int, default=None, help="The 1-bsae starting position of the sequence to be processed")
    parser.add_argument('--ctgEnd', type=int, default=None, help="The inclusive ending position of the sequence to be processed")
    parser.add_argument('--samtools', type=str, default="samtools", help="Path to the 'samtools', default: %(default)s")
    parser.add_argument('--considerleftedge', type=param.str2bool, nargs='?', const=True, default=True,
                        help="Count the left-most base-pairs of a read for coverage even if the starting position of a read is "
                             "after the starting position of a tensor, default: %(default)s")
    parser.add_argument('--dcov', type=int, default=250, help="Cap depth per position at %(default)d")
    parser.add_argument('--minCoverage', type=int, default=0, help="Minimum coverage required to generate a tensor, default: %(default)d")
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    OutputAlnTensor(args)


if __name__ == "__main__":
    main()
