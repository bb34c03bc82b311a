"""ExtractVariantCandidates: alignments -> variant-candidate positions, on the native extractor (csrc/candidates.cpp).

Python-3 counterpart of reference dataPrepScripts/ExtractVariantCandidates.py with the same command line (:256-312) and the
same output rows `ctg pos refBase total k0 n0 .. k6 n6` (:35-39), first stage of the reference's calling pipeline
(`ExtractVariantCandidates | CreateTensor | callVar`, callVarBam.py:56-66).  The per-base Python loop (:127-243) runs in C++
(`cvb_candidates_*`, include/cvb200.h).  `samtools` is used for BAM / FASTA access when installed; `.sam` / `.sam.gz` text
and plain FASTA are read directly otherwise.
"""
import argparse
import ctypes
import gzip
import os
import sys

import numpy as np

from . import _lib, param
from .CreateTensor import _chunks, _load_reference, _open_alignments


class Candidates(object):
    """thin handle over cvb_candidates_* (one contig / region)"""

    def __init__(self, ctgName, ref_seq, ref_start=None, ctgStart=None, ctgEnd=None, bed=None, minMQ=0, minCoverage=4,
                 threshold=0.125, outputProb=None, seed=0, threads=None):
        """threads: host threads per feed() call (None = CVB_HOST_THREADS or up to 16 cores); the result does not depend on it"""
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        ref = ref_seq.encode("ascii", "replace") if isinstance(ref_seq, str) else bytes(ref_seq)
        if bed is None:
            bb = be = None
            nb = -1
        else:
            bb = np.ascontiguousarray([b for b, _ in bed], dtype=np.int64)
            be = np.ascontiguousarray([e for _, e in bed], dtype=np.int64)
            nb = len(bed)
        region = ctgStart is not None and ctgEnd is not None
        _lib.check(self._lib.cvb_candidates_create(
            ctgName.encode(), ref, len(ref), int(ref_start or 0), int(ctgStart) if region else -1, int(ctgEnd) if region else -1,
            bb.ctypes.data if nb > 0 else None, be.ctypes.data if nb > 0 else None, nb, int(minMQ), float(minCoverage),
            float(threshold), -1.0 if outputProb is None else float(outputProb), int(seed) & 0xFFFFFFFFFFFFFFFF,
            ctypes.byref(self._h)))
        if threads is None:
            threads = int(os.environ.get("CVB_HOST_THREADS", 0)) or min(16, os.cpu_count() or 1)
        _lib.check(self._lib.cvb_candidates_set_threads(self._h, int(threads)))

    def feed(self, sam_bytes, final=False):
        b = sam_bytes.encode("ascii", "replace") if isinstance(sam_bytes, str) else sam_bytes
        if isinstance(b, np.ndarray):                     # (uint8 buffer from _SamTextView: no copy)
            _lib.check(self._lib.cvb_candidates_feed(self._h, b.ctypes.data, b.size, 1 if final else 0))
            return
        _lib.check(self._lib.cvb_candidates_feed(self._h, b, len(b), 1 if final else 0))

    def take(self):
        """(rows as bytes, "...\\n" each; positions int64, 1-based) finished so far"""
        nb = int(self._lib.cvb_candidates_pending_bytes(self._h))
        npos = int(self._lib.cvb_candidates_pending(self._h))
        text = np.empty(max(nb, 1), np.uint8)
        pos = np.empty(max(npos, 1), np.int64)
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self._lib.cvb_candidates_take(self._h, text.ctypes.data, nb, ctypes.byref(a), pos.ctypes.data, npos,
                                                 ctypes.byref(b)))
        return text[:a.value].tobytes(), pos[:b.value].copy()

    def stats(self):
        s = (ctypes.c_int64 * 4)()
        _lib.check(self._lib.cvb_candidates_stats(self._h, s))
        return dict(sam_rows=int(s[0]), reads_processed=int(s[1]), malformed=int(s[2]), open_positions=int(s[3]))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvb_candidates_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def extract_candidates(sam_source, ctgName, ref_seq, ref_start=None, **opts):
    """generator of (rows bytes, positions) over a position-sorted SAM stream; the last item carries the stats as .stats"""
    c = Candidates(ctgName, ref_seq, ref_start, **opts)
    try:
        for b in _chunks(sam_source):
            c.feed(b)
            t, p = c.take()
            if len(p):
                yield t, p
        c.feed(b"", final=True)
        t, p = c.take()
        extract_candidates.last_stats = c.stats()
        if len(p):
            yield t, p
    finally:
        c.close()


def _load_bed(bed_fn, ctgName):
    """:89-105 -- half-open intervals of this contig as the reference builds them"""
    from .utils_v2 import open_maybe_gzip
    out, seen = [], set()
    with open_maybe_gzip(bed_fn) as f:
        for row in f:
            row = row.strip().split()
            if len(row) < 3:
                continue
            seen.add(row[0])
            if row[0] != ctgName:
                continue
            begin, end = int(row[1]), int(row[2]) - 1
            if end == begin:
                end += 1
            if end > begin:
                out.append((begin, end))
    if ctgName not in seen:
        sys.exit("ctgName is not in the bed file, are you using the correct bed file (%s)?" % bed_fn)
    return out


def MakeCandidates(args):
    if args.gen4Training:
        args.minCoverage = 0
        args.threshold = 0
        args.outputProb = (args.candidates * 2.) / args.genomeSize
    else:
        args.outputProb = None
    ref_seq, ref_start = _load_reference(args)              # converts ctgStart to the reference's 1-based value (:63)
    bed = _load_bed(args.bed_fn, args.ctgName) if args.bed_fn is not None else None
    proc, sam = _open_alignments(args)
    if args.can_fn != "PIPE":
        out = gzip.open(args.can_fn, "wb", compresslevel=6)
    else:
        out = sys.stdout.buffer
    for text, _ in extract_candidates(sam, args.ctgName, ref_seq, ref_start, ctgStart=args.ctgStart, ctgEnd=args.ctgEnd, bed=bed,
                                      minMQ=args.minMQ, minCoverage=args.minCoverage, threshold=args.threshold,
                                      outputProb=args.outputProb, seed=int.from_bytes(os.urandom(8), "little")):
        out.write(text)
    sam.close()
    if proc is not None:
        proc.wait()
    if out is not sys.stdout.buffer:
        out.close()
    else:
        out.flush()
    if extract_candidates.last_stats["reads_processed"] == 0:
        print("No read has been process, either the genome region you specified has no read cover, or please check the "
              "correctness of your BAM input (%s)." % args.bam_fn, file=sys.stderr)
        sys.exit(0)


def main():
    parser = argparse.ArgumentParser(description="Generate variant candidates using alignments")
    parser.add_argument('--bam_fn', type=str, default="input.bam", help="Sorted bam file input (or .sam / .sam.gz text), default: %(default)s")
    parser.add_argument('--ref_fn', type=str, default="ref.fa", help="Reference fasta file input, default: %(default)s")
    parser.add_argument('--bed_fn', type=str, default=None,
                        help="Call variant only in these regions, works in intersection with ctgName, ctgStart and ctgEnd, optional")
    parser.add_argument('--can_fn', type=str, default="PIPE", help="Pile-up count output, use PIPE for standard output, default: %(default)s")
    parser.add_argument('--threshold', type=float, default=0.125,
                        help="Minimum allele frequence of the 1st non-reference allele for a site to be considered as a condidate "
                             "site, default: %(default)f")
    parser.add_argument('--minCoverage', type=float, default=4, help="Minimum coverage required to call a variant, default: %(default)f")
    parser.add_argument('--minMQ', type=int, default=0, help="Minimum Mapping Quality, default: %(default)d")
    parser.add_argument('--gen4Training', type=param.str2bool, nargs='?', const=True, default=False,
                        help="Output all genome positions as candidate for model training, default: %(default)s")
    parser.add_argument('--candidates', type=int, default=7000000, help="Use with gen4Training, default: %(default)s")
    parser.add_argument('--genomeSize', type=int, default=3000000000, help="Use with gen4Training, default: %(default)s")
    parser.add_argument('--ctgName', type=str, default="chr17", help="The name of sequence to be processed, default: %(default)s")
    parser.add_argument('--ctgStart', type=int, default=None, help="The 1-bsae starting position of the sequence to be processed")
    parser.add_argument('--ctgEnd', type=int, default=None, help="The inclusive ending position of the sequence to be processed")
    parser.add_argument('--samtools', type=str, default="samtools", help="Path to the 'samtools', default: %(default)s")
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    MakeCandidates(args)


if __name__ == "__main__":
    main()
