"""callVarBam -- call variants straight from alignments; counterpart of reference clairvoyante/callVarBam.py with the same
command line (:146-213).

The reference starts three processes joined by pipes (:113-131): `ExtractVariantCandidates.py | CreateTensor.py | callVar.py`
(or `GetTruth.py` first when `--vcf_fn` gives the candidate sites), each a Python interpreter, the first two running
per-base loops.  Here the three stages run in ONE process on the native stages:
    candidates  csrc/candidates.cpp  (ExtractVariantCandidates.py)        pass 1 over the alignments
    tensors     csrc/pileup.cpp      (CreateTensor.py)                    pass 2, batches handed to the model in memory
    calls       callVar.Test         (callVar.py:180-216)                 predict on the GPU | format VCF, overlapped
so no tensor text is printed or parsed.  Alignments and reference come from `samtools` when it is installed (the same
`view -F 2308` / `faidx` commands, issued once per pass like the reference's two stages) or from `.sam` / `.sam.gz` + FASTA.
`--pypy`, `--threads` and `--delay` are accepted for command-line compatibility; there is nothing left for them to tune.
"""
import argparse
import os
import sys
import types

import numpy as np

from . import CreateTensor as CT, ExtractVariantCandidates as EVC, callVar, param


def _vcf_positions(vcf_fn, ctgName, ctgStart, ctgEnd):
    """candidate sites from a VCF (the reference runs dataPrepScripts/GetTruth.py:47-58 for this): rows of this contig
    inside [ctgStart, ctgEnd] (ctgStart already the 1-based value)"""
    from .utils_v2 import open_maybe_gzip
    out = []
    with open_maybe_gzip(vcf_fn) as f:
        for row in f:
            row = row.strip().split()
            if not row or row[0][0] == "#" or row[0] != ctgName:
                continue
            p = int(row[1])
            if ctgStart is not None and ctgEnd is not None and (p < ctgStart or p > ctgEnd):
                continue
            out.append(p)
    return out


class _ChunkCache(object):
    """keeps the alignment text of pass 1 for pass 2 while it fits (CVB_SAM_CACHE_MB, default 2048): the reference decodes the
    BAM twice (one `samtools view` per stage, callVarBam.py:113-126); here the second decode is skipped when memory allows"""

    def __init__(self):
        self.limit = int(os.environ.get("CVB_SAM_CACHE_MB", 2048)) << 20
        self.chunks, self.bytes, self.complete = [], 0, False

    def tee(self, chunks):
        keep = self.limit > 0
        for b in chunks:
            if keep:
                self.bytes += len(b)
                if self.bytes > self.limit:
                    keep, self.chunks = False, []
                else:
                    self.chunks.append(b)
            yield b
        self.complete = keep


class _AlignmentFeed(object):
    """stands where callVar.Test expects `utils`: GetTensor(tensor_fn, num) yields (endFlag, n, X, pos)"""

    def __init__(self, args, ref_seq, ref_start, positions, cache=None):
        self.args, self.ref_seq, self.ref_start, self.positions, self.cache = args, ref_seq, ref_start, positions, cache

    def GetTensor(self, tensor_fn, num):
        a = self.args
        if self.cache is not None and self.cache.complete:
            proc, sam, source = None, None, iter(self.cache.chunks)
        else:
            proc, sam = CT._open_alignments(a)
            source = sam
        try:
            for item in CT.GetTensorFromAlignments(source, self.ref_seq, self.positions, a.ctgName, num, self.ref_start,
                                                   dcov=a.dcov, considerleftedge=a.considerleftedge):
                yield item
        finally:
            if sam is not None:
                sam.close()
            if proc is not None:
                proc.wait()


def Run(args, model=None):
    if args.ctgName is None:
        sys.exit("--ctgName must be specified. You can call variants on multiple chromosomes simultaneously.")
    if args.v2:
        sys.exit("clairvoyante_b200 implements the v3 / v3_slim networks only (--v2 is out of scope)")
    if not (args.ctgStart is not None and args.ctgEnd is not None and int(args.ctgStart) <= int(args.ctgEnd)):
        args.ctgStart = args.ctgEnd = None                       # callVarBam.py:88-91
    ref_seq, ref_start = CT._load_reference(args)                # ctgStart becomes the 1-based value both stages compare with
    # ---- stage 1: candidate positions
    if args.vcf_fn is None:
        bed = EVC._load_bed(args.bed_fn, args.ctgName) if args.bed_fn is not None else None
        proc, sam = CT._open_alignments(args)
        cache = _ChunkCache()
        pos = [p for _, p in EVC.extract_candidates(cache.tee(CT._chunks(sam)), args.ctgName, ref_seq, ref_start, ctgStart=args.ctgStart, ctgEnd=args.ctgEnd,
                                                    bed=bed, minCoverage=args.minCoverage, threshold=args.threshold)]
        sam.close()
        if proc is not None:
            proc.wait()
        positions = np.concatenate(pos) if pos else np.empty((0,), np.int64)
    else:
        cache = None
        positions = np.asarray(_vcf_positions(args.vcf_fn, args.ctgName, args.ctgStart, args.ctgEnd), np.int64)
    if args.ctgStart is not None:                                # CreateTensor's own candidate filter (CreateTensor.py:70-71)
        positions = positions[(positions >= args.ctgStart) & (positions <= args.ctgEnd)]
    # ---- stages 2 + 3
    if model is None:
        if args.slim:
            from . import clairvoyante_v3_slim as cv
        else:
            from . import clairvoyante_v3 as cv
        model = cv.Clairvoyante()
        model.init()
        model.restoreParameters(os.path.abspath(args.chkpnt_fn))
    cv_args = types.SimpleNamespace(tensor_fn="(alignments)", call_fn=args.call_fn, qual=args.qual or None, sampleName=args.sampleName,
                                    showRef=False, ref_fn=args.ref_fn if os.path.isfile(args.ref_fn + ".fai") else None)
    callVar.Test(cv_args, model, _AlignmentFeed(args, ref_seq, ref_start, positions, cache))
    return len(positions)


def main():
    parser = argparse.ArgumentParser(description="Call variants using a trained Clairvoyante model and a BAM file")
    parser.add_argument('--chkpnt_fn', type=str, default=None, help="Input a Clairvoyante model")
    parser.add_argument('--ref_fn', type=str, default="ref.fa", help="Reference fasta file input, default: %(default)s")
    parser.add_argument('--bed_fn', type=str, default=None,
                        help="Call variant only in these regions, works in intersection with ctgName, ctgStart and ctgEnd, optional")
    parser.add_argument('--bam_fn', type=str, default="bam.bam", help="BAM file input (or .sam / .sam.gz text), default: %(default)s")
    parser.add_argument('--call_fn', type=str, default=None, help="Output variant predictions")
    parser.add_argument('--vcf_fn', type=str, default=None,
                        help="Candidate sites VCF file input, if provided, variants will only be called at the sites in the VCF file")
    parser.add_argument('--threshold', type=float, default=0.125,
                        help="Minimum allele frequence of the 1st non-reference allele for a site to be considered as a condidate "
                             "site, default: %(default)f")
    parser.add_argument('--minCoverage', type=float, default=4, help="Minimum coverage required to call a variant, default: %(default)d")
    parser.add_argument('--qual', type=int, default=None,
                        help="If set, variant with equal or higher quality will be marked PASS, or LowQual otherwise, optional")
    parser.add_argument('--sampleName', type=str, default="SAMPLE", help="Define the sample name to be shown in the VCF file")
    parser.add_argument('--ctgName', type=str, default=None, help="The name of sequence to be processed")
    parser.add_argument('--ctgStart', type=int, default=None, help="The 1-bsae starting position of the sequence to be processed")
    parser.add_argument('--ctgEnd', type=int, default=None, help="The inclusive ending position of the sequence to be processed")
    parser.add_argument('--considerleftedge', type=param.str2bool, nargs='?', const=True, default=True,
                        help="Count the left-most base-pairs of a read for coverage even if the starting position of a read is "
                             "after the starting position of a tensor, default: %(default)s")
    parser.add_argument('--dcov', type=int, default=250, help="Cap depth per position at %(default)s")
    parser.add_argument('--samtools', type=str, default="samtools", help="Path to the 'samtools', default: %(default)s")
    parser.add_argument('--pypy', type=str, default="pypy", help="(ignored: no interpreter is started)")
    parser.add_argument('--v3', type=param.str2bool, nargs='?', const=True, default=True, help="Use Clairvoyante version 3")
    parser.add_argument('--v2', type=param.str2bool, nargs='?', const=True, default=False, help="Use Clairvoyante version 2")
    parser.add_argument('--slim', type=param.str2bool, nargs='?', const=True, default=False, help="Use the slim version of Clairvoyante")
    parser.add_argument('--threads', type=int, default=None, help="(ignored)")
    parser.add_argument('--delay', type=int, default=10, help="(ignored: one process, no thread-pool start-up to stagger)")
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    Run(args)


if __name__ == "__main__":
    main()
