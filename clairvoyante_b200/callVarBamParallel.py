"""callVarBamParallel -- print one callVarBam command per genome chunk; counterpart of reference
clairvoyante/callVarBamParallel.py (same options, :92-145; same chunking, :66-89): the contigs of the `.fai` index (the
major ones unless --includingAllContigs) are cut into --refChunkSize pieces and pieces without BED coverage are left out.
The commands run `python -m clairvoyante_b200.callVarBam`, one process per chunk on one GPU each (`CVB_DEVICE` /
`CUDA_VISIBLE_DEVICES` select it): sites are independent, there is no collective."""
import argparse
import os
import sys

from . import _driver as D, param

_NUMBERED = [str(i) for i in range(23)] + ["X", "Y"]
majorContigs = set(_NUMBERED) | {"chr" + c for c in _NUMBERED}


def _existing(fn, suffix=""):
    if not os.path.isfile(fn + suffix):
        sys.exit("Error: %s not found" % (fn + suffix))
    return os.path.abspath(fn)


def _bed_intervals(bed_fn):
    """{contig: [(begin, end)]} with the reference's end - 1 and never-empty rule (:56-60)"""
    covered = {}
    from .utils_v2 import open_maybe_gzip
    with open_maybe_gzip(bed_fn) as f:
        for fields in (line.split() for line in f):
            if len(fields) >= 3:
                lo, hi = int(fields[1]), int(fields[2]) - 1
                covered.setdefault(fields[0], []).append((lo, hi + 1 if hi == lo else hi))
    return covered


def _chunks(fai_fn, size, every_contig):
    """(contig, start, end) pieces in index order"""
    with open(fai_fn) as fai:
        for fields in (line.strip().split("\t") for line in fai):
            if not every_contig and fields[0] not in majorContigs:
                continue
            length = int(fields[1])
            for start in range(0, length, size):
                yield fields[0], start, min(start + size, length)


def commands(args):
    model = os.path.abspath(args.chkpnt_fn)
    bam, ref = _existing(args.bam_fn), _existing(args.ref_fn)
    _existing(args.ref_fn, ".fai")
    bed = _existing(args.bed_fn) if args.bed_fn is not None else None
    covered = _bed_intervals(bed) if bed else None
    extras = [("--vcf_fn %s" % _existing(args.vcf_fn)) if args.vcf_fn is not None else "", "--considerleftedge" if args.considerleftedge else "",
              ("--qual %d" % args.qual) if args.qual else "", "--slim" if args.slim else ""]
    lines = []
    for contig, start, end in _chunks(ref + ".fai", args.refChunkSize, args.includingAllContigs):
        if covered is not None and not any(lo < end and hi > start for lo, hi in covered.get(contig, ())):   # tree.search(start, end) empty
            continue
        opts = ["--chkpnt_fn", model, "--ref_fn", ref, "--bam_fn", bam] + (["--bed_fn", bed] if bed else []) + [
            "--ctgName", contig, "--ctgStart", str(start), "--ctgEnd", str(end),
            "--call_fn", "%s.%s_%d_%d.vcf" % (args.output_prefix, contig, start, end), "--threshold", "%f" % args.threshold,
            "--minCoverage", "%f" % args.minCoverage, "--samtools", args.samtools, "--sampleName", args.sampleName]
        lines.append(" ".join(["python -m clairvoyante_b200.callVarBam"] + opts + [x for x in extras if x]))
    return lines


def main():
    parser = argparse.ArgumentParser(description="Create commands for calling variants in parallel using a trained Clairvoyante model and a BAM file")
    add = parser.add_argument
    add('--chkpnt_fn', type=str, default=None, help="Model checkpoint")
    add('--ref_fn', type=str, default="ref.fa", help="Reference FASTA (with .fai), default: %(default)s")
    add('--bed_fn', type=str, default=None, help="Only chunks touching these regions, optional")
    add('--refChunkSize', type=int, default=10000000, help="Chunk length, default: %(default)s")
    add('--bam_fn', type=str, default="bam.bam", help="Alignments, default: %(default)s")
    add('--vcf_fn', type=str, default=None, help="Candidate sites VCF passed on to callVarBam, optional")
    add('--output_prefix', type=str, default=None, help="Prefix of the per-chunk VCF names")
    add('--includingAllContigs', type=param.str2bool, nargs='?', const=True, default=False, help="All contigs of the index, not only 1-22, X, Y")
    add('--tensorflowThreads', type=int, default=4, help="(ignored: no TensorFlow)")
    add('--threshold', type=float, default=0.2, help="Candidate allele-frequency threshold, default: %(default)f")
    add('--minCoverage', type=float, default=4, help="Candidate minimum coverage, default: %(default)d")
    add('--qual', type=int, default=None, help="PASS / LowQual cut, optional")
    add('--sampleName', type=str, default="SAMPLE", help="Sample column of the VCF")
    add('--considerleftedge', type=param.str2bool, nargs='?', const=True, default=True, help="Passed on to callVarBam, default: %(default)s")
    add('--samtools', type=str, default="samtools", help="samtools executable, default: %(default)s")
    add('--pypy', type=str, default="pypy", help="(ignored)")
    add('--delay', type=int, default=10, help="(ignored)")
    add('--slim', type=param.str2bool, nargs='?', const=True, default=False, help="slim network")
    print("\n".join(commands(D.parse(parser))))


if __name__ == "__main__":
    main()
