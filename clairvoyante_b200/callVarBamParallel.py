"""callVarBamParallel -- print one callVarBam command per genome chunk; counterpart of reference
clairvoyante/callVarBamParallel.py with the same command line (:92-145) and the same chunking (:66-89): contigs of the
`.fai` index (major contigs unless --includingAllContigs) cut into --refChunkSize pieces, chunks without BED coverage
skipped.  The printed commands run `python -m clairvoyante_b200.callVarBam`, one process per chunk on one GPU each
(`CVB_DEVICE` / `CUDA_VISIBLE_DEVICES` select it): sites are independent, no collective."""
import argparse
import gzip
import os
import sys

from . import param

majorContigs = {"chr" + str(a) for a in list(range(0, 23)) + ["X", "Y"]}.union({str(a) for a in list(range(0, 23)) + ["X", "Y"]})


def _check(fn, sfx=""):
    if not os.path.isfile(fn + sfx):
        sys.exit("Error: %s not found" % (fn + sfx))
    return os.path.abspath(fn)


def commands(args):
    chkpnt_fn = os.path.abspath(args.chkpnt_fn)
    bam_fn, ref_fn = _check(args.bam_fn), _check(args.ref_fn)
    fai_fn = _check(args.ref_fn, ".fai") + ".fai"
    bed_fn = _check(args.bed_fn) if args.bed_fn is not None else None
    tree = {}
    if bed_fn is not None:                                              # :51-64
        opener = gzip.open if bed_fn.endswith(".gz") else open
        with opener(bed_fn, "rt") as f:
            for row in f:
                row = row.strip().split()
                if len(row) < 3:
                    continue
                begin, end = int(row[1]), int(row[2]) - 1
                if end == begin:
                    end += 1
                tree.setdefault(row[0], []).append((begin, end))
    tail = " ".join(x for x in ("--vcf_fn %s" % _check(args.vcf_fn) if args.vcf_fn is not None else "",
                                "--considerleftedge" if args.considerleftedge else "",
                                "--qual %d" % args.qual if args.qual else "", "--slim" if args.slim else "") if x)
    out = []
    for line in open(fai_fn):                                            # :66-89
        fields = line.strip().split("\t")
        chromName = fields[0]
        if not args.includingAllContigs and str(chromName) not in majorContigs:
            continue
        regionStart, chromLength = 0, int(fields[1])
        while regionStart < chromLength:
            start, end = regionStart, min(regionStart + args.refChunkSize, chromLength)
            output_fn = "%s.%s_%d_%d.vcf" % (args.output_prefix, chromName, regionStart, end)
            bed_part = ""
            if bed_fn is not None:
                if not any(b < end and e > start for b, e in tree.get(chromName, [])):   # len(tree.search(start, end)) == 0
                    regionStart = end
                    continue
                bed_part = "--bed_fn %s " % bed_fn
            out.append("python -m clairvoyante_b200.callVarBam --chkpnt_fn %s --ref_fn %s --bam_fn %s %s--ctgName %s --ctgStart %d "
                       "--ctgEnd %d --call_fn %s --threshold %f --minCoverage %f --samtools %s --sampleName %s %s"
                       % (chkpnt_fn, ref_fn, bam_fn, bed_part, chromName, regionStart, end, output_fn, args.threshold,
                          args.minCoverage, args.samtools, args.sampleName, tail))
            regionStart = end
    return [c.rstrip() for c in out]


def main():
    parser = argparse.ArgumentParser(
        description="Create commands for calling variants in parallel using a trained Clairvoyante model and a BAM file")
    parser.add_argument('--chkpnt_fn', type=str, default=None, help="Input a Clairvoyante model")
    parser.add_argument('--ref_fn', type=str, default="ref.fa", help="Reference fasta file input, default: %(default)s")
    parser.add_argument('--bed_fn', type=str, default=None, help="Call variant only in these regions, optional, default: whole genome")
    parser.add_argument('--refChunkSize', type=int, default=10000000,
                        help="Divide job with smaller genome chunk size for parallelism, default: %(default)s")
    parser.add_argument('--bam_fn', type=str, default="bam.bam", help="BAM file input, default: %(default)s")
    parser.add_argument('--vcf_fn', type=str, default=None, help="Candidate sites VCF file input, optional")
    parser.add_argument('--output_prefix', type=str, default=None, help="Output prefix")
    parser.add_argument('--includingAllContigs', type=param.str2bool, nargs='?', const=True, default=False,
                        help="Call variants on all contigs, default: chr{1..22,X,Y} and {1..22,X,Y}")
    parser.add_argument('--tensorflowThreads', type=int, default=4, help="(ignored: no TensorFlow)")
    parser.add_argument('--threshold', type=float, default=0.2,
                        help="Minimum allele frequence of the 1st non-reference allele for a site to be considered as a condidate "
                             "site, default: %(default)f")
    parser.add_argument('--minCoverage', type=float, default=4, help="Minimum coverage required to call a variant, default: %(default)d")
    parser.add_argument('--qual', type=int, default=None,
                        help="If set, variant with equal or higher quality will be marked PASS, or LowQual otherwise, optional")
    parser.add_argument('--sampleName', type=str, default="SAMPLE", help="Define the sample name to be shown in the VCF file")
    parser.add_argument('--considerleftedge', type=param.str2bool, nargs='?', const=True, default=True,
                        help="Count the left-most base-pairs of a read for coverage, default: %(default)s")
    parser.add_argument('--samtools', type=str, default="samtools", help="Path to the 'samtools', default: %(default)s")
    parser.add_argument('--pypy', type=str, default="pypy", help="(ignored)")
    parser.add_argument('--delay', type=int, default=10, help="(ignored)")
    parser.add_argument('--slim', type=param.str2bool, nargs='?', const=True, default=False, help="Use the slim model")
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    print("\n".join(commands(args)))


if __name__ == "__main__":
    main()
