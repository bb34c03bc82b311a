"""GetTruth -- truth VCF -> `ctg pos ref alt gt1 gt2` rows (the --var_fn input of train.py / tensor2Bin.py and, through
their first two columns, the candidate list of CreateTensor.py); Python-3 counterpart of reference
dataPrepScripts/GetTruth.py (same command line, :89-108).  Per record (:52-79): genotype from the LAST column
(`/` and `|` alike, `.` = 0), alleles ordered, and a 1/2 site with two ALT alleles becomes 0/1 with the SHORTEST ALT.
`tabix` is used for region queries when both the index and the program exist (:41-47); otherwise the file is scanned."""
import argparse
import gzip
import os
import shlex
import shutil
import subprocess
import sys


def _genotype(sample_field):
    """allele indices of the GT sub-field, ascending: '1|0' -> (0, 1), './1' -> (0, 1)"""
    a, b = (int(x) for x in sample_field.split(":")[0].replace("/", "|").replace(".", "0").split("|"))
    return (a, b) if a < b else (b, a)


def truth_rows(lines, ctgName, ctgStart=None, ctgEnd=None):
    """ctgStart is the value the reference compares with (its --ctgStart + 1, :35-36)"""
    bounded = ctgStart is not None and ctgEnd is not None
    for line in lines:
        f = line.strip().split()
        if not f or f[0][0] == "#" or f[0] != ctgName:
            continue
        if bounded and not ctgStart <= int(f[1]) <= ctgEnd:
            continue
        g1, g2 = _genotype(f[-1])
        alt = f[4]
        if (g1, g2) == (1, 2) and "," in alt:       # two ALT alleles: reported as 0/1 with the shortest one (first of equals; :67-77)
            g1, g2 = 0, 1
            alt = min(alt.split(","), key=lambda s: 99 if len(s) >= 99 else len(s))
            if len(alt) >= 99:
                alt = ""
        yield " ".join((f[0], f[1], f[3], alt, str(g1), str(g2)))


def OutputVariant(args):
    ctgStart, ctgEnd = args.ctgStart, args.ctgEnd
    if ctgStart is not None and ctgEnd is not None:
        ctgStart += 1
    out = sys.stdout if args.var_fn == "PIPE" else gzip.open(args.var_fn, "wt")
    proc = None
    if ctgStart is not None and ctgEnd is not None and os.path.isfile(args.vcf_fn + ".tbi") and shutil.which("tabix"):
        proc = subprocess.Popen(shlex.split("tabix -f -p vcf %s %s:%s-%s" % (args.vcf_fn, args.ctgName, ctgStart, ctgEnd)),
                                stdout=subprocess.PIPE, bufsize=8388608, text=True)
        lines = proc.stdout
    else:
        with open(args.vcf_fn, "rb") as probe:
            gz = probe.read(2) == b"\x1f\x8b"                       # `gzip -fdc` passes plain files through
        lines = gzip.open(args.vcf_fn, "rt") if gz else open(args.vcf_fn, "rt")
    for r in truth_rows(lines, args.ctgName, ctgStart, ctgEnd):
        out.write(r + "\n")
    lines.close()
    if proc is not None:
        proc.wait()
    if out is not sys.stdout:
        out.close()


def main():
    parser = argparse.ArgumentParser(description="Extract variant type and allele from a Truth dataset")
    parser.add_argument('--vcf_fn', type=str, default="input.vcf", help="Truth VCF, default: %(default)s")
    parser.add_argument('--var_fn', type=str, default="PIPE", help="Output (PIPE = standard output), default: %(default)s")
    parser.add_argument('--ctgName', type=str, default="chr17", help="Contig to extract, default: %(default)s")
    parser.add_argument('--ctgStart', type=int, default=None, help="First position of the region (as the reference takes it)")
    parser.add_argument('--ctgEnd', type=int, default=None, help="Last position of the region, inclusive")
    from . import _driver as D
    OutputVariant(D.parse(parser))


if __name__ == "__main__":
    main()
