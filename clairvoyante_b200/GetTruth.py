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


def truth_rows(lines, ctgName, ctgStart=None, ctgEnd=None):
    """ctgStart is the value the reference compares with (its --ctgStart + 1, :35-36)"""
    for row in lines:
        row = row.strip().split()
        if not row or row[0][0] == "#":
            continue
        if row[0] != ctgName:
            continue
        if ctgStart is not None and ctgEnd is not None:
            if int(row[1]) < ctgStart or int(row[1]) > ctgEnd:
                continue
        last = row[-1]
        varType = last.split(":")[0].replace("/", "|").replace(".", "0").split("|")
        p1, p2 = varType
        p1, p2 = int(p1), int(p2)
        p1, p2 = (p1, p2) if p1 < p2 else (p2, p1)
        if p1 == 1 and p2 == 2 and row[4].find(",") != -1:
            p1, p2 = 0, 1
            shortestLen, shortestGT = 99, ""
            for i in row[4].split(","):
                if len(i) < shortestLen:
                    shortestLen, shortestGT = len(i), i
            row[4] = shortestGT
        yield " ".join([row[0], row[1], row[3], row[4], str(p1), str(p2)])


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
    parser.add_argument('--vcf_fn', type=str, default="input.vcf", help="Truth vcf file input, default: %(default)s")
    parser.add_argument('--var_fn', type=str, default="PIPE", help="Truth variants output, use PIPE for standard output, default: %(default)s")
    parser.add_argument('--ctgName', type=str, default="chr17", help="The name of sequence to be processed, default: %(default)s")
    parser.add_argument('--ctgStart', type=int, default=None, help="The 1-bsae starting position of the sequence to be processed")
    parser.add_argument('--ctgEnd', type=int, default=None, help="The inclusive ending position of the sequence to be processed")
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    OutputVariant(args)


if __name__ == "__main__":
    main()
