"""callVar -- call variants from candidate tensors with a trained model; Python-3 counterpart of reference
clairvoyante/callVar.py with the same command line (callVar.py:223-254) and the same VCF text.

  Run / Test    callVar.py:21-47, 180-216   three-way overlap: parse batch k+2 | predict batch k+1 | format batch k
  Output        callVar.py:50-153           per-site decisions; the argmax / QUAL / DP / AF arithmetic is done on
                                            whole batches with NumPy, strings only for the emitted records
  PrintVCFHeader callVar.py:156-178
"""
import argparse
import logging
import os
import sys
import time
from math import log
from threading import Thread

import numpy as np

from . import param

logging.basicConfig(format='%(message)s', level=logging.INFO)
num2base = "ACGT"
base2num = dict(zip("ACGT", (0, 1, 2, 3)))
maxVarLength = 5
inferIndelLengthMinimumAF = 0.125


def Run(args):
    logging.info("Loading model ...")
    if args.v2:
        sys.exit("clairvoyante_b200 implements the v3 / v3_slim networks only (--v2 is out of scope)")
    from . import utils_v2 as utils
    utils.SetupEnv()
    if args.slim:
        from . import clairvoyante_v3_slim as cv
    else:
        from . import clairvoyante_v3 as cv
    if args.threads is None:
        if args.tensor_fn == "PIPE":
            param.NUM_THREADS = 4
    else:
        param.NUM_THREADS = args.threads
    m = cv.Clairvoyante()
    m.init()
    m.restoreParameters(os.path.abspath(args.chkpnt_fn))
    Test(args, m, utils)


def _top2(a):
    s = np.sort(a, axis=1)
    return s[:, -1], s[:, -2]


def Output(args, call_fh, num, XBatch, posBatch, base, z, t, l):
    if num != len(base):
        sys.exit("Inconsistent shape between input tensor and output predictions %d/%d" % (num, len(base)))
    if num == 0:
        return
    F = param.flankingBaseNum
    varTypes = np.argmax(t, axis=1)
    emit = np.flatnonzero((varTypes != 0) | bool(args.showRef))
    if emit.size == 0:
        return
    zyg = np.argmax(z, axis=1)
    vlen = np.argmax(l, axis=1)
    t1, t2 = _top2(t); z1, z2 = _top2(z); l1, l2 = _top2(l)
    # float32 products, float64 ratio and log, truncation toward zero -- as in callVar.py:72
    ratio = ((t2 * z2 * l2).astype(np.float64) + 1e-300) / ((t1 * z1 * l1).astype(np.float64) + 1e-300)
    order = np.argsort(base, axis=1, kind="stable")[:, ::-1]          # ties: higher index first, like argsort()[::-1]
    # python sum() over float32 rows accumulates left to right in float32 (callVar.py:88-89)
    dp = np.zeros(num, np.float32)
    for src in (XBatch[:, F, :, 0], XBatch[:, F + 1, :, 1], XBatch[:, F + 1, :, 2], XBatch[:, F, :, 3]):
        s = np.zeros(num, np.float32)
        for k in range(4):
            s = s + src[:, k]
        dp = dp + s
    ins_cov = XBatch[:, F + 1, :, 1].sum(axis=1, dtype=np.float32)
    del_cov = XBatch[:, F + 1, :, 2].sum(axis=1, dtype=np.float32)
    out = []
    for j in emit:
        if dp[j] == 0:
            continue
        varType, varLength = int(varTypes[j]), int(vlen[j])
        chromosome, coordination, refSeq = posBatch[j].split(":")
        coordination = int(coordination)
        qual = int(-4.343 * log(ratio[j]))
        filt = "."
        if args.qual is not None:
            filt = "PASS" if qual >= args.qual else "LowQual"
        refBase = refSeq[F]; altBase = ""; inferred = 0; info = []; af = 0.0
        if varType <= 1:                                   # REF or SNP
            if varType == 1:
                b1, b2 = num2base[order[j, 0]], num2base[order[j, 1]]
                altBase = b1 if b1 != refBase else b2
            else:
                altBase = refBase
            af = XBatch[j, F, base2num[altBase], 3] / dp[j]
        elif varType == 2:                                 # INS
            if varLength == 0:
                varLength = 1
            af = ins_cov[j] / dp[j]
            if varLength != maxVarLength:
                for k in range(F + 1, F + varLength + 1):
                    altBase += num2base[int(np.argmax(XBatch[j, k, :, 1]))]
            else:
                for k in range(F + 1, 2 * F + 1):
                    ref_t, ins_t = XBatch[j, k, :, 0], XBatch[j, k, :, 1]
                    if k < (F + maxVarLength) or sum(ins_t) >= (inferIndelLengthMinimumAF * sum(ref_t)):
                        inferred += 1
                        altBase += num2base[int(np.argmax(ins_t))]
                    else:
                        break
            if inferred >= F:
                altBase = "<INS>"
                info.append("SVTYPE=INS")
            else:
                altBase = refBase + altBase
        else:                                              # DEL
            if varLength == 0:
                varLength = 1
            af = del_cov[j] / dp[j]
            if varLength == maxVarLength:
                for k in range(F + 1, 2 * F + 1):
                    if k < (F + maxVarLength) or sum(XBatch[j, k, :, 2]) >= (inferIndelLengthMinimumAF * sum(XBatch[j, k, :, 0])):
                        inferred += 1
                    else:
                        break
            if inferred >= F:
                altBase = "<DEL>"
                info.append("SVTYPE=DEL")
            elif varLength != maxVarLength:
                refBase = refSeq[F:F + varLength + 1]
                altBase = refSeq[F]
            else:
                refBase = refSeq[F:F + inferred + 1]
                altBase = refSeq[F]
        if 0 < inferred < F:
            info.append("LENGUESS=%d" % inferred)
        gt = "0/0" if varType == 0 else ("0/1" if zyg[j] == 0 else "1/1")
        out.append("%s\t%d\t.\t%s\t%s\t%d\t%s\t%s\tGT:GQ:DP:AF\t%s:%d:%d:%.4f" %
                   (chromosome, coordination, refBase, altBase, qual, filt, ";".join(info) if info else ".", gt, qual, dp[j], af))
    if out:
        call_fh.write("\n".join(out) + "\n")


def PrintVCFHeader(args, call_fh):
    lines = ['##fileformat=VCFv4.1',
             '##FILTER=<ID=PASS,Description="All filters passed">',
             '##FILTER=<ID=LowQual,Description="Confidence in this variant being real is below calling threshold.">',
             '##ALT=<ID=DEL,Description="Deletion">',
             '##ALT=<ID=INS,Description="Insertion of novel sequence">',
             '##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">',
             '##INFO=<ID=LENGUESS,Number=.,Type=Integer,Description="Best guess of the indel length">',
             '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
             '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">',
             '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read Depth">',
             '##FORMAT=<ID=AF,Number=1,Type=Float,Description="Estimated allele frequency in the range (0,1)">']
    if args.ref_fn is not None:
        with open(args.ref_fn + ".fai") as fai:
            for line in fai:
                f = line.strip().split("\t")
                lines.append("##contig=<ID=%s,length=%d>" % (f[0], int(f[1])))
    lines.append('#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s' % args.sampleName)
    call_fh.write("\n".join(lines) + "\n")


def Test(args, m, utils):
    """Pipeline of callVar.py:180-216: while batch k is written, batch k+1 is on the GPU and batch k+2 is parsed."""
    call_fh = open(args.call_fn, "w")
    PrintVCFHeader(args, call_fh)
    gen = utils.GetTensor(args.tensor_fn, param.predictBatchSize)
    logging.info("Calling variants ...")
    start = time.time()
    cur = next(gen)                                   # (endFlag, num, X, pos)
    m.predictNoRT(cur[2])
    preds = (m.predictBaseRTVal, m.predictZygosityRTVal, m.predictVarTypeRTVal, m.predictIndelLengthRTVal)
    nxt = next(gen) if cur[0] == 0 else None
    while True:
        workers = [Thread(target=Output, args=(args, call_fh, cur[1], cur[2], cur[3]) + preds)]
        if nxt is not None:
            workers.append(Thread(target=m.predictNoRT, args=(nxt[2],)))
        for w in workers:
            w.start()
        after = next(gen) if (nxt is not None and nxt[0] == 0) else None    # main thread parses meanwhile
        for w in workers:
            w.join()
        if nxt is None:
            break
        cur, nxt = nxt, after
        preds = (m.predictBaseRTVal, m.predictZygosityRTVal, m.predictVarTypeRTVal, m.predictIndelLengthRTVal)
    call_fh.close()
    logging.info("Total time elapsed: %.2f s" % (time.time() - start))


def main():
    parser = argparse.ArgumentParser(description="Call variants using a trained Clairvoyante model and tensors of candididate variants")
    parser.add_argument('--tensor_fn', type=str, default="PIPE", help="Tensor input, use PIPE for standard input")
    parser.add_argument('--chkpnt_fn', type=str, default=None, help="Input a checkpoint for testing or continue training")
    parser.add_argument('--call_fn', type=str, default=None, help="Output variant predictions")
    parser.add_argument('--qual', type=int, default=None,
                        help="If set, variant with equal or higher quality will be marked PASS, or LowQual otherwise, optional")
    parser.add_argument('--sampleName', type=str, default="SAMPLE", help="Define the sample name to be shown in the VCF file")
    parser.add_argument('--showRef', type=param.str2bool, nargs='?', const=True, default=False, help="Show reference calls, optional")
    parser.add_argument('--ref_fn', type=str, default=None,
                        help="Reference fasta file input, optional, print contig tags in the VCF header if set")
    parser.add_argument('--threads', type=int, default=None, help="Number of threads, optional")
    parser.add_argument('--v3', type=param.str2bool, nargs='?', const=True, default=True, help="Use Clairvoyante version 3")
    parser.add_argument('--v2', type=param.str2bool, nargs='?', const=True, default=False, help="Use Clairvoyante version 2")
    parser.add_argument('--slim', type=param.str2bool, nargs='?', const=True, default=False,
                        help="Train using the slim version of Clairvoyante, optional")
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    Run(args)


if __name__ == "__main__":
    main()
