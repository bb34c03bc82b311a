"""callVar -- call variants from candidate tensors with a trained model; Python-3 counterpart of reference
clairvoyante/callVar.py with the same command line (callVar.py:223-254) and the same VCF text.

  Run / Test    callVar.py:21-47, 180-216   three-way overlap: parse batch k+2 | predict batch k+1 | format batch k
  Output        callVar.py:50-153           per-site decisions and the record text: one native pass over the batch
                                            (cvb_vcf_records, csrc/vcf_text.cpp)
  PrintVCFHeader callVar.py:156-178
"""
import argparse
import logging
import os
import sys
import time
from threading import Thread

import numpy as np

from . import _lib, param

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


def Output(args, call_fh, num, XBatch, posBatch, base, z, t, l):
    """callVar.py:50-153.  The per-site record logic runs in the C library (cvb_vcf_records, csrc/vcf_text.cpp: one pass over the
    batch instead of a Python loop per site); oracle/callvar_output.py is its scalar restatement."""
    if num != len(base):
        sys.exit("Inconsistent shape between input tensor and output predictions %d/%d" % (num, len(base)))
    if num == 0:
        return
    lib = _lib.load()
    x = np.ascontiguousarray(XBatch, dtype=np.float32)
    heads = [np.ascontiguousarray(h, dtype=np.float32) for h in (base, z, t, l)]
    if x.shape[1:] != (2 * param.flankingBaseNum + 1, 4, param.matrixNum) or [h.shape[1] for h in heads] != [4, 2, 4, 6]:
        sys.exit("Output: unexpected tensor or prediction shape")
    pos = "\n".join(posBatch).encode("ascii", "replace")
    out = np.empty(len(pos) + 256 * num + 64, np.uint8)
    n = lib.cvb_vcf_records(x.ctypes.data, pos, len(pos), heads[0].ctypes.data, heads[1].ctypes.data, heads[2].ctypes.data,
                            heads[3].ctypes.data, num, 1 if args.showRef else 0, -1 if args.qual is None else int(args.qual),
                            out.ctypes.data, len(out))
    if n < 0:
        _lib.check(1)
    if n:
        call_fh.write(out[:n].tobytes().decode("ascii", "replace"))


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
    failed = []

    def guarded(fn, *a):
        # an exception (or sys.exit) on a worker thread would otherwise end only that thread: the VCF would silently miss a
        # batch and the process would still exit 0
        try:
            fn(*a)
        except BaseException as e:          # noqa: B902 -- SystemExit included
            failed.append(e)

    while True:
        workers = [Thread(target=guarded, args=(Output, args, call_fh, cur[1], cur[2], cur[3]) + preds)]
        if nxt is not None:
            workers.append(Thread(target=guarded, args=(m.predictNoRT, nxt[2])))
        for w in workers:
            w.start()
        after = next(gen) if (nxt is not None and nxt[0] == 0) else None    # main thread parses meanwhile
        for w in workers:
            w.join()
        if failed:
            call_fh.close()
            raise failed[0]
        if nxt is None:
            break
        cur, nxt = nxt, after
        preds = (m.predictBaseRTVal, m.predictZygosityRTVal, m.predictVarTypeRTVal, m.predictIndelLengthRTVal)
    call_fh.close()
    logging.info("Total time elapsed: %.2f s" % (time.time() - start))


def main():
    from . import _driver as D
    parser = argparse.ArgumentParser(description="Call variants using a trained Clairvoyante model and tensors of candididate variants")
    add = parser.add_argument
    add('--tensor_fn', type=str, default="PIPE", help="Tensor text (PIPE = standard input)")
    add('--chkpnt_fn', type=str, default=None, help="Model checkpoint")
    add('--call_fn', type=str, default=None, help="VCF to write")
    add('--qual', type=int, default=None, help="Mark records PASS at or above this quality, LowQual below; optional")
    add('--sampleName', type=str, default="SAMPLE", help="Sample column of the VCF")
    add('--showRef', type=param.str2bool, nargs='?', const=True, default=False, help="Also write reference calls")
    add('--ref_fn', type=str, default=None, help="Reference FASTA: its .fai supplies the ##contig header lines, optional")
    add('--threads', type=int, default=None, help="Accepted for compatibility (TensorFlow's thread count in the reference)")
    D.variant_options(parser)
    Run(D.parse(parser))


if __name__ == "__main__":
    main()
