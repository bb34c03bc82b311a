"""evaluate -- score a trained model on a labelled tensor set; Python-3 counterpart of reference clairvoyante/evaluate.py
with the same command line (evaluate.py:113-140) and the same report: top-1 / top-2 accuracy of the base-change head
(:77-89) and the confusion matrices of zygosity, variant type and indel length (:90-110), after predicting the whole set
in batches of predictBatchSize (:55-74)."""
import argparse
import logging
import time

import numpy as np

from . import _driver as D, param

logging.basicConfig(format='%(message)s', level=logging.INFO)


def Run(args):
    logging.info("Loading model ...")
    m, utils = D.new_model(args)
    m.restoreParameters(D.absolute(args.chkpnt_fn))
    Test(args, m, utils)


def predict_all(m, utils, XBlocks, total):
    """evaluate.py:55-74"""
    outs = [[], [], [], []]
    ptr = 0
    while ptr < total:
        XBatch, _, endFlag = utils.DecompressArray(XBlocks, ptr, param.predictBatchSize, total)
        for acc, o in zip(outs, m.predict(XBatch)):
            acc.append(o)
        ptr += param.predictBatchSize
        if endFlag != 0:
            break
    return [np.concatenate(o) for o in outs]


def report(bases, zs, ts, ls, YArray):
    """evaluate.py:77-110 -> dict(all, top1, top2, zygosity, varType, indelLength); also logged in the reference's wording"""
    logging.info("Version 2 model, evaluation on base change:")
    order = np.argsort(bases, axis=1, kind="stable")[:, ::-1]
    truth = np.argmax(YArray[:, 0:4], axis=1)
    top1 = int((order[:, 0] == truth).sum())
    top2 = top1 + int((order[:, 1] == truth).sum())
    n = len(bases)
    logging.info("all/top1/top2/top1p/top2p: %d/%d/%d/%.2f/%.2f" % (n, top1, top2, float(top1) / max(n, 1) * 100,
                                                                     float(top2) / max(n, 1) * 100))
    res = dict(all=n, top1=top1, top2=top2)
    for key, title, pred, lo, hi in (("zygosity", "Zygosity", zs, 4, 6), ("varType", "variant type", ts, 6, 10),
                                     ("indelLength", "indel length", ls, 10, 16)):
        logging.info("Version 2 model, evaluation on %s:" % title)
        ed = np.zeros((hi - lo, hi - lo), dtype=np.int64)
        np.add.at(ed, (np.argmax(YArray[:, lo:hi], axis=1), np.argmax(pred, axis=1)), 1)
        for row in ed:
            logging.info("\t".join(str(v) for v in row))
        res[key] = ed
    return res


def Test(args, m, utils):
    logging.info("Loading the dataset ...")
    if args.bin_fn is not None:
        total, XBlocks, YBlocks, _ = utils.load_bin(args.bin_fn)
    else:
        total, XBlocks, YBlocks, _ = utils.GetTrainingArray(args.tensor_fn, args.var_fn, args.bed_fn)
    logging.info("Dataset size: %d" % total)
    logging.info("Testing on the dataset ...")
    predictStart = time.time()
    bases, zs, ts, ls = predict_all(m, utils, XBlocks, total)
    logging.info("Prediciton time elapsed: %.2f s" % (time.time() - predictStart))
    YArray, _, _ = utils.DecompressArray(YBlocks, 0, total, total)
    return report(bases, zs, ts, ls, YArray)


def main():
    parser = argparse.ArgumentParser(description="Evaluate trained Clairvoyante model")
    D.dataset_options(parser)
    parser.add_argument('--chkpnt_fn', type=str, default=None, help="Checkpoint to evaluate")
    D.variant_options(parser)
    Run(D.parse(parser))


if __name__ == "__main__":
    main()
