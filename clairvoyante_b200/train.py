"""train -- adaptive-learning-rate training driver; Python-3 counterpart of reference clairvoyante/train.py with
the same command line (train.py:225-259) and the same schedule:

  * first trainingDatasetPercentage (90 %) of the shuffled set trains in batches of trainBatchSize (10,000), the
    rest validates in batches of predictBatchSize (1,000)                          train.py:67-102
  * the model call of batch k runs on a thread while batch k+1 is decompressed     train.py:86-109
  * one checkpoint per epoch `<prefix>-%06d`, resume epoch = last 6 chars          train.py:82,127-129
  * learning rate and L2 lambda x0.1 when the last five validation-loss differences zig-zag (or stall),
    at most maxLearningRateSwitch times                                            train.py:131-154
  * afterwards: predict the whole set and log top-1/top-2 + confusion matrices     train.py:169-218
"""
import argparse
import logging
import time

import numpy as np

from . import _driver as D, param

logging.basicConfig(format='%(message)s', level=logging.INFO)


def Run(args):
    logging.info("Initializing model ...")
    m, utils = D.new_model(args)
    if args.chkpnt_fn is not None:
        m.restoreParameters(D.absolute(args.chkpnt_fn))
    TrainAll(args, m, utils)


def next_batch_size(ptr, validationStart):
    """rows to fetch at dataset position `ptr` (train.py:95-102)"""
    return D.rows_wanted(ptr, validationStart)


def switch_needed(losses):
    """True when the last five differences of the validation loss alternate in sign, or the oldest is exactly
    zero (train.py:133-148)"""
    if len(losses) < 6:
        return False
    d = [losses[i] - losses[i + 1] for i in range(-6, -1)]
    if d[0] > 0:
        return d[1] < 0 and d[2] > 0 and d[3] < 0 and d[4] > 0
    if d[0] < 0:
        return d[1] > 0 and d[2] < 0 and d[3] > 0 and d[4] < 0
    return True


def _confusion(pred, truth, k):
    ed = np.zeros((k, k), dtype=np.int64)
    np.add.at(ed, (np.argmax(truth, axis=1), np.argmax(pred, axis=1)), 1)
    return ed


def TrainAll(args, m, utils):
    logging.info("Loading the training dataset ...")
    data = D.TrainingSet(args, utils)
    total, XBlocks, YBlocks = data.total, data.X, data.Y
    summaryWriter = D.announce_training(args, m, data)

    validationLosses = []
    trainingStart = time.time()
    switchesLeft = param.maxLearningRateSwitch
    sinceSwitch = 0
    for epoch in range(D.first_epoch(args), param.maxEpoch):
        _, validationLossSum = D.train_validate_epoch(m, data, epoch, summaryWriter)
        validationLosses.append((validationLossSum, epoch))
        if args.ochk_prefix is not None:
            m.saveParameters(D.absolute(D.checkpoint_name(args.ochk_prefix, epoch)))
        sinceSwitch += 1
        if sinceSwitch >= 6 and switch_needed([v for v, _ in validationLosses]):
            switchesLeft -= 1
            if switchesLeft == 0:
                break
            logging.info("New learning rate: %.2e" % m.setLearningRate())
            logging.info("New L2 regularization lambda: %.2e" % m.setL2RegularizationLambda())
            sinceSwitch = 0

    logging.info("Training time elapsed: %.2f s" % (time.time() - trainingStart))
    best = sorted(validationLosses)[0][1] if validationLosses else 0
    logging.info("Best validation loss at batch: %d" % best)

    logging.info("Testing on the training and validation dataset ...")
    predictStart = time.time()
    outs = [[], [], [], []]
    ptr = 0
    while ptr < total:
        XBatch, _, endFlag = utils.DecompressArray(XBlocks, ptr, param.predictBatchSize, total)
        for acc, o in zip(outs, m.predict(XBatch)):
            acc.append(o)
        ptr += param.predictBatchSize
        if endFlag != 0:
            break
    bases, zs, ts, ls = [np.concatenate(o) for o in outs]
    logging.info("Prediciton time elapsed: %.2f s" % (time.time() - predictStart))

    YArray, _, _ = utils.DecompressArray(YBlocks, 0, total, total)
    logging.info("Version 2 model, evaluation on base change:")
    order = np.argsort(bases, axis=1, kind="stable")[:, ::-1]
    truth = np.argmax(YArray[:, 0:4], axis=1)
    top1 = int((order[:, 0] == truth).sum())
    top2 = top1 + int((order[:, 1] == truth).sum())
    n = len(bases)
    logging.info("all/top1/top2/top1p/top2p: %d/%d/%d/%.2f/%.2f" % (n, top1, top2, float(top1) / n * 100, float(top2) / n * 100))
    for title, pred, lo, hi in (("Zygosity", zs, 4, 6), ("variant type", ts, 6, 10), ("indel length", ls, 10, 16)):
        logging.info("Version 2 model, evaluation on %s:" % title)
        for row in _confusion(pred, YArray[:, lo:hi], hi - lo):
            logging.info("\t".join(str(v) for v in row))


def main():
    parser = argparse.ArgumentParser(description="Train Clairvoyante")
    D.dataset_options(parser)
    D.optimiser_options(parser)
    D.variant_options(parser)
    Run(D.parse(parser))


if __name__ == "__main__":
    main()
