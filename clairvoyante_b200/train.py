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
import os
import sys
import time
from threading import Thread

import numpy as np

from . import param

logging.basicConfig(format='%(message)s', level=logging.INFO)


def Run(args):
    if args.v2:
        sys.exit("clairvoyante_b200 implements the v3 / v3_slim networks only (--v2 is out of scope)")
    from . import utils_v2 as utils
    if args.slim:
        from . import clairvoyante_v3_slim as cv
    else:
        from . import clairvoyante_v3 as cv
    utils.SetupEnv()
    m = cv.Clairvoyante()
    m.init()
    if args.chkpnt_fn is not None:
        m.restoreParameters(os.path.abspath(args.chkpnt_fn))
    TrainAll(args, m, utils)


def next_batch_size(ptr, validationStart):
    """rows to fetch at dataset position `ptr` (train.py:95-102)"""
    if ptr < validationStart:
        return min(param.trainBatchSize, validationStart - ptr) if (validationStart - ptr) < param.trainBatchSize else param.trainBatchSize
    if ptr % param.predictBatchSize != 0:
        return param.predictBatchSize - (ptr % param.predictBatchSize)
    return param.predictBatchSize


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
    if args.bin_fn is not None:
        total, XBlocks, YBlocks, posBlocks = utils.load_bin(args.bin_fn)       # noqa: F841  (positions kept for format parity)
    else:
        total, XBlocks, YBlocks, posBlocks = utils.GetTrainingArray(args.tensor_fn, args.var_fn, args.bed_fn)
    logging.info("The size of training dataset: {}".format(total))

    summaryWriter = m.summaryFileWriter(args.olog_dir) if args.olog_dir is not None else None

    logging.info("Start training ...")
    logging.info("Learning rate: %.2e" % m.setLearningRate(args.learning_rate))
    logging.info("L2 regularization lambda: %.2e" % m.setL2RegularizationLambda(args.lambd))

    validationLosses = []
    trainingStart = time.time()
    trainingTotal = int(total * param.trainingDatasetPercentage)
    validationStart = trainingTotal + 1
    numValItems = total - validationStart
    switchesLeft = param.maxLearningRateSwitch
    sinceSwitch = 0
    epoch = 1 if args.chkpnt_fn is None else int(args.chkpnt_fn[-param.parameterOutputPlaceHolder:]) + 1

    def fetch(ptr, size):
        X, nx, ex = utils.DecompressArray(XBlocks, ptr, size, total)
        Y, ny, ey = utils.DecompressArray(YBlocks, ptr, size, total)
        if nx != ny or ex != ey:
            sys.exit("Inconsistency between decompressed arrays: %d/%d" % (nx, ny))
        return X, Y, nx, ex

    while epoch < param.maxEpoch:
        epochStart = time.time()
        trainLossSum = validationLossSum = 0
        XBatch, YBatch, got, _ = fetch(0, param.trainBatchSize)
        ptr = got
        while True:
            training = ptr < validationStart
            worker = Thread(target=m.trainNoRT if training else m.getLossNoRT, args=(XBatch, YBatch))
            worker.start()
            XNext, YNext, got, endFlag = fetch(ptr, next_batch_size(ptr, validationStart))   # overlaps the model call
            worker.join()
            XBatch, YBatch = XNext, YNext
            if training:
                trainLossSum += m.trainLossRTVal
                if summaryWriter is not None:
                    summaryWriter.add_summary(m.trainSummaryRTVal, epoch)
            else:
                validationLossSum += m.getLossLossRTVal
            ptr += got
            if endFlag != 0:
                break
        validationLossSum += m.getLoss(XBatch, YBatch)
        logging.info(" ".join([str(epoch), "Training loss:", str(trainLossSum / trainingTotal), "Validation loss: ",
                               str(validationLossSum / numValItems)]))
        logging.info("Epoch time elapsed: %.2f s" % (time.time() - epochStart))
        validationLosses.append((validationLossSum, epoch))
        if args.ochk_prefix is not None:
            path = "%s-%%0%dd" % (args.ochk_prefix, param.parameterOutputPlaceHolder)
            m.saveParameters(os.path.abspath(path % epoch))
        sinceSwitch += 1
        if sinceSwitch >= 6 and switch_needed([v for v, _ in validationLosses]):
            switchesLeft -= 1
            if switchesLeft == 0:
                break
            logging.info("New learning rate: %.2e" % m.setLearningRate())
            logging.info("New L2 regularization lambda: %.2e" % m.setL2RegularizationLambda())
            sinceSwitch = 0
        epoch += 1

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
    parser.add_argument('--bin_fn', type=str, default=None,
                        help="Binary tensor input generated by tensor2Bin.py, tensor_fn, var_fn and bed_fn will be ignored")
    parser.add_argument('--tensor_fn', type=str, default="vartensors", help="Tensor input")
    parser.add_argument('--var_fn', type=str, default="truthvars", help="Truth variants list input")
    parser.add_argument('--bed_fn', type=str, default=None, help="High confident genome regions input in the BED format")
    parser.add_argument('--chkpnt_fn', type=str, default=None, help="Input a checkpoint for testing or continue training")
    parser.add_argument('--learning_rate', type=float, default=param.initialLearningRate,
                        help="Set the initial learning rate, default: %(default)s")
    parser.add_argument('--lambd', type=float, default=param.l2RegularizationLambda,
                        help="Set the l2 regularization lambda, default: %(default)s")
    parser.add_argument('--ochk_prefix', type=str, default=None, help="Prefix for checkpoint outputs at each learning rate change, optional")
    parser.add_argument('--olog_dir', type=str, default=None, help="Directory for tensorboard log outputs, optional")
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
