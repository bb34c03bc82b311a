"""What the training / evaluation drivers share: command-line option groups, model construction, the training-set view
with the reference's 90 / 10 split, and the epoch walk in which the model works on batch k (on a thread) while batch k+1 is
decompressed.  The drivers (train, trainNonstop, trainWithoutValidationNonstop, evaluate, calTrainDevDiff) keep the
reference's command lines, log lines and model-call sequences -- tests/test_reference_run_cpu.py holds them to what the
reference's own drivers did with a stub model -- but are written on top of this module instead of five copies of one loop."""
import logging
import os

import numpy as np
import sys
import time
from threading import Thread

from . import param

V2_MESSAGE = "clairvoyante_b200 implements the v3 / v3_slim networks only (--v2 is out of scope)"


# ---------------------------------------------------------------------------------------------------- command line
def dataset_options(parser):
    parser.add_argument('--bin_fn', type=str, default=None, help="Training set written by tensor2Bin (then the next three are ignored)")
    parser.add_argument('--tensor_fn', type=str, default="vartensors", help="Tensor text input")
    parser.add_argument('--var_fn', type=str, default="truthvars", help="Truth variant list")
    parser.add_argument('--bed_fn', type=str, default=None, help="High-confidence regions (BED)")


def variant_options(parser):
    parser.add_argument('--v3', type=param.str2bool, nargs='?', const=True, default=True, help="v3 network (default)")
    parser.add_argument('--v2', type=param.str2bool, nargs='?', const=True, default=False, help="v2 network (not available here)")
    parser.add_argument('--slim', type=param.str2bool, nargs='?', const=True, default=False, help="slim variant of the network")


def optimiser_options(parser):
    parser.add_argument('--chkpnt_fn', type=str, default=None, help="Checkpoint to continue from")
    parser.add_argument('--learning_rate', type=float, default=param.initialLearningRate, help="Initial learning rate, default: %(default)s")
    parser.add_argument('--lambd', type=float, default=param.l2RegularizationLambda, help="L2 regularisation lambda, default: %(default)s")
    parser.add_argument('--ochk_prefix', type=str, default=None, help="Prefix of the checkpoints written after every epoch")
    parser.add_argument('--olog_dir', type=str, default=None, help="Directory for summary logs, optional")


def parse(parser):
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    return args


def new_model(args):
    """(model, utils) for the variant the options select; the model is initialised, nothing is restored"""
    if args.v2:
        sys.exit(V2_MESSAGE)
    from . import utils_v2 as utils
    if args.slim:
        from . import clairvoyante_v3_slim as cv
    else:
        from . import clairvoyante_v3 as cv
    utils.SetupEnv()
    m = cv.Clairvoyante()
    m.init()
    return m, utils


# ---------------------------------------------------------------------------------------------------- training set
class TrainingSet(object):
    """blocks of a training set + the split the reference uses: rows [0, trainingTotal] train, the rest validates; the
    validation mean divides by total - validationStart (train.py:69-72)"""

    def __init__(self, args, utils):
        if args.bin_fn is not None:
            self.total, self.X, self.Y, _ = utils.load_bin(args.bin_fn)
        else:
            self.total, self.X, self.Y, _ = utils.GetTrainingArray(args.tensor_fn, args.var_fn, args.bed_fn)
        self.utils = utils
        self.trainingTotal = int(self.total * param.trainingDatasetPercentage)
        self.validationStart = self.trainingTotal + 1
        self.numValItems = self.total - self.validationStart

    def fetch(self, at, rows):
        """(X, Y, rows delivered, end flag) for rows [at, at + rows)"""
        X, nx, ex = self.utils.DecompressArray(self.X, at, rows, self.total)
        Y, ny, ey = self.utils.DecompressArray(self.Y, at, rows, self.total)
        if nx != ny or ex != ey:
            sys.exit("Inconsistency between decompressed arrays: %d/%d" % (nx, ny))
        # the batch also carries its raw counts (uint8 / int16) for the narrow host->device feed of train / getLoss
        # (cvb_train_step_host_x); packed natively on the reader's side of the train / fetch overlap
        pack = getattr(self.utils, "with_counts", None)
        if pack is not None and nx and os.environ.get("CVB_FEED", "counts") != "fp32" and getattr(X, "dtype", None) == np.float32:
            X = pack(X)
        return X, Y, nx, ex

    def rows_wanted(self, at):
        return rows_wanted(at, self.validationStart)


def rows_wanted(at, validationStart):
    """size of the batch that starts at row `at` when an epoch trains and validates (train.py:95-102): full training
    batches up to the split, then prediction-sized batches aligned to multiples of predictBatchSize"""
    if at < validationStart:
        return min(param.trainBatchSize, validationStart - at)
    off = at % param.predictBatchSize
    return param.predictBatchSize - off if off else param.predictBatchSize


class Walk(object):
    """One pass over the set.  `call_for(at)` names the model method for the batch that is in hand when the reader stands at
    row `at`; that call runs on a thread while the following batch (rows_after(at) rows) is fetched.  Iterating yields
    (at, called method) after every call; afterwards `.tail` = (X, Y, at) holds the LAST fetched batch, which no call has seen."""

    def __init__(self, data, first_rows, rows_after, call_for):
        self.data, self.first_rows, self.rows_after, self.call_for, self.tail = data, first_rows, rows_after, call_for, None

    def __iter__(self):
        X, Y, got, _ = self.data.fetch(0, self.first_rows)
        at = got
        while True:
            method = self.call_for(at)
            worker = Thread(target=method, args=(X, Y))
            worker.start()
            X2, Y2, got, end = self.data.fetch(at, self.rows_after(at))
            worker.join()
            X, Y = X2, Y2
            yield at, method
            at += got
            if end != 0:
                break
        self.tail = (X, Y, at)


def checkpoint_name(prefix, epoch):
    return ("%s-%%0%dd" % (prefix, param.parameterOutputPlaceHolder)) % epoch


def first_epoch(args):
    return 1 if args.chkpnt_fn is None else int(args.chkpnt_fn[-param.parameterOutputPlaceHolder:]) + 1


def announce_training(args, m, data):
    logging.info("The size of training dataset: {}".format(data.total))
    writer = m.summaryFileWriter(args.olog_dir) if args.olog_dir is not None else None
    logging.info("Start training ...")
    logging.info("Learning rate: %.2e" % m.setLearningRate(args.learning_rate))
    logging.info("L2 regularization lambda: %.2e" % m.setL2RegularizationLambda(args.lambd))
    return writer


def train_validate_epoch(m, data, epoch, writer):
    """the epoch of train.py / trainNonstop.py (:86-122): returns (training loss sum, validation loss sum) and logs the epoch line"""
    began = time.time()
    train_sum = val_sum = 0
    w = Walk(data, param.trainBatchSize, data.rows_wanted, lambda at: m.trainNoRT if at < data.validationStart else m.getLossNoRT)
    for at, method in w:
        if method == m.trainNoRT:
            train_sum += m.trainLossRTVal
            if writer is not None:
                writer.add_summary(m.trainSummaryRTVal, epoch)
        else:
            val_sum += m.getLossLossRTVal
    X, Y, _ = w.tail
    val_sum += m.getLoss(X, Y)
    logging.info(" ".join([str(epoch), "Training loss:", str(train_sum / data.trainingTotal), "Validation loss: ",
                           str(val_sum / data.numValItems)]))
    logging.info("Epoch time elapsed: %.2f s" % (time.time() - began))
    return train_sum, val_sum


def absolute(path):
    return os.path.abspath(path)
