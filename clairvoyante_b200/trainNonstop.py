"""trainNonstop -- train at a fixed learning rate, one checkpoint per epoch, until maxEpoch; counterpart of reference
clairvoyante/trainNonstop.py (same options, :131-170): train.py's epoch (first 90 % trains, the rest validates; :78-124)
without the learning-rate schedule.  `--ochk_prefix` is mandatory (:31-32)."""
import argparse
import logging
import sys
import time

from . import _driver as D, param

logging.basicConfig(format='%(message)s', level=logging.INFO)


def Run(args):
    logging.info("Initializing model ...")
    m, utils = D.new_model(args)
    if args.ochk_prefix is None:
        sys.exit("--chk_prefix must be defined in nonstop training mode")
    if args.chkpnt_fn is not None:
        m.restoreParameters(D.absolute(args.chkpnt_fn))
    TrainAll(args, m, utils)


def TrainAll(args, m, utils):
    logging.info("Loading the training dataset ...")
    data = D.TrainingSet(args, utils)
    writer = D.announce_training(args, m, data)
    began = time.time()
    for epoch in range(D.first_epoch(args), param.maxEpoch):
        D.train_validate_epoch(m, data, epoch, writer)
        m.saveParameters(D.absolute(D.checkpoint_name(args.ochk_prefix, epoch)))
    logging.info("Training time elapsed: %.2f s" % (time.time() - began))


def main():
    parser = argparse.ArgumentParser(description="Train Clairvoyante Nonstop")
    D.dataset_options(parser)
    D.optimiser_options(parser)
    D.variant_options(parser)
    Run(D.parse(parser))


if __name__ == "__main__":
    main()
