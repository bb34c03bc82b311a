"""trainWithoutValidationNonstop -- every row trains, fixed learning rate, one checkpoint per epoch until maxEpoch;
counterpart of reference clairvoyante/trainWithoutValidationNonstop.py (same options).  Kept from the reference: all batches
are trainBatchSize rows, the LAST batch of an epoch is fetched but never trained on (the loop ends when the reader hits the
end, :95-104), the epoch mean divides by the whole set, and the checkpoint path is used as given (not made absolute)."""
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
        epoch_began = time.time()
        loss_sum = 0
        for _at, _method in D.Walk(data, param.trainBatchSize, lambda at: param.trainBatchSize, lambda at: m.trainNoRT):
            loss_sum += m.trainLossRTVal
            if writer is not None:
                writer.add_summary(m.trainSummaryRTVal, epoch)
        logging.info(" ".join([str(epoch), "Training loss:", str(loss_sum / data.total)]))
        logging.info("Epoch time elapsed: %.2f s" % (time.time() - epoch_began))
        m.saveParameters(D.checkpoint_name(args.ochk_prefix, epoch))
    logging.info("Training time elapsed: %.2f s" % (time.time() - began))


def main():
    parser = argparse.ArgumentParser(description="Train Clairvoyante Nonstop without validation")
    D.dataset_options(parser)
    D.optimiser_options(parser)
    D.variant_options(parser)
    Run(D.parse(parser))


if __name__ == "__main__":
    main()
