"""calTrainDevDiff -- mean training-set and validation-set loss of a list of checkpoints; counterpart of reference
clairvoyante/calTrainDevDiff.py (same options, :80-110).  One getLossNoRT per batch of predictBatchSize, overlapped with the
next fetch (:55-72).  Kept from the reference: a batch's loss counts as "validation" when the reader -- which already
stands behind the batch being evaluated -- has reached validationStart (:70-71); the batch before the split is shortened to
end on it, the batches behind it are aligned to multiples of predictBatchSize (:61-64); the last batch is evaluated after
the loop (:73-76)."""
import argparse
import sys

from . import _driver as D, param


def Run(args):
    m, utils = D.new_model(args)
    CalcAll(args, m, utils)


def CalcAll(args, m, utils):
    data = D.TrainingSet(args, utils)

    def rows_after(at):
        size = param.predictBatchSize
        if at < data.validationStart and data.validationStart - at < size:
            return data.validationStart - at
        if at >= data.validationStart and at % size != 0:
            return size - at % size
        return size

    results = []
    for name in args.chkpnt_fn:
        m.restoreParameters(D.absolute(name))
        sums = {True: 0, False: 0}                 # keyed by "counts as validation"
        w = D.Walk(data, param.predictBatchSize, rows_after, lambda at: m.getLossNoRT)
        for at, _method in w:
            sums[at >= data.validationStart] += m.getLossLossRTVal
        X, Y, at = w.tail
        m.getLossNoRT(X, Y)
        # (the reference tests the pointer BEFORE it moves past the last batch, :75)
        last_at = at - len(X)
        sums[last_at >= data.validationStart] += m.getLossLossRTVal
        line = (name, sums[False] / data.trainingTotal, sums[True] / data.numValItems)
        print("%s\t%.10f\t%.10f" % line, file=sys.stderr)
        results.append(line)
    return results


def main():
    parser = argparse.ArgumentParser(description="Calculate the loss different between training dataset and validation dataset")
    D.dataset_options(parser)
    parser.add_argument('--chkpnt_fn', nargs='+', type=str, default=None, help="Checkpoints to evaluate")
    D.variant_options(parser)
    Run(D.parse(parser))


if __name__ == "__main__":
    main()
