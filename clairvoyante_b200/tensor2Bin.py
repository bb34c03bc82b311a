"""tensor2Bin -- text tensors (+ truth, BED) -> binary training file; counterpart of reference
clairvoyante/tensor2Bin.py:18-28: four consecutive pickles (total, X blocks, Y blocks, pos blocks).  Blocks are Blosc-1 / LZ4 frames around Python-2-style pickles (utils_v2.pack_array; unshuffled: smaller and faster to decode
for count tensors); `--blosc` additionally applies python-blosc's default byte shuffle and writes protocol-2 outer pickles,
i.e. what the reference itself writes, so that the file also loads in the reference's train.py (:40-44)."""
import argparse
import logging
import pickle

from . import _driver as D, param

logging.basicConfig(format='%(message)s', level=logging.INFO)


def Run(args):
    from . import utils_v2 as utils
    utils.SetupEnv()
    logging.info("Loading the dataset ...")
    blosc = bool(getattr(args, "blosc", False))
    parts = utils.GetTrainingArray(args.tensor_fn, args.var_fn, args.bed_fn, container="blosc" if blosc else "cvbz")
    logging.info("Writing to binary ...")
    with open(args.bin_fn, "wb") as fh:
        for p in parts:
            if blosc:
                pickle.dump(p, fh, protocol=2)       # loadable by Python 2 (bytes travel as latin-1 text through _codecs.encode)
            else:
                pickle.dump(p, fh)


def main():
    parser = argparse.ArgumentParser(description="Generate a binary format input tensor")
    D.dataset_options(parser)                    # here --bin_fn names the OUTPUT
    parser.add_argument('--blosc', type=param.str2bool, nargs='?', const=True, default=False,
                        help="Shuffled frames + protocol-2 outer pickles, exactly like the reference (default: unshuffled frames)")
    D.variant_options(parser)
    Run(D.parse(parser))


if __name__ == "__main__":
    main()
