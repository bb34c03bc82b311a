"""tensor2Bin -- text tensors (+ truth, BED) -> binary training file; counterpart of reference
clairvoyante/tensor2Bin.py:18-28: four consecutive pickles (total, X blocks, Y blocks, pos blocks).  Blocks are Blosc-1 / LZ4 frames around Python-2-style pickles (utils_v2.pack_array; unshuffled: smaller and faster to decode
for count tensors); `--blosc` additionally applies python-blosc's default byte shuffle and writes protocol-2 outer pickles,
i.e. what the reference itself writes, so that the file also loads in the reference's train.py (:40-44)."""
import argparse
import logging
import pickle
import sys

from . import param

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
    parser.add_argument('--tensor_fn', type=str, default="vartensors", help="Tensor input")
    parser.add_argument('--var_fn', type=str, default="truthvars", help="Truth variants list input")
    parser.add_argument('--bed_fn', type=str, default=None, help="High confident genome regions input in the BED format")
    parser.add_argument('--bin_fn', type=str, default=None, help="Output a binary tensor file")
    parser.add_argument('--blosc', type=param.str2bool, nargs='?', const=True, default=False,
                        help="Shuffled frames + protocol-2 outer pickles, exactly like the reference (default: unshuffled frames)")
    parser.add_argument('--v3', type=param.str2bool, nargs='?', const=True, default=True, help="Use Clairvoyante version 3")
    parser.add_argument('--v2', type=param.str2bool, nargs='?', const=True, default=False, help="Use Clairvoyante version 2")
    args = parser.parse_args()
    if len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit(1)
    Run(args)


if __name__ == "__main__":
    main()
