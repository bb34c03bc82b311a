"""PairWithNonVariants -- truth-variant tensors + a random sample of non-variant tensors -> one training tensor file;
counterpart of reference dataPrepScripts/PairWithNonVariants.py (same options, :138-153, plus --seed).  Every variant row is
kept (:99-106); a candidate row that lies inside the BED regions and not at a variant position is kept with probability
min(1, amp * #variants / #usable candidates) (:88-90, :107-123)."""
import argparse
import gzip
import logging
import random

from . import _driver as D
from .utils_v2 import _Regions

logging.basicConfig(format='%(message)s', level=logging.INFO)


def _text(fn):
    """a text file, gzip or plain (what `gzip -fdc` accepts)"""
    with open(fn, "rb") as probe:
        packed = probe.read(2) == b"\x1f\x8b"
    return gzip.open(fn, "rt") if packed else open(fn, "rt")


def _regions(bed_fn):
    if bed_fn is None:
        return None
    logging.info("Loading BED file ...")
    regions = _Regions()
    with _text(bed_fn) as f:
        for fields in (line.split() for line in f):
            if len(fields) >= 3:
                lo, hi = int(fields[1]), int(fields[2]) - 1      # :38-41: intervaltree's half-open [lo, hi), never empty
                regions.add(fields[0], lo, hi + 1 if hi == lo else hi)
    return regions


def _sites(fn):
    """(stripped row, (contig, position)) of every tensor row"""
    with _text(fn) as f:
        for line in f:
            row = line.strip()
            head = row.split(None, 2)
            if len(head) >= 2:
                yield row, (head[0], int(head[1]))


def Pair(args):
    regions = _regions(args.bed_fn)

    def eligible(site):
        return (regions is None or (site[0] in regions and regions.hit(*site))) and site not in variants

    logging.info("Counting the number of Truth Variants in %s ..." % args.tensor_var_fn)
    variants, n_variant_rows = set(), 0
    for _, site in _sites(args.tensor_var_fn):
        variants.add(site)
        n_variant_rows += 1
    logging.info("%d Truth Variants" % n_variant_rows)
    wanted = n_variant_rows * args.amp
    logging.info("%d non-variants to be picked" % wanted)
    logging.info("Counting the number of usable non-variants in %s ..." % args.tensor_can_fn)
    usable = sum(1 for _, site in _sites(args.tensor_can_fn) if eligible(site))
    logging.info("%d usable non-variant" % usable)
    keep = min(1.0, float(wanted) / usable) if usable else 1.0
    logging.info("%.2f of all non-variants are selected" % keep)
    draw = (random.Random(args.seed) if getattr(args, "seed", None) is not None else random).random
    written = [0, 0]
    with gzip.open(args.output_fn, "wt") as out:
        with _text(args.tensor_var_fn) as f:
            for line in f:
                out.write(line.strip() + "\n")
                written[0] += 1
        for row, site in _sites(args.tensor_can_fn):
            if eligible(site) and draw() < keep:
                out.write(row + "\n")
                written[1] += 1
    logging.info("%.2f/%.2f Truth Variants/Non-variants outputed" % tuple(written))
    return tuple(written)


def main():
    parser = argparse.ArgumentParser(description="Pair truth-variant tensors with sampled non-variant tensors")
    parser.add_argument('--tensor_can_fn', type=str, default=None, help="Tensors at candidate (mostly non-variant) positions")
    parser.add_argument('--tensor_var_fn', type=str, default=None, help="Tensors at the truth variants")
    parser.add_argument('--bed_fn', type=str, default=None, help="Regions a non-variant may come from (BED)")
    parser.add_argument('--output_fn', type=str, default=None, help="Output tensor file (gzip)")
    parser.add_argument('--amp', type=float, default=2, help="Non-variants to pick per truth variant, default: 2")
    parser.add_argument('--seed', type=int, default=None, help="Seed of the sampler (the reference samples unseeded)")
    Pair(D.parse(parser))


if __name__ == "__main__":
    main()
