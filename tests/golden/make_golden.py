"""Generates tests/golden/forward_{v3,v3_slim}.npz from the NumPy float64 oracle.
The reference cannot run here (TensorFlow 1.12 absent, Python 2 sources), so these are
ORACLE outputs on seeded weights/inputs, not reference outputs: they pin the oracle and the
synthetic generator against accidental change and give the GPU tests a fixed target.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from clairvoyante_b200 import initializers as I, synth  # noqa: E402
from oracle import cv_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
for variant, wseed, dseed, n in (("v3", 11, 12, 96), ("v3_slim", 21, 22, 96)):
    W = I.init_weights(variant, wseed)
    x = synth.make_sites(n, dseed)
    o = O.forward(W, x, variant)
    np.savez_compressed(os.path.join(HERE, "forward_%s.npz" % variant), weight_seed=wseed, data_seed=dseed, n=n,
                        x_head=x[:4], logits=o["logits"], out16=O.out16(o))
    print(variant, "written")
