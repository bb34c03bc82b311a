"""Per-head argmax parity on TRAINED-LIKE weights (north_star: "per-head argmax bit-exact (fp logits within 1e-3)").

With initialiser weights 92 % of the v3 sites have both zygosity logits at SELU's floor, so an "exempt where the oracle is
tied" rule exempted nearly every site of that head.  tests/golden/trained_{v3,v3_slim}.npz are weights after a few hundred
Adam steps of the oracle on labelled synthetic sites (tests/golden/make_trained_weights.py); the fixture
trained_forward_*.npz holds the oracle's logits / outputs for 4,096 sites and ENUMERATES the near-tie sites per head
(top-2 margin <= 2e-3; < 1 % on every head).  The bar: argmax identical on EVERY site outside those lists, at the reference's
predictBatchSize (callVar.py:184, param.py:12) and as one batch, on every arithmetic path and every input feed.
"""
import os

import numpy as np
import pytest

from clairvoyante_b200 import synth
from oracle import cv_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HEADS = ((0, 4), (4, 6), (6, 10), (10, 16))
HEAD_NAMES = ("base", "zygosity", "varType", "indelLength")


def load_trained(variant):
    with np.load(os.path.join(GOLD, "trained_%s.npz" % variant)) as z:
        W = {k.replace(".", "/"): z[k].astype(np.float32) for k in z.files}
    with np.load(os.path.join(GOLD, "trained_forward_%s.npz" % variant)) as z:
        fx = {k: z[k] for k in z.files}
    x, y = synth.make_labeled_sites(int(fx["n"]), int(fx["data_seed"]))
    return W, fx, x, y


def exempt_fractions(fx):
    return [len(fx["ties_%d" % h]) / float(fx["n"]) for h in range(4)]


@pytest.mark.parametrize("variant", ["v3", "v3_slim"])
def test_fixture_is_the_oracle_and_ties_are_rare(variant):
    """cpu: the committed logits are what the NumPy oracle computes from the committed weights, the tie lists are exactly the
    sites under the margin, and they are < 1 % of the sites on every head"""
    W, fx, x, y = load_trained(variant)
    idx = np.r_[0:48, 2000:2016]
    ref = O.forward(W, x[idx], variant)
    assert np.abs(ref["logits"] - fx["logits"][idx]).max() < 1e-9
    assert np.abs(O.out16(ref) - fx["out16"][idx]).max() < 1e-12
    for h, (a, b) in enumerate(HEADS):
        srt = np.sort(fx["logits"][:, a:b], 1)
        assert np.array_equal(np.nonzero(srt[:, -1] - srt[:, -2] <= float(fx["margin"]))[0], fx["ties_%d" % h])
        assert np.array_equal(fx["logits"][:, a:b].argmax(1), fx["argmax"][:, h])
        assert len(fx["ties_%d" % h]) < 0.01 * int(fx["n"]), HEAD_NAMES[h]
    # the network has learnt the implanted genotypes: this is not a degenerate constant predictor
    assert (fx["argmax"][:, 1] == y[:, 4:6].argmax(1)).mean() > 0.9
    assert (fx["argmax"][:, 2] == y[:, 6:10].argmax(1)).mean() > 0.9


def check_against_fixture(lg, out16, fx, tol=1e-3):
    """logits within tol, outputs within 2e-4, argmax identical outside the enumerated near-ties; returns #sites compared"""
    assert np.abs(lg - fx["logits"]).max() <= tol, "max |logit - oracle| = %g" % np.abs(lg - fx["logits"]).max()
    assert np.abs(out16 - fx["out16"]).max() <= 2e-4
    n = int(fx["n"])
    for h, (a, b) in enumerate(HEADS):
        keep = np.ones(n, bool)
        keep[fx["ties_%d" % h]] = False
        bad = np.nonzero((lg[:, a:b].argmax(1) != fx["argmax"][:, h]) & keep)[0]
        assert len(bad) == 0, "%s head: argmax differs at sites %s" % (HEAD_NAMES[h], bad[:10])
        pbad = np.nonzero((out16[:, a:b].argmax(1) != fx["argmax"][:, h]) & keep)[0]       # what callVar.py:58-66 takes
        if h > 0:                           # (the base head's sigmoid saturates: equal probabilities are not a logit tie)
            assert len(pbad) == 0, "%s head: argmax of the probabilities differs at sites %s" % (HEAD_NAMES[h], pbad[:10])


@pytest.mark.gpu
def test_fp16_mode_on_trained_weights():
    """BASELINE configs[2] (v3_slim, plain fp16): logits within the mode's stated tolerance 2e-3 * max(1, max |logit|) of the
    fp64 oracle, and the per-head argmax identical on every site whose oracle top-2 margin exceeds twice that tolerance --
    the sites under it are counted and must stay a small minority"""
    from clairvoyante_b200 import clairvoyante_v3_slim as cv
    W, fx, x, y = load_trained("v3_slim")
    m = cv.Clairvoyante()
    m.setComputeMode("fp16")
    m.setWeights(W)
    out16, lg = m.predictLogits(x)
    tol = 2e-3 * max(1.0, float(np.abs(fx["logits"]).max()))
    err = float(np.abs(lg - fx["logits"]).max())
    assert err <= tol, "max |logit - oracle| = %g (tolerance %g)" % (err, tol)
    exempt = []
    for h, (a, b) in enumerate(HEADS):
        srt = np.sort(fx["logits"][:, a:b], 1)
        clear = (srt[:, -1] - srt[:, -2]) > 2 * tol
        assert (lg[:, a:b].argmax(1) == fx["argmax"][:, h])[clear].all(), HEAD_NAMES[h]
        exempt.append(1.0 - clear.mean())
    print("v3_slim/fp16: max |logit - oracle| = %.3g (tolerance %.3g), exempt fraction per head %s, label agreement %s"
          % (err, tol, ["%.4f" % e for e in exempt],
             ["%.4f" % (lg[:, a:b].argmax(1) == fx["argmax"][:, h]).mean() for h, (a, b) in enumerate(HEADS)]))
    assert max(exempt[1:]) < 0.05
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant,mode", [("v3", "fp16x3"), ("v3", "fp32"), ("v3_slim", "fp16x3"), ("v3_slim", "fp32")])
def test_argmax_identical_on_trained_weights(variant, mode):
    from clairvoyante_b200 import utils_v2 as U
    if variant == "v3":
        from clairvoyante_b200 import clairvoyante_v3 as cv
    else:
        from clairvoyante_b200 import clairvoyante_v3_slim as cv
    W, fx, x, y = load_trained(variant)
    m = cv.Clairvoyante()
    m.setComputeMode(mode)
    m.setWeights(W)
    out16, lg = m.predictLogits(x)                                   # one batch of 4,096
    check_against_fixture(lg, out16, fx)
    bs = 1000                                                        # the reference's predictBatchSize
    parts = [m.predictLogits(x[i:i + bs]) for i in range(0, len(x), bs)]
    o2, l2 = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    assert np.array_equal(o2, out16) and np.array_equal(l2, lg)      # batching does not change a bit
    # narrow feeds: raw uint8 / int16 counts and fp16 values give the fp32 feed's bits
    cnt = U.pack_counts(x)
    assert cnt is not None and cnt.dtype == np.uint8
    for feed in (cnt, cnt.astype(np.int16), x.astype(np.float16), U.with_counts(x)):
        o3, l3 = m.predictLogits(feed)
        assert np.array_equal(o3, out16) and np.array_equal(l3, lg), getattr(feed, "dtype", None)
    base, z, t, l = m.predict(U.with_counts(x))
    assert np.array_equal(np.concatenate([base, z, t, l], 1), out16)
    print("%s/%s: exempt (near-tie) fraction per head: %s" % (variant, mode, ["%.4f" % f for f in exempt_fractions(fx)]))
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant,mode", [("v3", "fp16x3"), ("v3", "fp32"), ("v3_slim", "fp16x3"), ("v3_slim", "fp32")])
def test_cuda_matches_the_independent_tf_executor(variant, mode):
    """The CUDA path against tests/golden/opencv_tf_forward.npz -- the forward GraphDef executed by OpenCV's TensorFlow
    importer (tests/golden/tf_graphdef.py, make_golden_opencv.py), an implementation of TF's op semantics written by neither the
    reference's authors nor this repo's: logits within north_star's 1e-3, outputs within 2e-4, argmax identical wherever the
    fixture's top-2 margin exceeds 2e-3 -- with no oracle in the loop"""
    from clairvoyante_b200 import initializers as I
    if variant == "v3":
        from clairvoyante_b200 import clairvoyante_v3 as cv
    else:
        from clairvoyante_b200 import clairvoyante_v3_slim as cv
    with np.load(os.path.join(GOLD, "opencv_tf_forward.npz")) as z:
        fx = {k: z[k] for k in z.files}
    x, _ = synth.make_labeled_sites(int(fx["n"]), int(fx["data_seed"]))
    m = cv.Clairvoyante()
    m.setComputeMode(mode)
    for tag in ("trained", "init"):
        W = load_trained(variant)[0] if tag == "trained" else I.init_weights(variant, 4)
        m.setWeights(W)
        out16, lg = m.predictLogits(x)
        want_lg, want = fx["%s_%s_logits" % (variant, tag)], fx["%s_%s_out16" % (variant, tag)]
        assert np.abs(lg - want_lg).max() <= 1e-3, (tag, float(np.abs(lg - want_lg).max()))
        assert np.abs(out16 - want).max() <= 2e-4, tag
        for a, b in HEADS:
            srt = np.sort(want_lg[:, a:b], 1)
            clear = srt[:, -1] - srt[:, -2] > 2e-3
            assert (lg[:, a:b].argmax(1) == want_lg[:, a:b].argmax(1))[clear].all(), (tag, a)
    m.close()
