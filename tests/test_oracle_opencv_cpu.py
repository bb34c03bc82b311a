"""cpu: the oracle's network arithmetic against an INDEPENDENT executor of TensorFlow graph semantics.

TensorFlow 1.12 cannot run here, so the op kernels the oracle restates (conv2d SAME with even kernels, VALID max-pool, the
NHWC flatten order, dense, sigmoid / softmax heads) were "parity unpinned" in round 1: restated from documentation by the
oracle's author.  tests/golden/tf_graphdef.py writes the forward graphs of both variants as TensorFlow GraphDef protobufs (TF's
own proto schemas, shipped with TensorBoard) and OpenCV's dnn module -- which imports TensorFlow graphs and implements their
semantics itself -- executes them; tests/golden/opencv_tf_forward.npz holds its outputs (make_golden_opencv.py).  The oracle
must agree with that fixture, and, when cv2 is importable, with OpenCV executed again now."""
import os
import sys

import numpy as np
import pytest

from clairvoyante_b200 import initializers as I, synth
from oracle import cv_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)


@pytest.fixture(scope="module")
def fx():
    with np.load(os.path.join(GOLD, "opencv_tf_forward.npz")) as z:
        return {k: z[k] for k in z.files}


def _weights(variant, tag):
    if tag == "init":
        return I.init_weights(variant, 4)
    with np.load(os.path.join(GOLD, "trained_%s.npz" % variant)) as z:
        return {k.replace(".", "/"): z[k].astype(np.float32) for k in z.files}


def _other_same_rule(x, kernel, bias):
    """the plausible WRONG reading of SAME for even kernels: the extra row / column of padding at the top / left"""
    kh, kw = kernel.shape[:2]
    n, h, w, cin = x.shape
    pt, pl = (kh - 1) - (kh - 1) // 2, (kw - 1) - (kw - 1) // 2
    xp = np.zeros((n, h + kh - 1, w + kw - 1, cin))
    xp[:, pt:pt + h, pl:pl + w] = x
    out = np.zeros((n, h, w, kernel.shape[3]))
    for i in range(kh):
        for j in range(kw):
            out += xp[:, i:i + h, j:j + w] @ kernel[i, j]
    return out + bias


@pytest.mark.parametrize("i", range(6))
def test_conv_same_and_valid_pool_match_the_tf_importer(fx, i):
    """every conv geometry of both variants (1x4, 2x4, 3x4 with pools 5 / 4 / 3; 1x4, 3x4, 5x4 without): SAME puts the extra
    padding of an even kernel at the bottom / right (clairvoyante_v3.py:54-96); the other convention is far outside the tolerance"""
    k, b, x, y = (fx["layer%d_%s" % (i, s)].astype(np.float64) for s in "kbxy")
    pool = x.shape[1] - y.shape[1] + 1
    ref = O.maxpool_h(O.conv2d_same(x, k, b), pool)
    assert ref.shape == y.shape
    scale = np.abs(y).max()
    assert np.abs(ref - y).max() <= 2e-6 * scale                      # OpenCV accumulates in fp32
    wrong = O.maxpool_h(_other_same_rule(x, k, b), pool)
    assert np.abs(wrong - y).max() > 0.05 * scale                     # the test can tell the two conventions apart


@pytest.mark.parametrize("variant", ["v3", "v3_slim"])
@pytest.mark.parametrize("tag", ["trained", "init"])
def test_forward_matches_the_tf_importer(fx, variant, tag):
    """whole forward pass at phase = False, 64 sites: the four head outputs and their logits (base pre-sigmoid, the others
    SELU + 1e-10) as the oracle computes them == what OpenCV computes from the GraphDef"""
    W = _weights(variant, tag)
    x, _ = synth.make_labeled_sites(int(fx["n"]), int(fx["data_seed"]))
    ref = O.forward(W, x, variant)
    lg, out = fx["%s_%s_logits" % (variant, tag)], fx["%s_%s_out16" % (variant, tag)]
    # OpenCV computes in fp32: its logits carry ~1e-6 relative rounding (|logit| reaches 55 with initialiser weights), and a
    # probability moves by at most a quarter of its logit's error
    tol = 1e-5 * max(1.0, float(np.abs(lg).max()))
    assert np.abs(ref["logits"] - lg).max() <= tol
    assert np.abs(O.out16(ref) - out).max() <= max(5e-6, 0.25 * tol)
    for a, b in ((0, 4), (4, 6), (6, 10), (10, 16)):                  # and the calls are the same wherever the margin is real
        srt = np.sort(lg[:, a:b], 1)
        clear = srt[:, -1] - srt[:, -2] > 1e-4
        assert (ref["logits"][:, a:b].argmax(1) == lg[:, a:b].argmax(1))[clear].all()


@pytest.mark.parametrize("tag,variant", [("v3", "v3"), ("slim", "v3_slim")])
def test_tf_importer_equals_the_references_own_predict(fx, tag, variant):
    """no oracle in this one: reference_graph.npz holds what the REFERENCE'S OWN clairvoyante_v3*.py / selu.py return from
    predict() when executed on the TF stand-in (make_golden_reference_graph.py); OpenCV's TensorFlow importer, given the GraphDef
    of tf_graphdef.py with the same weights and sites, must return the same 16 numbers per site -- the reference's wiring and an
    independent implementation of TensorFlow's kernels meet"""
    with np.load(os.path.join(GOLD, "reference_graph.npz")) as z:
        want = z[tag + "/predict"]
    got = fx["%s_refgraph_out16" % variant]
    assert got.shape == want.shape == (256, 16)
    assert np.abs(got - want).max() <= 2e-5          # fp32 arithmetic in OpenCV, |logit| up to ~60 with initialiser weights
    for a, b in ((4, 6), (6, 10), (10, 16)):         # same call wherever the two leading probabilities are not a tie
        srt = np.sort(want[:, a:b], 1)               # (initialiser weights leave most v3 zygosity logits at SELU's floor: 0.5 / 0.5)
        clear = srt[:, -1] - srt[:, -2] > 1e-4
        assert (got[:, a:b].argmax(1) == want[:, a:b].argmax(1))[clear].all()


def test_fixture_is_what_opencv_computes_now(fx):
    """the committed fixture is reproducible: re-run the importer (skipped where cv2 is absent)"""
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "dnn"):
        pytest.skip("cv2 without dnn")
    import make_golden_opencv as M
    from tf_graphdef import forward_graph
    x, _ = synth.make_labeled_sites(int(fx["n"]), int(fx["data_seed"]))
    for variant in ("v3", "v3_slim"):
        pb, outs, logits = forward_graph(_weights(variant, "trained"), variant, len(x))
        ys = M.run_pb(pb, x, list(outs) + list(logits))
        got = np.concatenate([y.reshape(len(x), -1) for y in ys[:4]], 1)
        assert np.abs(got - fx["%s_trained_out16" % variant]).max() <= 5e-6      # (OpenCV dispatches on the host's vector ISA)
