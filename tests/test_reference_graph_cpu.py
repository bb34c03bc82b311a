"""The network oracles against the reference's own graph code: tests/golden/reference_graph.npz was produced by executing
clairvoyante_v3.py / clairvoyante_v3_slim.py / selu.py of the reference on a TensorFlow-1.x stand-in
(tests/golden/make_golden_reference_graph.py, tests/golden/tf1_stand_in.py): wiring, sizes, formulas and feeds are the
reference's, the op kernels are the stand-in's (torch float64).  Both oracle restatements must reproduce it to rounding."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clairvoyante_b200 import initializers   # noqa: E402
from oracle import cv_oracle as O, cv_oracle_torch as OT   # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_graph.npz"))
X = G["x"].astype(np.float32)
Y = G["y"]
VARIANTS = [("v3", "v3"), ("slim", "v3_slim")]


def _weights(v):
    return {k: np.asarray(a, np.float64) for k, a in initializers.init_weights(v, seed=0).items()}


@pytest.mark.parametrize("tag,variant", VARIANTS)
def test_parameter_count(tag, variant):
    assert int(G[tag + "/n_params"]) == sum(int(np.prod(s)) for _, s in O.variable_shapes(variant))


@pytest.mark.parametrize("tag,variant", VARIANTS)
def test_predict_and_getloss(tag, variant):
    W = _weights(variant)
    out = O.out16(O.forward(W, X, variant))
    assert np.abs(out - G[tag + "/predict"]).max() < 1e-12
    assert abs(O.loss(W, X, Y, variant, l2_lambda=0.0)["loss"] - float(G[tag + "/getloss"])) < 1e-9 * float(G[tag + "/getloss"])
    import torch
    o = OT.forward(OT.to_torch(W, torch.float64), torch.tensor(X, dtype=torch.float64), variant)
    t16 = torch.cat([o["base"], o["zygosity"], o["varType"], o["indelLength"]], 1).numpy()
    assert np.abs(t16 - G[tag + "/predict"]).max() < 1e-12


def _check_variables(W, prefix):
    """against the fixture's `compact` form: small variables in full, large ones by strided sample + sum + sum of squares"""
    for k, v in W.items():
        v = np.asarray(v, np.float64).reshape(-1)
        if prefix + k in G.files:
            ref = G[prefix + k]
            assert np.abs(v - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), k
        else:
            ref = G[prefix + k + "#sample"]
            got = v[::v.size // len(ref)][:len(ref)]
            assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), k
            sums = G[prefix + k + "#sums"]
            assert abs(v.sum() - sums[0]) <= 1e-9 * max(1.0, abs(sums[0])) + 1e-9 and abs((v * v).sum() - sums[1]) <= 1e-9 * sums[1], k


def _adam_steps(W, variant, steps, **fw):
    m = {k: np.zeros_like(v) for k, v in W.items()}
    v2 = {k: np.zeros_like(v) for k, v in W.items()}
    losses = []
    for t in range(1, steps + 1):
        l, g = OT.loss_and_grads(W, X, Y, variant, l2_lambda=1e-3, **fw)
        losses.append(l)
        for k in W:
            W[k], m[k], v2[k] = OT.tf_adam_step(W[k], g[k], m[k], v2[k], t, 1e-3)
    return losses, W


@pytest.mark.parametrize("tag,variant", VARIANTS)
def test_two_training_steps_without_dropout(tag, variant):
    losses, W = _adam_steps(_weights(variant), variant, 2)
    want = G[tag + "/train_losses"]
    assert np.allclose(losses, want, rtol=1e-10, atol=0)
    _check_variables(W, tag + "/after2/")


@pytest.mark.parametrize("tag,variant", VARIANTS)
def test_training_step_with_the_recorded_dropout_noise(tag, variant):
    mask = G[tag + "/dropout_mask4"].astype(np.float64)    # floor(keep_prob + U[0,1)) as the graph drew it (selu.py:55-57)
    losses, W = _adam_steps(_weights(variant), variant, 1, drop4_rate=0.5, drop4_mask=mask)
    assert abs(losses[0] - float(G[tag + "/dropout_loss"])) <= 1e-10 * abs(losses[0])
    _check_variables(W, tag + "/after_dropout/")
