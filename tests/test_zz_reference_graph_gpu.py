"""The CUDA path against tests/golden/reference_graph.npz -- what the reference's own graph code (clairvoyante_v3.py,
clairvoyante_v3_slim.py, selu.py) computed on the TensorFlow stand-in of tests/golden/tf1_stand_in.py: head outputs of
predict, getLoss, and the loss fetched by the first train step (dropout 0, lambda 1e-3).  Same bars as
tests/test_forward_gpu.py / test_train_gpu.py use against the oracle."""
import os

import numpy as np
import pytest

from clairvoyante_b200 import initializers as I

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_graph.npz"))
X = G["x"].astype(np.float32)
Y = G["y"]


def _model(variant, **kw):
    if variant == "v3":
        from clairvoyante_b200 import clairvoyante_v3 as cv
    else:
        from clairvoyante_b200 import clairvoyante_v3_slim as cv
    m = cv.Clairvoyante(**kw)
    m.setWeights(I.init_weights(variant, seed=0))
    return m


@pytest.mark.parametrize("tag,variant", [("v3", "v3"), ("slim", "v3_slim")])
def test_predict_getloss_and_first_train_loss(tag, variant):
    m = _model(variant, dropoutRateFC4=0.0, l2RegularizationLambda=1e-3, initialLearningRate=1e-3)
    base, z, t, l = m.predict(X)
    out = np.concatenate([base, z, t, l], axis=1)
    assert np.abs(out - G[tag + "/predict"]).max() <= 2e-4
    want = float(G[tag + "/getloss"])
    assert abs(float(m.getLoss(X, Y)) - want) <= 2e-5 * abs(want) + 1e-3
    m.init(seed=1)                                     # (Adam slots and step zeroed, as in tests/test_train_gpu.py)
    m.setWeights(I.init_weights(variant, seed=0))
    loss, _ = m.train(X, Y)
    want = float(G[tag + "/train_losses"][0])
    assert abs(float(loss) - want) <= 3e-5 * abs(want) + 1e-3
    m.close()
