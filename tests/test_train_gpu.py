"""GPU parity of the loss / backward / Adam path (clairvoyante_v3.py:140-152,174,183-227) against the
oracle: NumPy loss, torch-fp64 autograd gradients, restated TF-1.x Adam."""
import numpy as np
import pytest

from clairvoyante_b200 import dropout_rng, initializers as I, synth
from oracle import cv_oracle as O, cv_oracle_torch as OT

pytestmark = pytest.mark.gpu


VARIANTS = ["v3", "v3_slim"]
N4 = {"v3": 336, "v3_slim": 36}


def _model(W, variant="v3", **kw):
    if variant == "v3":
        from clairvoyante_b200 import clairvoyante_v3 as cv
    else:
        from clairvoyante_b200 import clairvoyante_v3_slim as cv
    m = cv.Clairvoyante(**kw)
    m.setWeights(W)
    return m


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", [1, 7, 1000, 5121])
def test_get_loss_matches_oracle(variant, n):
    W = I.init_weights(variant, 3)
    x, y = synth.make_sites(n, 4), synth.make_labels(n, 4)
    m = _model(W, variant)
    got = float(m.getLoss(x, y))
    idx = slice(0, min(n, 1500))
    ref = O.loss(W, x, y, variant, 0.0)["loss"] if n <= 1500 else None
    if ref is not None:
        assert abs(got - ref) <= 2e-5 * abs(ref) + 1e-3
    else:  # additivity over micro-chunks (5120 sites): loss(all) == loss(a) + loss(b)
        a = float(m.getLoss(x[:3000], y[:3000])); b = float(m.getLoss(x[3000:], y[3000:]))
        assert abs(got - (a + b)) <= 1e-5 * abs(got)
    assert m.getLoss(x[:0], y[:0]) == 0.0
    m.getLossNoRT(x[idx], y[idx])
    assert m.getLossLossRTVal is not None
    m.close()


# tolerance on max|g - g_fp64| / max|g_fp64| per variable.  fp32 SIMT and split-bf16 tensor kernels (~2^-16 per operand)
# share one bar; plain bf16 operands (2^-9 per operand, BASELINE config "bf16 compute") get the looser one.
GRAD_TOL = {"fp32": 2e-3, "bf16x3": 2e-3, "bf16": 5e-2}


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("rate", [0.0, 0.5])
@pytest.mark.parametrize("mode", ["bf16x3", "fp32", "bf16"])
def test_gradients_match_autograd(variant, rate, mode):
    W = I.init_weights(variant, 5)
    n = 300
    x, y = synth.make_sites(n, 6), synth.make_labels(n, 6)
    m = _model(W, variant, dropoutRateFC4=rate)
    m.setTrainMode(mode)
    seed = 0x1234ABCD
    loss, summary = m._train_step(x, y, apply_update=0, seed=seed)
    g = m.getGradients()
    mask = dropout_rng.keep_mask(seed, n, rate, width=N4[variant]) if rate > 0 else None
    ref_loss, ref_g = OT.loss_and_grads(W, x, y, variant, 0.0, drop4_rate=rate, drop4_mask=mask)
    for name in sorted(ref_g):
        assert _relerr(g[name], ref_g[name]) < GRAD_TOL[mode], name
    m.close()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("n", [300, 5121 + 77])
def test_fc5_dropout_gradients_match_autograd(variant, mode, n):
    """dropoutRateFC5 != 0 (clairvoyante_v3.py:121; param.py:22 defaults to 0 but the kwarg accepts any rate): SELU dropout
    on FC5's output, which feeds the zygosity / varType / indelLength heads, together with FC4's; masks replayed on the host
    (two micro-chunks at n = 5198: the counters run over the whole batch); getLoss stays dropout-free (phase = False)"""
    N5 = {"v3": 168, "v3_slim": 18}
    W = I.init_weights(variant, 15)
    x, y = synth.make_sites(n, 16), synth.make_labels(n, 16)
    m = _model(W, variant, dropoutRateFC4=0.5, dropoutRateFC5=0.25)
    m.setTrainMode(mode)
    seed = 0x77AA55
    plain = float(m.getLoss(x, y))
    loss, _ = m._train_step(x, y, apply_update=0, seed=seed)
    g = m.getGradients()
    m4 = dropout_rng.keep_mask(seed, n, 0.5, width=N4[variant])
    m5 = dropout_rng.keep_mask(seed ^ 0x5D5D5D5D5D5D5D5D, n, 0.25, width=N5[variant])
    ref_loss, ref_g = OT.loss_and_grads(W, x, y, variant, 0.0, drop4_rate=0.5, drop4_mask=m4, drop5_rate=0.25, drop5_mask=m5)
    assert abs(float(m.getLoss(x, y)) - plain) <= 1e-6 * abs(plain)
    for name in sorted(ref_g):
        assert _relerr(g[name], ref_g[name]) < GRAD_TOL[mode], name
    m.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_count_feed_trains_identically(variant):
    """cvb_train_step_host_x / cvb_loss_host_x: a batch that arrives as raw uint8 / int16 counts (a CountBatch, or the explicit
    arrays) is widened per micro-chunk on the device to exactly the float32 tensors of the ordinary call; losses and gradients
    agree to the rounding of their fp32 atomic sums (whose order varies from run to run on either feed).  Two micro-chunks,
    ragged sizes, both arithmetic paths."""
    from clairvoyante_b200 import utils_v2 as U
    W = I.init_weights(variant, 11)
    n = 5120 + 333
    x, y = synth.make_labeled_sites(n, 12)
    cnt = U.pack_counts(x)
    assert cnt is not None and cnt.dtype == np.uint8
    m = _model(W, variant, dropoutRateFC4=0.5)
    ref_loss = float(m.getLoss(x, y))
    for feed in (U.with_counts(x), cnt, cnt.astype(np.int16), x.astype(np.float16)):
        assert abs(float(m.getLoss(feed, y)) - ref_loss) <= 1e-6 * abs(ref_loss)   # (per-site terms meet in fp32 atomics)
    m.setTrainMode("fp32")
    l0, _ = m._train_step(x[:3000], y[:3000], apply_update=0, seed=5)
    g0 = m.getGradients()
    l1, _ = m._train_step(U.with_counts(x[:3000]), y[:3000], apply_update=0, seed=5)
    g1 = m.getGradients()
    assert abs(float(l0) - float(l1)) <= 1e-6 * abs(float(l0))
    for k in g0:
        assert _relerr(g1[k], g0[k]) < 1e-5, k          # (fp32 atomics in the weight-gradient sums: order, not values, may differ)
    m.setTrainMode("bf16x3")
    l2, _ = m._train_step(x, y, apply_update=0, seed=6)
    g2 = m.getGradients()
    l3, _ = m._train_step(cnt, y, apply_update=0, seed=6)
    g3 = m.getGradients()
    assert abs(float(l2) - float(l3)) <= 1e-6 * abs(float(l2))
    for k in g2:
        assert _relerr(g3[k], g2[k]) < 1e-4, k
    m.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_native_init_draws_the_reference_initialisers(variant):
    """cvb_init_weights (init(), clairvoyante_v3.py:177-178): truncated normal with stddev sqrt(2.6 / fan_in) cut at two sigma
    for conv / fc4 / fc5 kernels, glorot-uniform heads, zero biases, zero Adam slots and step; seeded and repeatable"""
    m = _model(I.init_weights(variant, 0), variant)
    m.train(synth.make_sites(64, 1), synth.make_labels(64, 1))          # leaves non-zero slots and step 1 behind
    m.init(seed=123)
    W = m.getWeights()
    assert m.getStep() == 0
    for name, shape in I.variable_shapes(variant):
        w = W[name]
        assert w.shape == tuple(shape)
        assert not m._get(name, 1, shape).any() and not m._get(name, 2, shape).any()
        if name.endswith("bias"):
            assert not w.any()
        elif name.startswith("Y"):
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            assert np.abs(w).max() <= lim and (w.size < 100 or np.abs(w).max() > 0.8 * lim)
        else:
            sd = np.sqrt(2.6 / np.prod(shape[:-1]))
            assert np.abs(w).max() <= 2.0 * sd * (1 + 1e-6)
            if w.size >= 2000:                                           # std of a normal truncated at 2 sigma = 0.8796 sigma
                assert abs(w.std() / sd - 0.8796) < 0.05 and abs(w.mean()) < 0.1 * sd
    m2 = _model(I.init_weights(variant, 0), variant)
    m2.init(seed=123)
    W2 = m2.getWeights()
    assert all(np.array_equal(W[k], W2[k]) for k in W)
    m2.init(seed=124)
    assert not np.array_equal(m2.getWeights()["fc4/kernel"], W["fc4/kernel"])
    m.close(); m2.close()


@pytest.mark.parametrize("variant,n", [("v3", 1), ("v3", 37), ("v3", 129), ("v3", 5121), ("v3_slim", 37), ("v3_slim", 5121)])
def test_tensor_and_simt_training_paths_agree(variant, n):
    """the tcgen05 contractions (split bf16) against the fp32 SIMT kernels on ragged batch sizes: K = sites of the
    weight gradient not a multiple of the 32-wide K block, M tiles with masked rows, two micro-chunks"""
    W = I.init_weights(variant, 13)
    x, y = synth.make_sites(n, 14), synth.make_labels(n, 14)
    m = _model(W, variant, dropoutRateFC4=0.5)
    out = {}
    for mode in ("fp32", "bf16x3"):
        m.setTrainMode(mode)
        loss, _ = m._train_step(x, y, apply_update=0, seed=99)
        out[mode] = (float(loss), m.getGradients(), float(m.getLoss(x, y)))
    assert abs(out["fp32"][2] - out["bf16x3"][2]) <= 3e-5 * abs(out["fp32"][2])
    # split bf16 keeps 16 mantissa bits per operand; weight gradients are sums over all sites with heavy cancellation
    # (fc5/kernel at n = 5121: 3e-4 of the largest entry).  The bar against fp64 autograd is test_gradients_match_autograd's.
    for k in out["fp32"][1]:
        assert _relerr(out["bf16x3"][1][k], out["fp32"][1][k]) < 1e-3, k
    m.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_train_step_loss_and_tf_adam_update(variant):
    """loss fetched is that of the pre-update weights incl. lambda*sum(0.5 w^2); update is TF-1.x Adam"""
    W = I.init_weights(variant, 7)
    n = 256
    x, y = synth.make_sites(n, 8), synth.make_labels(n, 8)
    lam, lr = 1e-3, 1e-3
    m = _model(W, variant, dropoutRateFC4=0.0, l2RegularizationLambda=lam, initialLearningRate=lr)
    m.init(seed=1); m.setWeights(W)
    loss, summary = m.train(x, y)
    ref = O.loss(W, x, y, variant, lam)
    assert abs(float(loss) - ref["loss"]) <= 3e-5 * abs(ref["loss"])
    for k in ("loss1", "loss2", "loss3", "loss4", "lossL2"):
        assert abs(summary[k] - ref[k]) <= 1e-4 * abs(ref[k]) + 1e-4, k
    _, g = OT.loss_and_grads(W, x, y, variant, lam)      # lambda inside the loss == g + lam*w
    W1 = m.getWeights()
    for name in W:
        exp, _, _ = OT.tf_adam_step(W[name].astype(np.float64), g[name], 0.0, 0.0, 1, lr)
        # step-1 Adam moves every weight by ~lr*sign(g); compare where the gradient is not vanishing
        big = np.abs(g[name]) > 1e-3 * np.abs(g[name]).max()
        assert np.abs(W1[name] - exp)[big].max() < 0.05 * lr, name
    # second step: t = 2 with the stored slots (oracle gradients taken at the GPU's own W1)
    m1 = {k: 0.1 * g[k] for k in g}; v1 = {k: 0.001 * g[k] ** 2 for k in g}
    _, g2 = OT.loss_and_grads(W1, x, y, variant, lam)
    m.train(x, y)
    W2 = m.getWeights()
    for name in W:
        exp, _, _ = OT.tf_adam_step(W1[name].astype(np.float64), g2[name], m1[name], v1[name], 2, lr)
        big = (np.abs(g[name]) > 1e-3 * np.abs(g[name]).max()) & (np.abs(g2[name]) > 1e-3 * np.abs(g2[name]).max())
        assert np.abs(W2[name] - exp)[big].max() < 0.05 * lr, name
    m.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_training_reduces_loss_and_checkpoint_roundtrip(variant, tmp_path):
    W = I.init_weights(variant, 9)
    x, y = synth.make_sites(2000, 10), synth.make_labels(2000, 10)
    m = _model(W, variant, initialLearningRate=1e-4)      # Adam moves every weight by ~lr per step: 1e-3 overshoots random weights
    m.init(seed=2)
    l0 = float(m.getLoss(x, y))
    for _ in range(10):
        m.train(x, y)
    l1 = float(m.getLoss(x, y))
    assert l1 < l0
    fn = str(tmp_path / "ck" / "model-000001")
    m.saveParameters(fn)
    m2 = _model(W, variant)
    m2.restoreParameters(fn)
    assert abs(float(m2.getLoss(x, y)) - l1) <= 1e-6 * abs(l1)
    a, b = m.train(x, y)[0], m2.train(x, y)[0]        # Adam slots + step restored -> identical next step (dropout differs: rate!=0)
    m.close(); m2.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_checkpoint_is_a_tensorflow_bundle(variant, tmp_path):
    """saveParameters writes the tf.train.Saver layout (clairvoyante_v3.py:243-251): restoring from the bundle alone
    brings back weights, Adam slots and the step; the .cvb.npz alone does the same"""
    import os
    from clairvoyante_b200 import tf_bundle
    W = I.init_weights(variant, 4)
    x, y = synth.make_sites(500, 3), synth.make_labels(500, 3)
    m = _model(W, variant, initialLearningRate=1e-4, dropoutRateFC4=0.0)
    m.init(seed=5); m.setWeights(W)
    for _ in range(3):
        m.train(x, y)
    fn = str(tmp_path / "out" / "m-000003")
    m.saveParameters(fn)
    ent = tf_bundle.read_bundle(fn)
    assert len(ent) == 18 * 3 + 2 and abs(float(ent["beta1_power"]) - 0.9 ** 4) < 1e-7
    W3 = m.getWeights()
    assert all(np.array_equal(ent[k], W3[k]) for k in W3)
    la = m.train(x, y)[0]
    for drop in (".cvb.npz", ".index"):                       # bundle only, then npz only
        os.rename(fn + drop, fn + drop + ".bak")
        m2 = _model(W, variant, initialLearningRate=1e-4, dropoutRateFC4=0.0)
        m2.restoreParameters(fn)
        assert m2.getStep() == 3
        W2 = m2.getWeights()
        assert all(np.array_equal(W2[k], W3[k]) for k in W3)
        lb = m2.train(x, y)[0]
        assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
        assert _relerr(m2.getWeights()["fc4/kernel"], m.getWeights()["fc4/kernel"]) < 1e-6   # same Adam slots -> same update
        m2.close()
        os.rename(fn + drop + ".bak", fn + drop)
    other = "v3_slim" if variant == "v3" else "v3"
    m3 = _model(I.init_weights(other, 0), other)
    with pytest.raises(ValueError):
        m3.restoreParameters(fn)                              # wrong variant: shapes differ
    m3.close(); m.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_data_parallel_halves_equal_full_batch(variant):
    """gradient of a batch == sum of the gradients of its shards (SUM loss): what the DP all-reduce relies on"""
    W = I.init_weights(variant, 11)
    n = 400
    x, y = synth.make_sites(n, 12), synth.make_labels(n, 12)
    m = _model(W, variant, dropoutRateFC4=0.0)
    m._train_step(x, y, apply_update=0, seed=1); full = m.getGradients()
    m._train_step(x[:150], y[:150], apply_update=0, seed=1); a = m.getGradients()
    m._train_step(x[150:], y[150:], apply_update=0, seed=1); b = m.getGradients()
    for k in full:
        assert _relerr(a[k] + b[k], full[k]) < 2e-5, k
    m.close()


def test_set_learning_rate_and_lambda_semantics():
    W = I.init_weights("v3", 0)
    m = _model(W)
    assert m.setLearningRate(0.01) == 0.01 and abs(m.setLearningRate() - 0.001) < 1e-12       # clairvoyante_v3.py:229-234
    assert m.setL2RegularizationLambda(0.5) == 0.5 and abs(m.setL2RegularizationLambda() - 0.05) < 1e-12
    m.close()
