"""Generates tests/golden/reference_graph.npz by executing the reference's OWN model files -- clairvoyante/clairvoyante_v3.py,
clairvoyante_v3_slim.py and selu.py, read from /root/reference at generation time, unmodified -- on the TensorFlow-1.x
stand-in of tests/golden/tf1_stand_in.py (TensorFlow itself is not installable here).

What the fixture therefore pins: everything the reference's graph code decides (layers, sizes, paddings, pool windows, head
wiring, SELU / dropout_selu formulas, epsilon, loss composition, regularised variables, the feeds of predict / getLoss /
train).  What it does not pin: the numerical kernels of the TensorFlow ops, which the stand-in restates on torch float64 from
their documented semantics (header of tf1_stand_in.py).  So for the network arithmetic this is "reference wiring on
stand-in ops", not a TensorFlow run; DESIGN.md section 9 says so.

Per variant: reference-initialiser weights (seed 0) are assigned to the graph's variables by their TensorFlow names, then the
reference's public methods are called on 256 synthetic sites:
  predict(X)                     -> the four head outputs
  getLoss(X, Y)                  -> loss with phase False, lambda 0
  train(X, Y) x 2, dropout 0     -> the two losses and every variable after the second Adam step
  train(X, Y), dropoutRateFC4 .5 -> loss, the U[0,1) noise the graph drew (recorded), every variable after the step
Variables with more than 20,000 elements (fc4/kernel, fc5/kernel) are stored as a strided sample of 8,192 elements plus
their sum and sum of squares (`compact` below) to keep the fixture small; the others in full.

    python tests/golden/make_golden_reference_graph.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CV_DIR = "/root/reference/clairvoyante"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import tf1_stand_in as TF   # noqa: E402
from clairvoyante_b200 import initializers, synth   # noqa: E402


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(CV_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


SAMPLE = 8192


def compact(a):
    """full array when small, else (strided sample, [sum, sum of squares])"""
    a = np.asarray(a, np.float64).reshape(-1)
    if a.size <= 20000:
        return {"": a}
    step = a.size // SAMPLE
    return {"#sample": a[::step][:SAMPLE].copy(), "#sums": np.array([a.sum(), (a * a).sum()])}


def main():
    installed = TF.install()
    load("param")
    load("selu")
    fixture = {}
    x = synth.make_sites(256, seed=21)
    y = synth.make_labels(256, seed=21).astype(np.float64)
    fixture["x"], fixture["y"] = x.astype(np.int16), y
    assert np.array_equal(fixture["x"], x)
    for variant, modname in (("v3", "clairvoyante_v3"), ("slim", "clairvoyante_v3_slim")):
        cv = load(modname)
        W = initializers.init_weights("v3" if variant == "v3" else "v3_slim", seed=0)

        def fresh():
            m = cv.Clairvoyante()
            m.init()
            byname = m.g.by_name()
            assert sorted(byname) == sorted(W), (sorted(byname), sorted(W))
            for k, v in W.items():
                byname[k].assign(v)
            return m

        m = fresh()
        assert len(m.g.variables) == 18
        fixture[variant + "/n_params"] = np.array(sum(int(np.prod(v.shape)) for v in m.g.variables))
        base, z, t, l = m.predict(x)
        fixture[variant + "/predict"] = np.concatenate([base, z, t, l], axis=1)
        fixture[variant + "/getloss"] = np.array(m.getLoss(x, y))
        m.dropoutRateFC4Val = 0.0
        m.setLearningRate(1e-3)
        m.setL2RegularizationLambda(1e-3)
        l1, _ = m.train(x, y)
        l2, _ = m.train(x, y)
        fixture[variant + "/train_losses"] = np.array([l1, l2])
        for k, v in m.g.by_name().items():
            for sfx, a in compact(v.value.detach().numpy()).items():
                fixture[variant + "/after2/" + k + sfx] = a
        # one step with the reference's default dropout (0.5 on fc4, 0.0 on fc5), noise recorded
        m = fresh()
        draws = []
        rs = np.random.RandomState(99)

        def hook(shape):
            u = rs.random_sample(shape)
            draws.append(u)
            return u
        TF.set_uniform_hook(hook)
        ld, _ = m.train(x, y)
        TF.set_uniform_hook(None)
        assert len(draws) == 2 and draws[0].shape[0] == 256, [d.shape for d in draws]
        fixture[variant + "/dropout_loss"] = np.array(ld)
        fixture[variant + "/dropout_mask4"] = np.floor(0.5 + draws[0]).astype(np.uint8)   # selu.py:55-57, keep_prob 0.5
        for k, v in m.g.by_name().items():
            for sfx, a in compact(v.value.detach().numpy()).items():
                fixture[variant + "/after_dropout/" + k + sfx] = a
        print("%-5s params %d  getLoss %.6f  train %.6f -> %.6f  with dropout %.6f" % (
            variant, int(fixture[variant + "/n_params"]), float(fixture[variant + "/getloss"]), l1, l2, ld))
    np.savez_compressed(os.path.join(HERE, "reference_graph.npz"), **fixture)
    for k in installed + ["param", "selu", "clairvoyante_v3", "clairvoyante_v3_slim"]:
        sys.modules.pop(k, None)
    print("written", os.path.join(HERE, "reference_graph.npz"))


if __name__ == "__main__":
    main()
