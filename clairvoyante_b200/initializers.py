"""Reference weight initialisers, NumPy, seeded.

conv*/fc4/fc5 kernels: tf.contrib.layers.variance_scaling_initializer(factor=2.0,
mode='FAN_IN', uniform=False) (clairvoyante_v3.py:57,72,87,106,116) = truncated
normal, stddev sqrt(1.3*factor/fan_in), resampled outside +-2 sigma.
Head kernels: tf.layers.dense default = glorot_uniform (clairvoyante_v3.py:125-135).
All biases zero.  The reference is unseeded; a seed here only makes runs repeatable.
"""
import math
import numpy as np

H_IN, W_IN, C_IN = 33, 4, 4
HEAD_NAMES = ("YBaseChangeSigmoid", "YZygosityFC", "YVarTypeFC", "YIndelLengthFC")
HEAD_SIZES = (4, 2, 4, 6)
VARIANTS = {
    "v3": dict(convs=[(1, 16, 5), (2, 32, 4), (3, 48, 3)], fc4=336, fc5=168),
    "v3_slim": dict(convs=[(1, 8, 1), (3, 16, 1), (5, 32, 1)], fc4=36, fc5=18),
}


def variable_shapes(variant):
    """[(tf_name, shape)] in graph-creation order (jupyter_nb/visualization.ipynb:100-121)."""
    spec = VARIANTS[variant]
    out = []
    cin, h = C_IN, H_IN
    for i, (kh, cout, pool) in enumerate(spec["convs"], 1):
        out.append(("conv%d/kernel" % i, (kh, 4, cin, cout)))
        out.append(("conv%d/bias" % i, (cout,)))
        cin = cout
        h -= pool - 1
    out.append(("fc4/kernel", (h * W_IN * cin, spec["fc4"])))
    out.append(("fc4/bias", (spec["fc4"],)))
    out.append(("fc5/kernel", (spec["fc4"], spec["fc5"])))
    out.append(("fc5/bias", (spec["fc5"],)))
    for name, n, src in zip(HEAD_NAMES, HEAD_SIZES, (spec["fc4"],) + (spec["fc5"],) * 3):
        out.append((name + "/kernel", (src, n)))
        out.append((name + "/bias", (n,)))
    return out


def _truncated_normal(rng, shape, stddev):
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * stddev).astype(np.float32)


def init_weights(variant="v3", seed=0, head_gain=1.0):
    """dict tf_name -> float32 ndarray.  head_gain scales the 4 head kernels only
    (1.0 = reference initialiser); the synthetic benchmark keeps 1.0."""
    rng = np.random.default_rng(seed)
    W = {}
    for name, shape in variable_shapes(variant):
        if name.endswith("/bias"):
            W[name] = np.zeros(shape, np.float32)
        elif name.split("/")[0] in HEAD_NAMES:
            limit = math.sqrt(6.0 / (shape[0] + shape[1])) * head_gain
            W[name] = rng.uniform(-limit, limit, shape).astype(np.float32)
        else:
            fan_in = int(np.prod(shape[:-1]))
            W[name] = _truncated_normal(rng, shape, math.sqrt(1.3 * 2.0 / fan_in))
    return W
