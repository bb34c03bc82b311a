"""CPU oracle for the Clairvoyante v3 / v3_slim hot path.  TEST INFRASTRUCTURE ONLY.

This is a NumPy restatement of what the reference's TensorFlow graph computes
(forward, loss, and -- in cv_oracle_torch.py -- gradients and the TF-1.x Adam
update).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product path
(clairvoyante_b200/) never does and fails loudly without its CUDA library.

PARITY: forward pinned (e), training arithmetic UNPINNED.  The reference's arithmetic lives in TensorFlow 1.12
(requirements.txt:1), which is not vendored under /root/reference and not
installable here (no wheel, no network, Python-2 sources).  The reference ships
no tests, no golden vectors and no weights.  What pins this
oracle is (a) the activation/parameter shapes recorded in
jupyter_nb/visualization.ipynb:112-121,475-739 (checked in
tests/test_oracle.py), (b) a second, independent restatement on torch-CPU
(cv_oracle_torch.py: F.conv2d / F.max_pool2d in NCHW) that must agree with
this one, (c) hand-computed padding cases for TF `SAME` semantics, and (d) the
reference's OWN graph code -- clairvoyante_v3.py, clairvoyante_v3_slim.py,
selu.py, unmodified -- executed at fixture-generation time on a stand-in for the
TensorFlow-1.x API (tests/golden/tf1_stand_in.py, op kernels on torch float64):
tests/golden/reference_graph.npz holds its predict / getLoss / train results and
tests/test_reference_graph_cpu.py requires this oracle to reproduce them to
rounding.  (d) ties the wiring, formulas and feeds to the reference's source.
(e) TensorFlow's FORWARD op semantics -- Conv2D SAME with even kernels, VALID
max-pool, the NHWC flatten order, MatMul + BiasAdd, sigmoid / softmax -- are
pinned against an independent executor of TensorFlow graphs: the forward
GraphDef of both variants (tests/golden/tf_graphdef.py, TF's proto schemas)
run by OpenCV's dnn TensorFlow importer; tests/golden/opencv_tf_forward.npz
holds its outputs and tests/test_oracle_opencv_cpu.py requires this oracle to
match them (6e-7 on the head outputs).  Training: gradients are torch autograd
of that forward function (cv_oracle_torch.py); TF-1.x Adam is restated from its
documented update and cross-checked against torch.optim.Adam with epsilon moved
to TensorFlow's position (tests/test_oracle.py) -- that placement of epsilon is
the one thing still "unpinned".

Reference anchors (all paths relative to /root/reference):
  graph v3      clairvoyante/clairvoyante_v3.py:54-138
  graph slim    clairvoyante/clairvoyante_v3_slim.py:54-118
  SELU          clairvoyante/selu.py:21-25
  SELU dropout  clairvoyante/selu.py:34-69
  loss          clairvoyante/clairvoyante_v3.py:140-152
  feed prep     clairvoyante/utils_v2.py:45-46

TF semantics restated explicitly: NHWC activations, HWIO kernels,
cross-correlation (no kernel flip), stride 1, `SAME` zero padding with
pad_total = k-1, pad_before = pad_total // 2 (so the width-4 kernel on W=4
pads 1 left / 2 right and the height-2 kernel pads 0 top / 1 bottom),
`VALID` max-pool over H only, tf.layers.dense = x @ W + b, softmax and
log_softmax computed with the row max subtracted.
"""
import numpy as np

SELU_ALPHA = 1.6732632423543772848170429916717   # selu.py:23
SELU_SCALE = 1.0507009873554804934193349852946   # selu.py:24
DROPOUT_ALPHA = -1.7580993408473766               # selu.py:34

# (kernel_h, out_channels, pool_h) per conv layer; kernel_w is always 4, pool_w 1.
VARIANTS = {
    # clairvoyante_v3.py:9-12
    "v3": dict(convs=[(1, 16, 5), (2, 32, 4), (3, 48, 3)], fc4=336, fc5=168),
    # clairvoyante_v3_slim.py:9-11 (no pooling layers: :63,:72)
    "v3_slim": dict(convs=[(1, 8, 1), (3, 16, 1), (5, 32, 1)], fc4=36, fc5=18),
}
H_IN, W_IN, C_IN = 33, 4, 4            # param.py:6-7 -> (2*16+1, 4, 4)
HEAD_SIZES = (4, 2, 4, 6)              # clairvoyante_v3.py:8
HEAD_NAMES = ("YBaseChangeSigmoid", "YZygosityFC", "YVarTypeFC", "YIndelLengthFC")


def variable_shapes(variant):
    """TF variable names -> shapes, in graph creation order (visualization.ipynb:100-121)."""
    spec = VARIANTS[variant]
    shapes = []
    cin, h = C_IN, H_IN
    for i, (kh, cout, pool) in enumerate(spec["convs"], 1):
        shapes.append(("conv%d/kernel" % i, (kh, 4, cin, cout)))
        shapes.append(("conv%d/bias" % i, (cout,)))
        cin = cout
        h = h - (pool - 1)
    flat = h * W_IN * cin
    shapes.append(("fc4/kernel", (flat, spec["fc4"])))
    shapes.append(("fc4/bias", (spec["fc4"],)))
    shapes.append(("fc5/kernel", (spec["fc4"], spec["fc5"])))
    shapes.append(("fc5/bias", (spec["fc5"],)))
    for name, n, src in zip(HEAD_NAMES, HEAD_SIZES, (spec["fc4"],) + (spec["fc5"],) * 3):
        shapes.append((name + "/kernel", (src, n)))
        shapes.append((name + "/bias", (n,)))
    return shapes


def selu(x):
    """selu.py:21-25: scale * where(x >= 0, x, alpha * elu(x)); elu(x) = exp(x) - 1 for x < 0."""
    neg = np.minimum(x, 0)
    return SELU_SCALE * np.where(x >= 0.0, x, SELU_ALPHA * (np.exp(neg) - 1.0))


def _same_pad(k):
    total = k - 1
    before = total // 2
    return before, total - before


def conv2d_same(x, kernel, bias):
    """tf.layers.conv2d(padding='same', strides=1) on NHWC input, HWIO kernel
    (clairvoyante_v3.py:54-60,69-75,84-90)."""
    n, h, w, cin = x.shape
    kh, kw, kcin, cout = kernel.shape
    assert kcin == cin
    pt, pb = _same_pad(kh)
    pl, pr = _same_pad(kw)
    xp = np.zeros((n, h + pt + pb, w + pl + pr, cin), dtype=x.dtype)
    xp[:, pt:pt + h, pl:pl + w, :] = x
    out = np.zeros((n, h, w, cout), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            out += xp[:, i:i + h, j:j + w, :] @ kernel[i, j]
    return out + bias


def maxpool_h(x, p):
    """tf.layers.max_pooling2d(pool_size=(p,1), strides=1) ('valid'); clairvoyante_v3.py:63-66."""
    if p == 1:
        return x
    h = x.shape[1]
    out = x[:, 0:h - p + 1]
    for j in range(1, p):
        out = np.maximum(out, x[:, j:j + h - p + 1])
    return out


def softmax(z):
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=1, keepdims=True)


def log_softmax(z):
    m = z.max(axis=1, keepdims=True)
    s = z - m
    return s - np.log(np.exp(s).sum(axis=1, keepdims=True))


def dropout_selu(x, rate, mask=None):
    """selu.py:38-64 with fixedPointMean 0, fixedPointVar 1.  `mask` is the
    binary keep tensor floor(keep_prob + U[0,1)); rate == 0 returns x (selu.py:52-53)."""
    keep = 1.0 - rate
    if keep == 1.0:
        return x
    a = np.sqrt(1.0 / (keep * ((1.0 - keep) * DROPOUT_ALPHA ** 2 + 1.0)))
    b = -a * ((1.0 - keep) * DROPOUT_ALPHA)
    return a * (x * mask + DROPOUT_ALPHA * (1.0 - mask)) + b


def forward(weights, x, variant="v3", dtype=np.float64, return_all=False,
            drop4_rate=0.0, drop4_mask=None):
    """Forward pass in `dtype`.  Returns dict with the 4 head outputs
    (clairvoyante_v3.py:124-138), the 16 pre-activation logits per site
    (base: pre-sigmoid; others: SELU(FC)+1e-10) and, if asked, every layer.

    weights: dict TF-name -> ndarray (shapes from variable_shapes()).
    x: (N,33,4,4) already channel-subtracted (utils_v2.py:46).
    """
    spec = VARIANTS[variant]
    W = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    a = np.asarray(x, dtype=dtype).reshape(-1, H_IN, W_IN, C_IN)
    layers = {}
    for i, (kh, cout, pool) in enumerate(spec["convs"], 1):
        a = selu(conv2d_same(a, W["conv%d/kernel" % i], W["conv%d/bias" % i]))
        layers["conv%d" % i] = a
        a = maxpool_h(a, pool)
        layers["pool%d" % i] = a
    flat = a.reshape(a.shape[0], int(np.prod(a.shape[1:])))               # (h, w, c) order, c fastest; :99-102
    fc4 = selu(flat @ W["fc4/kernel"] + W["fc4/bias"])
    d4 = dropout_selu(fc4, drop4_rate, drop4_mask)
    fc5 = selu(d4 @ W["fc5/kernel"] + W["fc5/bias"])
    d5 = fc5                                       # dropoutRateFC5 = 0.0 (param.py:22)
    eps = dtype(1e-10)
    base_logit = d4 @ W["YBaseChangeSigmoid/kernel"] + W["YBaseChangeSigmoid/bias"]
    base = 1.0 / (1.0 + np.exp(-base_logit))
    zl = selu(d5 @ W["YZygosityFC/kernel"] + W["YZygosityFC/bias"]) + eps
    tl = selu(d5 @ W["YVarTypeFC/kernel"] + W["YVarTypeFC/bias"]) + eps
    ll = selu(d5 @ W["YIndelLengthFC/kernel"] + W["YIndelLengthFC/bias"]) + eps
    out = dict(base=base, zygosity=softmax(zl), varType=softmax(tl), indelLength=softmax(ll),
               logits=np.concatenate([base_logit, zl, tl, ll], axis=1))
    if return_all:
        layers.update(fc4=fc4, fc5=fc5, flat=flat)
        out["layers"] = layers
    return out


def loss(weights, x, y, variant="v3", l2_lambda=0.0, dtype=np.float64, **fw):
    """clairvoyante_v3.py:140-151.  SUM over the batch (not mean); the L2 term
    is lambda * sum_{non-bias vars} 0.5*||v||^2 (tf.nn.l2_loss)."""
    o = forward(weights, x, variant, dtype, **fw)
    y = np.asarray(y, dtype=dtype)
    lg = o["logits"]
    l1 = ((o["base"] - y[:, 0:4]) ** 2).sum()
    l2 = (-y[:, 4:6] * log_softmax(lg[:, 4:6])).sum()
    l3 = (-y[:, 6:10] * log_softmax(lg[:, 6:10])).sum()
    l4 = (-y[:, 10:16] * log_softmax(lg[:, 10:16])).sum()
    reg = sum(0.5 * (np.asarray(v, dtype=dtype) ** 2).sum()
              for k, v in weights.items() if "bias" not in k) * l2_lambda
    return dict(loss=l1 + l2 + l3 + l4 + reg, loss1=l1, loss2=l2, loss3=l3, loss4=l4, lossL2=reg)


def out16(o):
    """Pack the four head outputs the way the C-ABI returns them: [base4 | zyg2 | type4 | len6]."""
    return np.concatenate([o["base"], o["zygosity"], o["varType"], o["indelLength"]], axis=1)
