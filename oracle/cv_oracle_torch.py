"""Second, independent CPU restatement of the Clairvoyante v3 / v3_slim graph
on torch-CPU (oneDNN conv, MKL GEMM).  TEST INFRASTRUCTURE ONLY -- see the
header of cv_oracle.py for who may import oracle/ and why parity is unpinned.

Roles:
  * cross-checks cv_oracle.py (NumPy, NHWC, explicit loops) with a different
    formulation (NCHW F.conv2d / F.max_pool2d, explicit F.pad for TF `SAME`);
  * autograd gives reference gradients for the training kernels
    (clairvoyante_v3.py:174 `AdamOptimizer(...).minimize(loss)`);
  * `tf_adam_step` restates TF-1.x Adam (python/training/adam.py in TF 1.12:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; var -= lr_t*m/(sqrt(v)+eps));
  * fp32 + all host threads = the stand-in for "reference TF-CPU on host"
    that bench.py times.
"""
import math
import torch
import torch.nn.functional as F

from . import cv_oracle as O

SELU_ALPHA = O.SELU_ALPHA
SELU_SCALE = O.SELU_SCALE


def selu(x):
    # selu.py:21-25
    return SELU_SCALE * torch.where(x >= 0.0, x, SELU_ALPHA * torch.expm1(torch.clamp(x, max=0.0)))


def to_torch(weights, dtype=torch.float32, requires_grad=False):
    return {k: torch.tensor(v, dtype=dtype, requires_grad=requires_grad) for k, v in weights.items()}


def _dropout_selu(t, rate, mask):
    # selu.py:34-69 with the keep mask given (fixedPointMean 0, fixedPointVar 1)
    keep = 1.0 - rate
    al = O.DROPOUT_ALPHA
    a_ = math.sqrt(1.0 / (keep * ((1.0 - keep) * al * al + 1.0)))
    b_ = -a_ * ((1.0 - keep) * al)
    return a_ * (t * mask + al * (1.0 - mask)) + b_


def forward(W, x, variant="v3", drop4_rate=0.0, drop4_mask=None, head_act=True, drop5_rate=0.0, drop5_mask=None):
    """W: dict name->torch tensor (TF layouts: HWIO / [in,out]); x: (N,33,4,4) NHWC tensor.
    head_act=False is NOT the reference graph: it drops the SELU of the three softmax heads (clairvoyante_v3.py:128-137) and
    exists only for the warm-up phase of tests/golden/make_trained_weights.py (a head whose two pre-activations are both far
    below zero sits at SELU's floor with a vanishing gradient and never recovers)."""
    spec = O.VARIANTS[variant]
    a = x.reshape(-1, O.H_IN, O.W_IN, O.C_IN).permute(0, 3, 1, 2)      # NCHW
    for i, (kh, cout, pool) in enumerate(spec["convs"], 1):
        k = W["conv%d/kernel" % i].permute(3, 2, 0, 1)                 # HWIO -> OIHW
        pt = (kh - 1) // 2
        pb = (kh - 1) - pt
        a = F.pad(a, (1, 2, pt, pb))                                   # W: (1,2), H: (pt,pb)
        a = selu(F.conv2d(a, k, W["conv%d/bias" % i]))
        if pool > 1:
            a = F.max_pool2d(a, kernel_size=(pool, 1), stride=1)
    flat = a.permute(0, 2, 3, 1).reshape(a.shape[0], -1)               # back to (h,w,c) order
    fc4 = selu(flat @ W["fc4/kernel"] + W["fc4/bias"])
    d4 = fc4
    if drop4_rate > 0.0:
        keep = 1.0 - drop4_rate
        al = O.DROPOUT_ALPHA
        a_ = math.sqrt(1.0 / (keep * ((1.0 - keep) * al * al + 1.0)))
        b_ = -a_ * ((1.0 - keep) * al)
        d4 = a_ * (fc4 * drop4_mask + al * (1.0 - drop4_mask)) + b_
    fc5 = selu(d4 @ W["fc5/kernel"] + W["fc5/bias"])
    if drop5_rate > 0.0:                                               # dropout5 feeds the three softmax heads
        fc5 = _dropout_selu(fc5, drop5_rate, drop5_mask)               # (clairvoyante_v3.py:121,128-135)
    base_logit = d4 @ W["YBaseChangeSigmoid/kernel"] + W["YBaseChangeSigmoid/bias"]
    hact = selu if head_act else (lambda t: t)
    zl = hact(fc5 @ W["YZygosityFC/kernel"] + W["YZygosityFC/bias"]) + 1e-10
    tl = hact(fc5 @ W["YVarTypeFC/kernel"] + W["YVarTypeFC/bias"]) + 1e-10
    ll = hact(fc5 @ W["YIndelLengthFC/kernel"] + W["YIndelLengthFC/bias"]) + 1e-10
    return dict(base=torch.sigmoid(base_logit), zygosity=torch.softmax(zl, 1),
                varType=torch.softmax(tl, 1), indelLength=torch.softmax(ll, 1),
                logits=torch.cat([base_logit, zl, tl, ll], 1))


def loss(W, x, y, variant="v3", l2_lambda=0.0, **fw):
    """clairvoyante_v3.py:140-151 (sum over batch; l2_loss = 0.5*sum(v^2) on non-bias vars)."""
    o = forward(W, x, variant, **fw)
    lg = o["logits"]
    l = ((o["base"] - y[:, 0:4]) ** 2).sum()
    l = l + (-y[:, 4:6] * torch.log_softmax(lg[:, 4:6], 1)).sum()
    l = l + (-y[:, 6:10] * torch.log_softmax(lg[:, 6:10], 1)).sum()
    l = l + (-y[:, 10:16] * torch.log_softmax(lg[:, 10:16], 1)).sum()
    reg = sum(0.5 * (v ** 2).sum() for k, v in W.items() if "bias" not in k)
    return l + l2_lambda * reg


def loss_and_grads(weights, x, y, variant="v3", l2_lambda=0.0, dtype=torch.float64, **fw):
    W = to_torch(weights, dtype, requires_grad=True)
    xt = torch.tensor(x, dtype=dtype)
    yt = torch.tensor(y, dtype=dtype)
    for k in ("drop4_mask", "drop5_mask"):
        if fw.get(k) is not None:
            fw = dict(fw, **{k: torch.tensor(fw[k], dtype=dtype)})
    l = loss(W, xt, yt, variant, l2_lambda, **fw)
    l.backward()
    return float(l.detach()), {k: v.grad.numpy() for k, v in W.items()}


def tf_adam_step(var, grad, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF-1.x AdamOptimizer dense update at step t (1-based); NumPy in, NumPy out."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = m + (grad - m) * (1.0 - beta1)
    v = v + (grad * grad - v) * (1.0 - beta2)
    var = var - lr_t * m / (v ** 0.5 + eps)
    return var, m, v


class CpuPredictor(object):
    """fp32, all host threads: the timed stand-in for the reference's
    `Clairvoyante.predict` on TF-CPU (clairvoyante_v3.py:257-267)."""

    def __init__(self, weights, variant="v3", threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.threads = torch.get_num_threads()
        self.W = to_torch(weights, torch.float32)
        self.variant = variant

    @torch.no_grad()
    def predict(self, x):
        o = forward(self.W, torch.from_numpy(x), self.variant)
        return (o["base"].numpy(), o["zygosity"].numpy(), o["varType"].numpy(), o["indelLength"].numpy())
