"""A stand-in for the slice of the TensorFlow 1.x API that the reference's model files call -- GENERATION-TIME TOOL for
tests/golden/make_golden_reference_graph.py, not part of the product and not the oracle.

Purpose: execute the reference's OWN graph-building code (clairvoyante/clairvoyante_v3.py, clairvoyante_v3_slim.py, selu.py,
read from /root/reference) without TensorFlow, so that everything those files decide -- layer order and sizes, which tensor
feeds which head, paddings, pool windows, the SELU / dropout formulas of selu.py, the loss composition, which variables are
regularised, what predict / getLoss / train feed for phase, dropout and lambda -- comes from the reference and not from a
restatement.  What does NOT come from the reference are the kernels of the ~25 TensorFlow ops themselves; they are
implemented here on torch (float64) from TensorFlow's documented semantics:

  tf.layers.conv2d     NHWC cross-correlation, HWIO kernel `name/kernel`, bias `name/bias`, stride 1, padding "same" =
                       total k-1 per axis, the smaller half before (TF's SAME rule), then `activation`
  tf.layers.max_pooling2d   window pool_size, strides, padding "valid"
  tf.layers.dense      x @ kernel + bias, kernel [in, out], then `activation`
  tf.nn.softmax / log_softmax / sigmoid / elu / l2_loss (= sum(v^2) / 2), tf.where, tf.pow, tf.slice (size -1 = to the end),
  tf.reduce_sum (all axes), tf.add, tf.add_n, tf.reshape, tf.constant, random_uniform [0, 1), floor, sqrt
  tf.train.AdamOptimizer(lr).minimize   TF-1.x Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m, v updates;
                       var -= lr_t * m / (sqrt(v) + 1e-8), b1 = 0.9, b2 = 0.999, t counted from 1
  smart_cond(pred, f1, f2)   lazily evaluates the branch `pred` selects at run time

The graph is symbolic (Node = a closure over its inputs), Session.run evaluates fetches with a per-run memo so that the loss
that is fetched and the loss that is differentiated see the same dropout mask, and applies optimiser updates after all
fetches were computed (TensorFlow returns the pre-update loss).  `install()` puts the fake modules into sys.modules."""
import sys
import types

import numpy as np
import torch

DT = torch.float64
_STATE = {"graph": None, "uniform_hook": None}


class _Shape(object):
    def __init__(self, dims):
        self.dims = dims

    def assert_is_compatible_with(self, other):
        return None

    def as_list(self):
        return list(self.dims) if self.dims is not None else None


class Node(object):
    dtype = "float"

    def __init__(self, fn, deps=(), name=None, shape=None):
        self.fn, self.deps, self.shape = fn, tuple(deps), shape
        self.name = (name or "node") + ":0"
        self.op = types.SimpleNamespace(name=name or "node")

    def eval(self, ctx):
        if id(self) not in ctx.memo:
            ctx.memo[id(self)] = self.fn(ctx)
        return ctx.memo[id(self)]

    def get_shape(self):
        return _Shape(self.shape)

    def set_shape(self, s):
        return None

    def _bin(self, other, f, name):
        o = as_node(other)
        shape = self.shape if (o.shape is None or (self.shape is not None and len(self.shape) >= len(o.shape))) else o.shape
        return Node(lambda c: f(self.eval(c), o.eval(c)), (self, o), name, shape)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, "add")
    def __radd__(self, o): return as_node(o)._bin(self, lambda a, b: a + b, "add")
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, "sub")
    def __rsub__(self, o): return as_node(o)._bin(self, lambda a, b: a - b, "sub")
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, "mul")
    def __rmul__(self, o): return as_node(o)._bin(self, lambda a, b: a * b, "mul")
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b, "div")
    def __rtruediv__(self, o): return as_node(o)._bin(self, lambda a, b: a / b, "div")
    def __ge__(self, o): return self._bin(o, lambda a, b: a >= b, "ge")
    def __neg__(self): return Node(lambda c: -self.eval(c), (self,), "neg", self.shape)
    __hash__ = object.__hash__


def as_node(v):
    if isinstance(v, Node):
        return v
    t = torch.as_tensor(np.asarray(v, dtype=np.float64), dtype=DT)
    return Node(lambda c: t, (), "const", list(t.shape))


class Variable(Node):
    def __init__(self, name, shape, init):
        Node.__init__(self, lambda c: self.value, (), name, list(shape))
        self.value = torch.tensor(np.asarray(init(shape), np.float64), dtype=DT, requires_grad=True)
        self.m = torch.zeros_like(self.value)
        self.v = torch.zeros_like(self.value)

    def assign(self, a):
        a = np.asarray(a, np.float64)
        assert list(a.shape) == list(self.value.shape), (self.name, a.shape, tuple(self.value.shape))
        self.value = torch.tensor(a, dtype=DT, requires_grad=True)


class Graph(object):
    def __init__(self):
        self.variables, self.summaries, self.step = [], [], 0

    def as_default(self):
        g = self

        class _Ctx(object):
            def __enter__(self_):
                self_.prev, _STATE["graph"] = _STATE["graph"], g
                return g

            def __exit__(self_, *a):
                _STATE["graph"] = self_.prev
        return _Ctx()

    def variable(self, name, shape, init):
        v = Variable(name, shape, init)
        self.variables.append(v)
        return v

    def by_name(self):
        return {v.op.name: v for v in self.variables}


class _RunCtx(object):
    def __init__(self, feed):
        self.feed, self.memo, self.after = feed, {}, []


class Session(object):
    def __init__(self, graph=None, config=None):
        self.graph = graph

    def run(self, fetches, feed_dict=None):
        feed = {}
        for k, v in (feed_dict or {}).items():
            feed[id(k)] = torch.as_tensor(np.asarray(v, dtype=np.float64 if k.dtype == "float" else bool))
        ctx = _RunCtx(feed)
        single = not isinstance(fetches, (tuple, list))
        out = []
        for f in ([fetches] if single else fetches):
            r = f.eval(ctx) if f is not None else None
            out.append(r.detach().numpy().copy() if isinstance(r, torch.Tensor) else r)
        for fn in ctx.after:
            fn()
        return out[0] if single else tuple(out)

    def close(self):
        return None


def placeholder(dtype, shape=None, name=None):
    n = Node(None, (), name, list(shape) if shape is not None else None)
    n.dtype = dtype
    n.fn = lambda c: c.feed[id(n)]
    return n


def _variance_scaling_initializer(factor=2.0, mode="FAN_IN", uniform=False, seed=None, dtype=None):
    def init(shape):      # (every variable is assigned explicitly before it is used; this only has to have the right shape)
        fan_in = int(np.prod(shape[:-1]))
        return np.random.RandomState(0).standard_normal(shape) * np.sqrt(factor / max(fan_in, 1))
    return init


def _glorot(shape):
    lim = np.sqrt(6.0 / (int(np.prod(shape[:-1])) + shape[-1]))
    return np.random.RandomState(1).uniform(-lim, lim, shape)


def _conv2d(inputs, filters, kernel_size, kernel_initializer=None, padding="valid", activation=None, name=None, strides=1):
    g = _STATE["graph"]
    kh, kw = kernel_size
    cin = inputs.shape[-1]
    K = g.variable(name + "/kernel", [kh, kw, cin, filters], kernel_initializer or _glorot)
    B = g.variable(name + "/bias", [filters], lambda s: np.zeros(s))
    assert padding == "same" and strides == 1

    def fn(c):
        x = inputs.eval(c).permute(0, 3, 1, 2)                      # NHWC -> NCHW
        w = K.eval(c).permute(3, 2, 0, 1)                           # HWIO -> OIHW
        ph, pw = kh - 1, kw - 1
        x = torch.nn.functional.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        y = torch.nn.functional.conv2d(x, w) + B.eval(c).view(1, -1, 1, 1)
        return y.permute(0, 2, 3, 1)
    out = Node(fn, (inputs, K, B), name + "/BiasAdd", inputs.shape[:-1] + [filters])
    return activation(out) if activation is not None else out


def _max_pooling2d(inputs, pool_size, strides, padding="valid", name=None):
    ph, pw = pool_size
    assert strides == 1 and padding == "valid"

    def fn(c):
        x = inputs.eval(c).permute(0, 3, 1, 2)
        return torch.nn.functional.max_pool2d(x, (ph, pw), stride=1).permute(0, 2, 3, 1)
    s = inputs.shape
    return Node(fn, (inputs,), name, [s[0], s[1] - ph + 1, s[2] - pw + 1, s[3]])


def _dense(inputs, units, kernel_initializer=None, activation=None, name=None):
    g = _STATE["graph"]
    K = g.variable(name + "/kernel", [inputs.shape[-1], units], kernel_initializer or _glorot)
    B = g.variable(name + "/bias", [units], lambda s: np.zeros(s))
    out = Node(lambda c: inputs.eval(c) @ K.eval(c) + B.eval(c), (inputs, K, B), name + "/BiasAdd", [inputs.shape[0], units])
    return activation(out) if activation is not None else out


def _unary(f, opname):
    def op(x, name=None):
        x = as_node(x)
        return Node(lambda c: f(x.eval(c)), (x,), name or opname, x.shape)
    return op


def _reshape(x, shape, name=None):
    return Node(lambda c: x.eval(c).reshape([int(s) for s in shape]), (x,), name or "reshape", [None if s == -1 else s for s in shape])


def _slice(x, begin, size, name=None):
    def fn(c):
        t = x.eval(c)
        idx = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
        return t[idx]
    return Node(fn, (x,), name or "slice", None)


def _where(cond, a, b):
    a, b = as_node(a), as_node(b)
    return Node(lambda c: torch.where(cond.eval(c), a.eval(c), b.eval(c)), (cond, a, b), "where", a.shape)


def _pow(x, y, name=None):
    x, y = as_node(x), as_node(y)
    return Node(lambda c: torch.pow(x.eval(c), y.eval(c)), (x, y), name or "pow", x.shape)


def _add_n(nodes):
    nodes = [as_node(n) for n in nodes]
    return Node(lambda c: sum(n.eval(c) for n in nodes), nodes, "add_n", [])


def _random_uniform(shape, seed=None, dtype=None):
    def fn(c):
        s = [int(v) for v in shape.eval(c)]
        u = _STATE["uniform_hook"](s) if _STATE["uniform_hook"] is not None else np.random.random_sample(s)
        return torch.as_tensor(np.asarray(u, np.float64), dtype=DT)
    return Node(fn, (shape,), "random_uniform", None)


def _shape_of(x):
    return Node(lambda c: torch.tensor(list(x.eval(c).shape)), (x,), "shape", None)


def _smart_cond(pred, fn1, fn2):
    if not isinstance(pred, Node):
        return fn1() if pred else fn2()
    a, b = fn1(), fn2()
    return Node(lambda c: a.eval(c) if bool(pred.eval(c)) else b.eval(c), (pred,), "cond", a.shape)


class _Adam(object):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = as_node(learning_rate), beta1, beta2, epsilon

    def minimize(self, loss):
        g = _STATE["graph"]

        def fn(c):
            vs = list(g.variables)
            grads = torch.autograd.grad(loss.eval(c), [v.value for v in vs], allow_unused=True)
            lr = float(self.lr.eval(c))

            def apply():
                g.step += 1
                t = g.step
                lr_t = lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
                for v, gr in zip(vs, grads):
                    gr = torch.zeros_like(v.value) if gr is None else gr
                    v.m = self.b1 * v.m + (1 - self.b1) * gr
                    v.v = self.b2 * v.v + (1 - self.b2) * gr * gr
                    v.value = (v.value.detach() - lr_t * v.m / (torch.sqrt(v.v) + self.eps)).requires_grad_(True)
            c.after.append(apply)
            return None
        return Node(fn, (loss,), "Adam", None)


class _NameScope(object):
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return "scope"

    def __exit__(self, *a):
        return False


def install():
    """puts `tensorflow` and the sub-modules selu.py imports into sys.modules; returns the list of names to remove again"""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.bool = "float", "bool"
    tf.Graph, tf.Session, tf.placeholder = Graph, Session, placeholder
    tf.ConfigProto = lambda **kw: None
    tf.layers = types.SimpleNamespace(conv2d=_conv2d, max_pooling2d=_max_pooling2d, dense=_dense)
    tf.nn = types.SimpleNamespace(
        sigmoid=_unary(torch.sigmoid, "sigmoid"), softmax=_unary(lambda t: torch.softmax(t, -1), "softmax"),
        log_softmax=_unary(lambda t: torch.log_softmax(t, -1), "log_softmax"), elu=_unary(torch.nn.functional.elu, "elu"),
        l2_loss=_unary(lambda t: (t * t).sum() / 2, "l2_loss"))
    tf.where, tf.pow, tf.reshape, tf.slice, tf.add_n = _where, _pow, _reshape, _slice, _add_n
    tf.add = lambda a, b, name=None: as_node(a)._bin(b, lambda x, y: x + y, name or "add")
    tf.constant = lambda value, **kw: as_node(value)
    tf.reduce_sum = lambda x, name=None: Node(lambda c: x.eval(c).sum(), (x,), name or "sum", [])
    tf.trainable_variables = lambda: list(_STATE["graph"].variables)
    tf.global_variables_initializer = lambda: Node(lambda c: None, (), "init", None)
    tf.summary = types.SimpleNamespace(scalar=lambda n, t: _STATE["graph"].summaries.append((n, t)), histogram=lambda n, t: None,
                                       merge_all=lambda: Node(lambda c: None, (), "merged", None), FileWriter=lambda *a, **k: None)
    tf.train = types.SimpleNamespace(AdamOptimizer=_Adam, Saver=lambda *a, **k: None)
    contrib_layers = types.ModuleType("tensorflow.contrib.layers")
    contrib_layers.variance_scaling_initializer = _variance_scaling_initializer
    contrib = types.ModuleType("tensorflow.contrib")
    contrib.layers = contrib_layers
    tf.contrib = contrib
    mods = {"tensorflow": tf, "tensorflow.contrib": contrib, "tensorflow.contrib.layers": contrib_layers}

    def sub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        mods[name] = m
        return m
    sub("tensorflow.python")
    sub("tensorflow.python.framework")
    sub("tensorflow.python.ops")
    sub("tensorflow.contrib.layers.python")
    sub("tensorflow.contrib.layers.python.layers")
    mods["tensorflow.python.framework"].ops = sub("tensorflow.python.framework.ops", name_scope=_NameScope,
                                                  convert_to_tensor=lambda v, dtype=None, name=None: as_node(v))
    mods["tensorflow.python.framework"].tensor_shape = sub("tensorflow.python.framework.tensor_shape", scalar=lambda: _Shape([]))
    mods["tensorflow.python.framework"].tensor_util = sub("tensorflow.python.framework.tensor_util",
                                                          constant_value=lambda t: None)   # placeholders are never constant
    mods["tensorflow.python.ops"].math_ops = sub("tensorflow.python.ops.math_ops", floor=_unary(torch.floor, "floor"),
                                                 sqrt=_unary(torch.sqrt, "sqrt"), pow=_pow)
    mods["tensorflow.python.ops"].random_ops = sub("tensorflow.python.ops.random_ops", random_uniform=_random_uniform)
    mods["tensorflow.python.ops"].array_ops = sub("tensorflow.python.ops.array_ops", shape=_shape_of,
                                                  identity=lambda x: Node(lambda c: x.eval(c), (x,), "identity", x.shape))
    mods["tensorflow.contrib.layers.python.layers"].utils = sub("tensorflow.contrib.layers.python.layers.utils", smart_cond=_smart_cond)
    for k, v in mods.items():
        sys.modules[k] = v
    return list(mods)


def set_uniform_hook(fn):
    """fn(shape) -> array of U[0,1) draws used by random_uniform (lets the generator record the dropout noise)"""
    _STATE["uniform_hook"] = fn
