"""TEST INFRASTRUCTURE.  The forward graphs of clairvoyante_v3 / clairvoyante_v3_slim as TensorFlow GraphDef protobufs, built
with the TF proto schemas that ship with TensorBoard (no TensorFlow in this image), so that an INDEPENDENT executor of
TensorFlow graph semantics -- OpenCV's dnn TensorFlow importer, cv2.dnn.readNetFromTensorflow -- can run them:

    Conv2D(padding=SAME, NHWC) + BiasAdd     clairvoyante_v3.py:54-60,69-75,84-90   (tf.layers.conv2d)
    MaxPool(ksize=(p,1), strides 1, VALID)   clairvoyante_v3.py:63-66,78-81,93-96   (tf.layers.max_pooling2d)
    Reshape([-1, flat])                      clairvoyante_v3.py:98-99               (NHWC flatten order)
    MatMul + BiasAdd                         clairvoyante_v3.py:101-137             (tf.layers.dense)
    Sigmoid / Softmax heads                  clairvoyante_v3.py:124-137
    selu                                     selu.py:21-25, written as scale*alpha*Elu(x) + scale*(1-alpha)*Relu(x)
                                             (the importer has no Select op; the identity is exact: elu(x) = relu(x) = x for
                                             x >= 0, relu(x) = 0 for x < 0)
    phase = False: dropout_selu is the identity (selu.py:48-69 through utils.smart_cond)

What this pins: the TensorFlow op-kernel semantics the oracle restates from documentation -- the SAME rule for even kernels
(one row / column more padding at the bottom / right), VALID pooling, the flatten order, the head wiring -- against an
implementation written by neither the reference's authors nor this repo's.  The wiring itself is pinned separately by the
reference's own graph code run on tf1_stand_in.py (make_golden_reference_graph.py)."""
import numpy as np
from tensorboard.compat.proto import graph_pb2, tensor_pb2, tensor_shape_pb2, types_pb2

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


def _shape(dims):
    return tensor_shape_pb2.TensorShapeProto(dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(d)) for d in dims])


class G(object):
    def __init__(self):
        self.g = graph_pb2.GraphDef()

    def node(self, name, op, inputs=(), **attrs):
        n = self.g.node.add()
        n.name, n.op = name, op
        n.input.extend(inputs)
        if op not in ("Const", "Placeholder"):
            n.attr["T"].type = types_pb2.DT_FLOAT
        for k, v in attrs.items():
            if isinstance(v, bytes):
                n.attr[k].s = v
            elif isinstance(v, (list, tuple)):
                n.attr[k].list.i.extend(v)
        return name

    def const(self, name, arr):
        arr = np.asarray(arr)
        n = self.g.node.add()
        n.name, n.op = name, "Const"
        dt = types_pb2.DT_FLOAT if arr.dtype == np.float32 else types_pb2.DT_INT32
        n.attr["dtype"].type = dt
        n.attr["value"].tensor.CopyFrom(tensor_pb2.TensorProto(dtype=dt, tensor_shape=_shape(arr.shape), tensor_content=arr.tobytes()))
        return name

    def placeholder(self, name, shape):
        n = self.g.node.add()
        n.name, n.op = name, "Placeholder"
        n.attr["dtype"].type = types_pb2.DT_FLOAT
        n.attr["shape"].shape.CopyFrom(_shape(shape))
        return name

    def selu(self, name, x):
        # selu.py:21-25  scale * where(x >= 0, x, alpha * elu(x))  ==  scale*alpha*elu(x) + scale*(1-alpha)*relu(x)
        e = self.node(name + "/elu", "Elu", [x])
        r = self.node(name + "/relu", "Relu", [x])
        ce = self.const(name + "/ce", np.float32(SELU_SCALE * SELU_ALPHA))
        cr = self.const(name + "/cr", np.float32(SELU_SCALE * (1.0 - SELU_ALPHA)))
        me = self.node(name + "/me", "Mul", [e, ce])
        mr = self.node(name + "/mr", "Mul", [r, cr])
        return self.node(name, "Add", [me, mr])

    def conv(self, name, x, k, b):
        w = self.const(name + "/kernel", k.astype(np.float32))
        bb = self.const(name + "/bias", b.astype(np.float32))
        c = self.node(name + "/Conv2D", "Conv2D", [x, w], strides=[1, 1, 1, 1], dilations=[1, 1, 1, 1], padding=b"SAME", data_format=b"NHWC")
        return self.node(name + "/BiasAdd", "BiasAdd", [c, bb], data_format=b"NHWC")

    def pool(self, name, x, p):
        return self.node(name, "MaxPool", [x], ksize=[1, p, 1, 1], strides=[1, 1, 1, 1], padding=b"VALID", data_format=b"NHWC")

    def dense(self, name, x, k, b):
        w = self.const(name + "/kernel", k.astype(np.float32))
        bb = self.const(name + "/bias", b.astype(np.float32))
        m = self.node(name + "/MatMul", "MatMul", [x, w])
        return self.node(name + "/BiasAdd", "BiasAdd", [m, bb])


def forward_graph(W, variant, n):
    """clairvoyante_v3.py:54-137 / clairvoyante_v3_slim.py:54-118 at phase = False (dropout is the identity)"""
    g = G()
    x = g.placeholder("X", [n, 33, 4, 4])
    pools = (5, 4, 3) if variant == "v3" else (1, 1, 1)
    h = x
    for i, p in zip((1, 2, 3), pools):
        h = g.selu("selu%d" % i, g.conv("conv%d" % i, h, W["conv%d/kernel" % i], W["conv%d/bias" % i]))
        if p > 1:
            h = g.pool("pool%d" % i, h, p)
    flat = W["fc4/kernel"].shape[0]
    shp = g.const("flat/shape", np.array([-1, flat], np.int32))
    f = g.node("flat", "Reshape", [h, shp])
    h4 = g.selu("selu4", g.dense("fc4", f, W["fc4/kernel"], W["fc4/bias"]))
    h5 = g.selu("selu5", g.dense("fc5", h4, W["fc5/kernel"], W["fc5/bias"]))
    base = g.dense("YBaseChangeSigmoid", h4, W["YBaseChangeSigmoid/kernel"], W["YBaseChangeSigmoid/bias"])   # :125, from the FC4 branch
    outs, logits = [g.node("YBase", "Sigmoid", [base])], [base]
    eps = g.const("eps", np.float32(1e-10))
    for nm in ("YZygosityFC", "YVarTypeFC", "YIndelLengthFC"):
        s = g.selu(nm + "/selu", g.dense(nm, h5, W[nm + "/kernel"], W[nm + "/bias"]))
        lg = g.node(nm + "/eps", "Add", [s, eps])                                                             # :127-136
        logits.append(lg)
        outs.append(g.node(nm + "/softmax", "Softmax", [lg]))
    return g.g.SerializeToString(), outs, logits


def layer_graph(kernel, bias, pool, n):
    """one conv2d(SAME) + bias [+ max-pool (pool,1) VALID] on an [n,33,4,cin] input"""
    g = G()
    h = g.conv("conv", g.placeholder("X", [n, 33, 4, kernel.shape[2]]), kernel, bias)
    if pool > 1:
        g.pool("pool", h, pool)
    return g.g.SerializeToString()
