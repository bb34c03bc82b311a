"""Generates tests/golden/opencv_tf_forward.npz: the outputs of OpenCV's TensorFlow importer (cv2.dnn, an independent
executor of TF graph semantics) on the forward GraphDefs of tf_graphdef.py, for the committed trained-like weights and for
initialiser weights, 64 synthetic sites each.      python tests/golden/make_golden_opencv.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import cv2  # noqa: E402
from tf_graphdef import forward_graph, layer_graph  # noqa: E402
from clairvoyante_b200 import initializers as I, synth  # noqa: E402

N, SEED = 64, 11
LAYERS = ((1, 4, 16, 5), (2, 16, 32, 4), (3, 32, 48, 3), (1, 4, 8, 1), (3, 8, 16, 1), (5, 16, 32, 1))   # kh, cin, cout, pool


def run_pb(pb, feeds, fetch):
    with tempfile.NamedTemporaryFile(suffix=".pb") as f:      # (the importer's in-memory overload crashes in this build)
        f.write(pb)
        f.flush()
        net = cv2.dnn.readNetFromTensorflow(f.name)
    net.setInput(np.ascontiguousarray(feeds.transpose(0, 3, 1, 2)))       # OpenCV blobs are NCHW
    # (the importer fuses BiasAdd into the layer of the MatMul / Conv2D before it: that layer's output is the biased value)
    ys = net.forward([f.replace("/BiasAdd", "/MatMul") for f in fetch]) if fetch else [net.forward()]
    return [np.asarray(y) for y in ys]


def trained(variant):
    with np.load(os.path.join(HERE, "trained_%s.npz" % variant)) as z:
        return {k.replace(".", "/"): z[k].astype(np.float32) for k in z.files}


def main():
    out = dict(n=np.int64(N), data_seed=np.int64(SEED), opencv=np.bytes_(cv2.__version__))
    x, _ = synth.make_labeled_sites(N, SEED)
    for variant in ("v3", "v3_slim"):
        for tag, W in (("trained", trained(variant)), ("init", I.init_weights(variant, 4))):
            pb, outs, logits = forward_graph(W, variant, N)
            ys = run_pb(pb, x, list(outs) + list(logits))
            out["%s_%s_out16" % (variant, tag)] = np.concatenate([y.reshape(N, -1) for y in ys[:4]], 1)
            out["%s_%s_logits" % (variant, tag)] = np.concatenate([y.reshape(N, -1) for y in ys[4:]], 1)
    # the inputs and weights of reference_graph.npz (the reference's own predict() on the stand-in): same sites, seed-0 initialisers
    with np.load(os.path.join(HERE, "reference_graph.npz")) as z:
        xr = z["x"].astype(np.float32)
    for variant in ("v3", "v3_slim"):
        pb, outs, logits = forward_graph(I.init_weights(variant, 0), variant, len(xr))
        ys = run_pb(pb, xr, list(outs))
        out["%s_refgraph_out16" % variant] = np.concatenate([y.reshape(len(xr), -1) for y in ys], 1)
    rng = np.random.default_rng(SEED)
    for i, (kh, cin, cout, pool) in enumerate(LAYERS):
        k = rng.standard_normal((kh, 4, cin, cout)).astype(np.float32)
        b = rng.standard_normal(cout).astype(np.float32)
        xi = rng.standard_normal((2, 33, 4, cin)).astype(np.float32)
        pb = layer_graph(k, b, pool, 2)
        y = run_pb(pb, xi, None)[0].transpose(0, 2, 3, 1)
        out["layer%d_k" % i], out["layer%d_b" % i], out["layer%d_x" % i], out["layer%d_y" % i] = k, b, xi, y
    np.savez_compressed(os.path.join(HERE, "opencv_tf_forward.npz"), **out)
    print("wrote opencv_tf_forward.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    main()
