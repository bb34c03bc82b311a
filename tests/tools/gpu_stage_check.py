"""First-light check on a GPU box: per-stage error of the CUDA forward against the oracle,
plus a quick device-resident throughput number.  Not a test, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from clairvoyante_b200 import initializers as I, synth
from clairvoyante_b200 import clairvoyante_v3, clairvoyante_v3_slim
from oracle import cv_oracle as O

for variant, mod in (("v3", clairvoyante_v3), ("v3_slim", clairvoyante_v3_slim)):
    W = I.init_weights(variant, 6)
    x = synth.make_sites(200, 8)
    m = mod.Clairvoyante()
    m.setWeights(W)
    out16, lg = m.predictLogits(x)
    ref = O.forward(W, x, variant, return_all=True)
    L = ref["layers"]
    if variant == "v3":
        p2 = m.debugRead("p2", 200).reshape(200, 28, 4, 32)
        print(variant, "p2 pad rows zero:", bool((p2[:, 0] == 0).all() and (p2[:, 27] == 0).all()),
              "p2 err", np.abs(p2[:, 1:27] - L["pool2"]).max())
        p3 = m.debugRead("p3", 200).reshape(200, 24, 4, 48)
        print(variant, "p3 err", np.abs(p3 - L["pool3"]).max())
    else:
        p2 = m.debugRead("p2", 200).reshape(200, 37, 4, 16)
        print(variant, "p2 err", np.abs(p2[:, 2:35] - L["conv2"]).max())
        p3 = m.debugRead("p3", 200).reshape(200, 33, 4, 32)
        print(variant, "p3 err", np.abs(p3 - L["conv3"]).max())
    h4 = m.debugRead("h4", 200)
    print(variant, "h4 err", np.abs(h4 - L["fc4"]).max())
    print(variant, "logit err", np.abs(lg - ref["logits"]).max(), "out16 err", np.abs(out16 - O.out16(ref)).max())
    # quick throughput, device resident
    n = 262144
    xd = torch.from_numpy(synth.make_sites(65536, 1)).cuda().repeat(4, 1, 1, 1).contiguous()
    od = torch.empty((n, 16), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        m.predictDevice(xd.data_ptr(), n, od.data_ptr(), None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        m.predictDevice(xd.data_ptr(), n, od.data_ptr(), None, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(variant, "device-resident: %.3f ms per %d sites -> %.3f M sites/s" % (ms, n, n / ms / 1e3))
    xh = synth.make_sites(65536, 1)
    t = time.time(); m.predict(xh); t1 = time.time() - t
    t = time.time(); m.predict(xh); t2 = time.time() - t
    print(variant, "host path pageable 65536 sites: %.1f ms first, %.1f ms second -> %.3f M sites/s" % (t1 * 1e3, t2 * 1e3, 65536 / t2 / 1e6))
    m.close()
