"""GPU probe for the tensor-core path: fp32 SIMT vs split-fp16 tcgen05 FC4 vs oracle, plus
structured-weight cases that localise descriptor / swizzle mistakes.  Dumps to gpurun_out/."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from clairvoyante_b200 import initializers as I, synth, clairvoyante_v3 as cv
from oracle import cv_oracle as O

os.makedirs("gpurun_out", exist_ok=True)
W = I.init_weights("v3", 6)
n = 300
x = synth.make_sites(n, 8)
ref = O.forward(W, x, "v3", return_all=True)
m = cv.Clairvoyante()
m.setWeights(W)
o32, l32 = m.predictLogits(x)
h4_32 = m.debugRead("h4", n)
print("fp32   : h4 err %.3g logit err %.3g" % (np.abs(h4_32 - ref["layers"]["fc4"]).max(), np.abs(l32 - ref["logits"]).max()))
m.setComputeMode("fp16x3")
try:
    otc, ltc = m.predictLogits(x)
    h4_tc = m.debugRead("h4", n)
    L = ref["layers"]
    if os.environ.get("CVB_TC_CONV3", "1") != "0":
        p2s = m.debugRead("p2_split", n).reshape(n, 28, 4, 32)
        print("   p2 split err %.3g pad rows zero %s" % (np.abs(p2s[:, 1:27] - L["pool2"]).max(), bool((p2s[:, 0] == 0).all() and (p2s[:, 27] == 0).all())))
    p3s = m.debugRead("p3_split", n).reshape(n, 24, 4, 48)
    e3 = np.abs(p3s - L["pool3"])
    print("   p3 split err %.3g (mean %.3g); by h: %s" % (e3.max(), e3.mean(), np.round(e3.max((0, 2, 3)), 4)))
    print("      by w: %s ; by site%%8: %s" % (np.round(e3.max((0, 1, 3)), 4), np.round([e3[i::8].max() for i in range(8)], 4)))
    np.savez_compressed("gpurun_out/tc_probe_p3.npz", p3s=p3s, ref=L["pool3"])
    e = np.abs(h4_tc - ref["layers"]["fc4"])
    print("fp16x3 : h4 err %.3g (mean %.3g) logit err %.3g ; vs fp32-simt h4 %.3g" % (e.max(), e.mean(), np.abs(ltc - ref["logits"]).max(), np.abs(h4_tc - h4_32).max()))
    print("   err by column block of 16:", np.round(e.max(0).reshape(21, 16).max(1), 4))
    print("   err by site block of 32  :", np.round(e.max(1)[:288].reshape(9, 32).max(1), 4))
    np.savez_compressed("gpurun_out/tc_probe.npz", h4_tc=h4_tc, h4_32=h4_32, ref=ref["layers"]["fc4"])
    if e.max() > 1e-3:
        # structured: W4 = one-hot in k -> h4[:, j] = selu(p3[:, k0])
        for k0 in (0, 1, 8, 16, 31, 32, 33, 64, 100, 4607):
            W2 = dict(W); w4 = np.zeros_like(W["fc4/kernel"]); w4[k0, :] = 1.0; W2["fc4/kernel"] = w4
            W2["fc4/bias"] = np.zeros(336, np.float32)
            m.setWeights(W2); m.predictLogits(x); h = m.debugRead("h4", n)
            want = O.selu(ref["layers"]["flat"][:, k0])
            # which k does the kernel appear to pick?
            flat = ref["layers"]["flat"]
            d = np.abs(O.selu(flat)[:, :, None] - h[:, None, :1]).sum(0)[:, 0]
            print("   one-hot k0=%d: err %.3g ; best matching k = %d" % (k0, np.abs(h - want[:, None]).max(), int(d.argmin())))
    # throughput
    N = 18944 * 8
    xd = torch.from_numpy(synth.make_sites(18944, 1)).cuda().repeat(8, 1, 1, 1).contiguous()
    od = torch.empty((N, 16), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    m.setWeights(W)
    for mode in ("fp32", "fp16x3"):
        m.setComputeMode(mode)
        for _ in range(2):
            m.predictDevice(xd.data_ptr(), N, od.data_ptr(), None, st)
        torch.cuda.synchronize()
        m.profileBegin()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            m.predictDevice(xd.data_ptr(), N, od.data_ptr(), None, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        pr = m.profileRead()
        print("%s: %.3f ms per %d sites -> %.2f M sites/s ; per-kernel ms/launch %s" % (mode, ms, N, N / ms / 1e3, {k: round(v[0] / max(v[1], 1), 4) for k, v in pr.items()}))
except Exception as ex:
    print("fp16x3 FAILED:", repr(ex))
m.close()
