"""per-call device time of small batches around the reference's predictBatchSize, with and without PDL (CVB_PDL is read once per
process: run twice).   python tools/small_n_probe.py <variant>"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clairvoyante_b200 import clairvoyante_v3 as cv, clairvoyante_v3_slim as cvs, initializers as I, synth
variant = sys.argv[1] if len(sys.argv) > 1 else "v3_slim"
m = (cv if variant == "v3" else cvs).Clairvoyante()
m.setWeights(I.init_weights(variant, 0))
if len(sys.argv) > 2:
    m.setComputeMode(sys.argv[2])
pool = synth.make_sites(4096, 1)
xd = torch.from_numpy(pool).cuda(); od = torch.empty((4096, 16), device="cuda")
st = torch.cuda.current_stream().cuda_stream
out = {}
for n in (896, 960, 992, 999, 1000, 1001, 1008, 1016, 1023, 1024, 1025, 1056, 1152):
    for _ in range(5):
        m.predictDevice(xd.data_ptr(), n, od.data_ptr(), None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(200):
        m.predictDevice(xd.data_ptr(), n, od.data_ptr(), None, st)
    e1.record(); torch.cuda.synchronize()
    out[n] = round(e0.elapsed_time(e1) * 1e3 / 200, 1)
print(variant, m.computeMode, "PDL", os.environ.get("CVB_PDL", "1"), json.dumps(out))
