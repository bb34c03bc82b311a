"""timing of the tensor path under CVB_ABLATE (set in the environment): prints per-kernel ms per 18,944-site launch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clairvoyante_b200 import clairvoyante_v3 as cv, initializers as I, synth
m = cv.Clairvoyante(); m.setWeights(I.init_weights("v3", 0))
N = 18944 * 8
xd = torch.from_numpy(synth.make_sites(18944, 1)).cuda().repeat(8, 1, 1, 1).contiguous(); od = torch.empty((N, 16), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): m.predictDevice(xd.data_ptr(), N, od.data_ptr(), None, st)
torch.cuda.synchronize(); m.profileBegin()
for _ in range(3): m.predictDevice(xd.data_ptr(), N, od.data_ptr(), None, st)
pr = m.profileRead()
print("ABLATE=%s" % os.environ.get("CVB_ABLATE", "0"), {k: round(v[0] / max(v[1], 1), 4) for k, v in pr.items()})
