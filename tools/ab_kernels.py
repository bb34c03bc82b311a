"""A/B of one build against a saved baseline on one B200: per-kernel ms per chunk (cvb_profile_* events) and whether the logits
are BIT-IDENTICAL to the run tagged "0" (same MMAs in the same order on the same operands), else the largest difference.
Round 2 used it for every kernel experiment of profiles/r02_ab_log.md (then called ab_resident.py; its first subject, the
CVB_CONV_RESIDENT switch, is gone: conv3 keeps its taps resident, conv2 streams them).  Run every setting under its own
timeout on the GPU box: a mis-programmed descriptor hangs the kernel, it does not fail.
    python tools/ab_kernels.py <variant: v3|slim> <tag> [ENV=VALUE,...]      # writes gpurun_out/ab_<variant>_<tag>.npy/json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
variant, setting = sys.argv[1], sys.argv[2]
extra = sys.argv[3] if len(sys.argv) > 3 else ""          # diagnostic switches for this run, e.g. CVB_PDL=0,CVB_COMPUTE=fp16
for kv in filter(None, extra.split(",")):
    k, v = kv.split("=")
    os.environ[k] = v
tag = setting + ("_" + extra.replace("=", "").replace(",", "_") if extra else "")
import numpy as np  # noqa: E402
import torch  # noqa: E402
from clairvoyante_b200 import initializers as I, synth  # noqa: E402
if variant == "v3":
    from clairvoyante_b200 import clairvoyante_v3 as cv
else:
    from clairvoyante_b200 import clairvoyante_v3_slim as cv

os.makedirs("gpurun_out", exist_ok=True)
m = cv.Clairvoyante()
m.setWeights(I.init_weights("v3" if variant == "v3" else "v3_slim", 0))
chunk = 18944 if variant == "v3" else 33152
x = synth.make_sites(chunk + 1234, 1)          # one full chunk and a ragged one
_, logits = m.predictLogits(x)
np.save("gpurun_out/ab_%s_%s.npy" % (variant, tag), logits)
ref_fn = "gpurun_out/ab_%s_0.npy" % variant
same = bool(np.array_equal(np.load(ref_fn), logits)) if tag != "0" and os.path.exists(ref_fn) else None
maxdiff = float(np.abs(np.load(ref_fn) - logits).max()) if tag != "0" and os.path.exists(ref_fn) else None

N = chunk * 8
xd = torch.from_numpy(x[:chunk]).cuda().repeat(8, 1, 1, 1).contiguous()
od = torch.empty((N, 16), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    m.predictDevice(xd.data_ptr(), N, od.data_ptr(), None, st)
torch.cuda.synchronize()
m.profileBegin()
for _ in range(4):
    m.predictDevice(xd.data_ptr(), N, od.data_ptr(), None, st)
pr = m.profileRead()
out = {"variant": variant, "tag": setting, "extra": extra, "bit_identical_to_default": same, "max_abs_logit_diff": maxdiff,
       "ms_per_chunk": {k: round(v[0] / max(v[1], 1), 4) for k, v in pr.items()}}
print(json.dumps(out))
json.dump(out, open("gpurun_out/ab_%s_%s.json" % (variant, tag), "w"))
m.close()
