"""A/B of the resident-weight conv kernels in the TRAINING step (CVB_CONV_RESIDENT bit 4): gradients of one step without
update against the default kernels (same MMAs in the same order; the atomically added split-K partial sums make the last bits
vary from run to run, so the largest relative difference is what to read -- expect <= 1e-6), then ms per step.  One setting per process:
    python tools/ab_resident_train.py <variant: v3|slim> <0|4>     # writes gpurun_out/ab_resident_train_<variant>_<setting>.npz/json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
variant, setting = sys.argv[1], sys.argv[2]
os.environ["CVB_CONV_RESIDENT"] = setting
import numpy as np  # noqa: E402
from clairvoyante_b200 import initializers as I, synth  # noqa: E402
if variant == "v3":
    from clairvoyante_b200 import clairvoyante_v3 as cv
else:
    from clairvoyante_b200 import clairvoyante_v3_slim as cv

os.makedirs("gpurun_out", exist_ok=True)
name = "v3" if variant == "v3" else "v3_slim"
m = cv.Clairvoyante(dropoutRateFC4=0.0)
m.init(seed=1)
m.setWeights(I.init_weights(name, 0))
x, y = synth.make_sites(2777, 5), synth.make_labels(2777, 5)
m._train_step(x, y, apply_update=0, seed=7)
g = m.getGradients()
np.savez("gpurun_out/ab_resident_train_%s_%s.npz" % (variant, setting), **{k.replace("/", "__"): v for k, v in g.items()})
ref_fn = "gpurun_out/ab_resident_train_%s_0.npz" % variant
same = worst = None
if setting != "0" and os.path.exists(ref_fn):
    r = np.load(ref_fn)
    same = all(np.array_equal(r[k.replace("/", "__")], v) for k, v in g.items())
    # (split-K partial sums are added with atomics, so even two runs of the default kernels differ in the last bits)
    worst = max(float(np.abs(r[k.replace("/", "__")] - v).max() / max(float(np.abs(v).max()), 1e-30)) for k, v in g.items())
xb, yb = synth.make_sites(10000, 6), synth.make_labels(10000, 6)
for _ in range(3):
    m.train(xb, yb)
t = time.perf_counter()
for _ in range(10):
    m.train(xb, yb)
ms = (time.perf_counter() - t) / 10 * 1e3
out = {"variant": variant, "CVB_CONV_RESIDENT": setting, "gradients_bit_identical_to_default": same, "max_rel_gradient_diff": worst, "ms_per_10000_tensor_step": round(ms, 3)}
print(json.dumps(out))
json.dump(out, open("gpurun_out/ab_resident_train_%s_%s.json" % (variant, setting), "w"))
m.close()
