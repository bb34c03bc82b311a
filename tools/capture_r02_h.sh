#!/bin/bash
# round 2, call 8: v3_slim tensor pipeline (+ fp16 mode), KCH = 512 split-K, conv2 bias variants
mkdir -p gpurun_out
timeout 120 python - <<'PY' 2>&1 | tail -12 | tee gpurun_out/r02h_slim_first.log
import numpy as np, sys
sys.path.insert(0, ".")
from clairvoyante_b200 import clairvoyante_v3_slim as cv, initializers as I, synth
from oracle import cv_oracle as O
W = I.init_weights("v3_slim", 3); x = synth.make_sites(300, 5)
ref = O.forward(W, x, "v3_slim", return_all=True)
m = cv.Clairvoyante(); m.setWeights(W)
for mode in ("fp16x3", "fp16", "fp32"):
    m.setComputeMode(mode)
    o, lg = m.predictLogits(x)
    print(mode, "max |logit - oracle| =", float(np.abs(lg - ref["logits"]).max()), "max|logit|", float(np.abs(ref["logits"]).max()))
m.close()
PY
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_trained_parity.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -14 | tee gpurun_out/r02h_tests.log
timeout 90 python tools/ab_resident.py v3 1 2>&1 | tail -1 | tee gpurun_out/r02h_ab.log
timeout 90 python tools/ab_resident.py slim 1 2>&1 | tail -1 | tee -a gpurun_out/r02h_ab.log
CVB_COMPUTE=fp16 timeout 90 python tools/ab_resident.py slim 1 CVB_COMPUTE=fp16 2>&1 | tail -1 | tee -a gpurun_out/r02h_ab.log
CVB_SLIM_TC=0 timeout 90 python tools/ab_resident.py slim 1 CVB_SLIM_TC=0 2>&1 | tail -1 | tee -a gpurun_out/r02h_ab.log
