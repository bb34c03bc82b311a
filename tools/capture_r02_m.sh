#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/train_small_probe.py 2>&1 | tail -1 | tee gpurun_out/r02m_train_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02m_train1250_launches.csv python - > gpurun_out/r02m_ncu.log 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
from clairvoyante_b200 import clairvoyante_v3 as cv, synth
import os
os.environ["CVB_TRAIN_GRAPH"] = "0"
m = cv.Clairvoyante(); m.init(seed=1)
x, y = synth.make_labeled_sites(1250, 3)
for _ in range(3):
    m.train(x, y)
PY
python tools/launch_summary.py gpurun_out/r02m_train1250_launches.csv 2>/dev/null | tail -45
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py 2>&1 | tail -8 | tee gpurun_out/r02m_sanitizer.log
timeout 60 python tools/small_n_probe.py v3_slim 2>&1 | tail -1
