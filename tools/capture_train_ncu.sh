#!/bin/bash
# ncu --set full over the kernels of one v3 training step (10,000 tensors = two micro-chunks): profiles/r02_train_summary.md
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r02_train
# skip the warm-up step (and the first micro-chunk of the profiled one), then one micro-chunk's worth of launches
timeout 800 ncu --set full --clock-control none --import-source on --launch-skip 150 -c 62 -o ${P}_full -f python tools/train_profile.py v3 1 > ${P}_ncu.log 2>&1
python tools/ncu_summary.py ${P}_full.ncu-rep gpurun_out/r02_train_launches.csv > ${P}_summary.md 2>/dev/null
ls -la ${P}_full.ncu-rep; tail -3 ${P}_ncu.log; grep -c "^### " ${P}_summary.md
