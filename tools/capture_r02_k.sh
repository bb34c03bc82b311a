#!/bin/bash
mkdir -p gpurun_out
for v in v3_slim v3; do
  timeout 120 python tools/small_n_probe.py $v 2>&1 | tail -1
  CVB_PDL=0 timeout 120 python tools/small_n_probe.py $v 2>&1 | tail -1
done | tee gpurun_out/r02k_probe.log
