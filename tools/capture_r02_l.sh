#!/bin/bash
mkdir -p gpurun_out
for off in 0 1 2 4 3 6; do
  CVB_PDL_OFF=$off timeout 120 python tools/small_n_probe.py v3_slim 2>&1 | tail -1 | sed "s/^/off=$off /"
done | tee gpurun_out/r02l_probe.log
