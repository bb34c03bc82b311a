#!/bin/bash
# gpurun --timeout 600 -- bash tools/capture_resident.sh     (each setting under its own timeout: a wrong descriptor hangs)
set -o pipefail
mkdir -p gpurun_out
for v in v3 slim; do
  for s in 0 1 2 3; do
    [ "$v" = slim ] && [ "$s" -ge 2 ] && continue
    timeout 90 python tools/ab_resident.py $v $s 2>&1 | tail -2 || echo "{\"variant\": \"$v\", \"CVB_CONV_RESIDENT\": \"$s\", \"failed\": true}"
  done
  [ "$v" = slim ] && { timeout 90 python tools/ab_resident.py slim 0 CVB_SLIM_FC4_TC=1 2>&1 | tail -2 || echo "{\"variant\": \"slim\", \"extra\": \"CVB_SLIM_FC4_TC=1\", \"failed\": true}"; }
  for s in 0 4; do
    timeout 120 python tools/ab_resident_train.py $v $s 2>&1 | tail -1 || echo "{\"variant\": \"$v\", \"train\": true, \"CVB_CONV_RESIDENT\": \"$s\", \"failed\": true}"
  done
done | tee gpurun_out/ab_resident.log
