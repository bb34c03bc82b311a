#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/r02r_tests.log
