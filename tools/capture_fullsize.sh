#!/bin/bash
# the full-size property tests at BASELINE configs[1]'s 4 Mi sites (the suite's default is 1 Mi) and a second default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CVB_FULLSIZE_SITES=4194304 timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_fullsize_4mi.log
timeout 600 python bench.py > gpurun_out/r02_bench_run2.json 2> gpurun_out/r02_bench_run2.err
cat gpurun_out/r02_fullsize_4mi.log; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_run2.json')); print(d['value'], d['e2e']['value'], d['slim']['fp16x3']['value'], d['slim']['fp16']['value'], d['train']['global_batch_10000']['value'], d['small_batch']['host_api_counts_pipelined']['value'], d['clocks'])"
