cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t1250_launches.csv python tools/train_profile.py v3 1 1250 > gpurun_out/t1250_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/t1250_launches.csv 0.5 > gpurun_out/t1250_summary.txt 2>&1
timeout 200 python tools/train_small_probe.py 2>&1 | tail -2 > gpurun_out/t_probe.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
cat gpurun_out/t1250_summary.txt | head -60; cat gpurun_out/t_probe.log; python -c "
import json; d=json.load(open('gpurun_out/check_bench.json')); print(json.dumps(d['small_batch'])); print(d['value'], d['e2e']['value'], d['train']['global_batch_10000']['value'])"
