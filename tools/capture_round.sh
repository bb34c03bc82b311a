#!/bin/bash
# end-of-round capture on one B200 (gpurun --timeout 2400 -- bash tools/capture_round.sh): GPU tests, smoke, the bench lines,
# an ncu launch list of the bench command, `ncu --set full` of the forward kernels of both variants, and from that report
# profiles/traffic.json stamped with the hash of THIS build (bench.py refuses a traffic file from another build)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r02
H=$(python -c "from clairvoyante_b200 import _lib; print(_lib.source_hash())")
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > ${P}_tests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 >> ${P}_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 1 --sites 151552 --cpu-seconds 1 --no-extra > ${P}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_conv_slab|k_fc4_tc|k_tail_tc|k_v3_c1_reg" --launch-skip 10 -c 5 -o ${P}_full_v3 -f python bench.py --steps 1 --warmup 1 --sites 151552 --cpu-seconds 1 --no-extra > ${P}_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_conv_slab|k_gemm_tc|k_tail|k_slim_c1_reg" --launch-skip 10 -c 5 -o ${P}_full_slim -f python bench.py --variant v3_slim --steps 1 --warmup 1 --sites 265216 --cpu-seconds 1 --no-extra > ${P}_ncu3.log 2>&1
python tools/ncu_summary.py --traffic ${P}_full_v3.ncu-rep ${P}_traffic_v3.json $H > /dev/null
python tools/ncu_summary.py --traffic ${P}_full_slim.ncu-rep ${P}_traffic_slim.json $H > /dev/null
python - <<PY
import json
a = json.load(open("${P}_traffic_v3.json")); b = json.load(open("${P}_traffic_slim.json"))
a["v3_slim"] = b["v3_slim"]; a["_detail"].update(b["_detail"]); a["report"] = [a["report"], b["report"]]
json.dump(a, open("profiles/traffic.json", "w"), indent=1)
json.dump(a, open("${P}_traffic.json", "w"), indent=1)
PY
python tools/ncu_summary.py ${P}_full_v3.ncu-rep ${P}_launches.csv > ${P}_v3_summary.md 2>/dev/null
python tools/ncu_summary.py ${P}_full_slim.ncu-rep > ${P}_slim_summary.md 2>/dev/null
timeout 600 python bench.py > ${P}_bench.json 2> ${P}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_ref.json 2>&1
timeout 400 python tools/batch_sweep.py > ${P}_sweep.log 2>&1; cp gpurun_out/batch_sweep.json ${P}_batch_sweep.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_train_launches.csv python tools/train_profile.py v3 1 > ${P}_ncu4.log 2>&1
cat ${P}_tests.log; tail -c 400 ${P}_bench.json
