cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/train
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > ${P}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> ${P}_tests.log
timeout 200 python tools/train_bench.py 20 > ${P}_train_bench.json 2> ${P}_train_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_train_launches.csv python tools/train_profile.py v3 1 > ${P}_ncu3.log 2>&1
cat ${P}_tests.log ${P}_train_bench.json
