set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1h_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> gpurun_out/r1h_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_train_launches.csv python tools/train_profile.py v3 1 > gpurun_out/r1h_train_ncu.log 2>&1
timeout 120 python tools/train_profile.py v3 10 > gpurun_out/r1h_train.log 2>&1
cat gpurun_out/r1h_tests.log gpurun_out/r1h_train.log
