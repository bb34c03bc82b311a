cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r1p_tests.log
timeout 120 python tools/train_profile.py v3 20 > gpurun_out/r1p_train.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1p_train_launches.csv python tools/train_profile.py v3 1 > gpurun_out/r1p_ncu.log 2>&1
cat gpurun_out/r1p_tests.log gpurun_out/r1p_train.log
