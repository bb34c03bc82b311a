cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CVB_TRAIN_SLIM_TC=1 timeout 200 python -m pytest tests/test_train_gpu.py -m gpu -q -k "slim" 2>&1 | grep -E "^E  |passed|failed|Error" | head -12 > gpurun_out/r1s_tests.log
CVB_TRAIN_SLIM_TC=1 timeout 60 python tools/train_profile.py v3_slim 20 >> gpurun_out/r1s_tests.log 2>&1
cat gpurun_out/r1s_tests.log | tail -14
