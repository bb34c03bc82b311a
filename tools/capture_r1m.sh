cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py -m gpu -q -x -k "golden or gradients or tensor_and_simt" 2>&1 | tail -3 > gpurun_out/r1m_tests.log
timeout 200 python bench.py --steps 3 --warmup 3 --sites 1212416 --cpu-seconds 1 > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err
timeout 120 python tools/train_profile.py v3 20 > gpurun_out/r1m_train.log 2>&1
cat gpurun_out/r1m_tests.log gpurun_out/r1m_train.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r1m_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v["ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
PY
