cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "golden" 2>&1 | tail -4 > gpurun_out/r1j_tests.log
timeout 200 python bench.py --steps 3 --warmup 3 --sites 1212416 --cpu-seconds 1 > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err
cat gpurun_out/r1j_tests.log
python - <<'PY'
import json
for f in ("gpurun_out/r1j_bench.json", "gpurun_out/r1j_bench_old.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k: round(v["ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r1j_bench.err
