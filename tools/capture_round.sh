set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r1g_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> gpurun_out/r1g_tests.log
timeout 200 python bench.py > gpurun_out/r1g_bench_v3.json 2> gpurun_out/r1g_bench_v3.err
timeout 200 python bench.py --variant v3_slim > gpurun_out/r1g_bench_slim.json 2> gpurun_out/r1g_bench_slim.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1g_bench_ref.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 1 --sites 151552 --cpu-seconds 1 > gpurun_out/r1g_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_slab|k_fc4_tc|k_tail_tc|k_v3_c1" --launch-skip 20 -c 5 -o gpurun_out/r1g_full -f python bench.py --steps 2 --warmup 1 --sites 151552 --cpu-seconds 1 > gpurun_out/r1g_ncu2.log 2>&1
cat gpurun_out/r1g_tests.log
