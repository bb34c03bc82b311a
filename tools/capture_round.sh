# end-of-round capture on one B200: tests, smoke, bench lines, ncu launch lists + full capture of the forward kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r1n
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > ${P}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> ${P}_tests.log
timeout 300 python bench.py > ${P}_bench_v3.json 2> ${P}_bench_v3.err
timeout 300 python bench.py --variant v3_slim > ${P}_bench_slim.json 2> ${P}_bench_slim.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_ref.json 2>&1
timeout 200 python tools/train_bench.py 20 > ${P}_train_bench.json 2> ${P}_train_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 1 --sites 151552 --cpu-seconds 1 > ${P}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_slab|k_fc4_tc|k_tail_tc|k_v3_c1_reg" --launch-skip 20 -c 5 -o ${P}_full -f python bench.py --steps 2 --warmup 1 --sites 151552 --cpu-seconds 1 > ${P}_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_train_launches.csv python tools/train_profile.py v3 1 > ${P}_ncu3.log 2>&1
cat ${P}_tests.log
