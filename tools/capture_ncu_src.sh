cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_slab|k_v3_c1_reg" --launch-skip 12 -c 3 -o gpurun_out/r1k_src -f python bench.py --steps 2 --warmup 1 --sites 151552 --cpu-seconds 1 > gpurun_out/r1k_ncu.log 2>&1
ls -la gpurun_out/r1k_src.ncu-rep
tail -3 gpurun_out/r1k_ncu.log
