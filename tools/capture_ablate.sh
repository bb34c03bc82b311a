cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/ablate.log
for a in 0 14 6 8 16 32 48; do
  CVB_ABLATE=$a timeout 100 python tools/ablate.py >> gpurun_out/ablate.log 2>&1
done
grep ABLATE gpurun_out/ablate.log
