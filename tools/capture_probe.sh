cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for k in 0 8 1; do
  echo "== kshift $k" >> gpurun_out/probe.log
  CVB_DBG_KSHIFT=$k timeout 100 python tools/train_profile.py v3 1 300 >> gpurun_out/probe.log 2>&1 | tail -3
done
cat gpurun_out/probe.log | grep -v "^  File" | tail -30
