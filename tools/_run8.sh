cd $GRAFT_REPO_ROOT
bash tools/exchange_sweep.sh > gpurun_out/r2_xchg_sweep.log 2>&1
cat gpurun_out/r2_xchg_sweep.log
