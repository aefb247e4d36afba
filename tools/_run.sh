cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_misc7.log
: > $L
timeout 900 python -m pytest tests/test_gpu_wreath.py -q 2>&1 | tail -5 >> $L
cat $L
