cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_misc4.log
: > $L
timeout 900 python -m pytest tests/test_gpu_wreath.py tests/test_gpu_sharded.py tests/test_gpu_api.py -q 2>&1 | tail -12 >> $L
cat $L
