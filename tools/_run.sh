cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_misc5.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q 2>&1 | tail -5 >> $L
timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
timeout 120 python tools/level_step.py 11 1 5 >> $L 2>&1
timeout 120 python tools/level_step.py 10 8 5 >> $L 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_l12_v54.csv python tools/level_step.py 12 1 1 > gpurun_out/ncu_l12.log 2>&1
cat $L
