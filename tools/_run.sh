cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_sweep4.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -5 >> $L
timeout 120 python tools/level_step.py 10 1 5 >> $L 2>&1
timeout 120 python tools/level_step.py 11 1 5 >> $L 2>&1
timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
SNFFT_WINDOWS="12:7,11/11:7,10/10:7,9" timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
SNFFT_WINDOWS="12:8,11/11:10/10:9" timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_l12_v4.csv python tools/level_step.py 12 1 1 > gpurun_out/ncu_l12.log 2>&1
cat $L
