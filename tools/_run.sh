cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_misc10.log
: > $L
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $L
timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
SNFFT_PHASE_A=2 timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
timeout 120 python tools/level_step.py 11 1 5 >> $L 2>&1
timeout 120 python tools/level_step.py 10 8 5 >> $L 2>&1
timeout 120 python tools/level_step.py 9 64 5 >> $L 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_l12_v7.csv python tools/level_step.py 12 1 1 > gpurun_out/ncu_l12.log 2>&1
cat $L
