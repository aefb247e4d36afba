cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_sweep8.log
: > $L
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 >> $L
SNFFT_PHASE_A=1 timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
SNFFT_PHASE_A=2 timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
timeout 120 python tools/level_step.py 11 1 5 >> $L 2>&1
timeout 120 python tools/level_step.py 10 8 5 >> $L 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -5 gpurun_out/r2_bench_a.err >> $L
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_l12_v52.csv python tools/level_step.py 12 1 1 > gpurun_out/ncu_l12.log 2>&1
cat $L
