cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_final1.log
: > $L
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $L
python -c "import __graft_entry__ as g; g.build(); g.smoke()" >> $L 2>&1
timeout 600 python bench.py > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err
echo "bench rc=$?" >> $L
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_1gpu_ref.json 2>> gpurun_out/r2_final_bench_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-coset-split --no-wreath --sustain-seconds 0 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:phase_kernel -c 9 -o gpurun_out/r2_phase_v6 python tools/level_step.py 12 1 1 > gpurun_out/ncu_full.log 2>&1
cat $L
