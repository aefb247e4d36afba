cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_misc3.log
: > $L
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 >> $L
timeout 200 python tools/chunk_sweep.py >> $L 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench rc=$?" >> $L
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_b.err
echo "ref rc=$?" >> $L
cat $L
