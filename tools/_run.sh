cd $GRAFT_REPO_ROOT
L=gpurun_out/r2_misc9.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q 2>&1 | tail -3 >> $L
for w in "12:8,11" "12:8,10,11" "12:8,9,11" "12:7,9,11" "12:7,9,10,11" "12:8,10,11/11:8,9,10" "12:8,10,11/11:7,9,10/10:7,9"; do
SNFFT_WINDOWS="$w" timeout 200 python tools/level_step.py 12 1 5 >> $L 2>&1
done
cat $L
