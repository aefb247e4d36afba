cd $GRAFT_REPO_ROOT
N=$(nvidia-smi -L | wc -l)
L=gpurun_out/r2_final_${N}gpu.log
: > $L
export PYTHONFAULTHANDLER=1
timeout 900 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload s12 --steps 5 --warmup 3 > gpurun_out/r2_final_bench_${N}gpu_s12.json 2> gpurun_out/r2_final_bench_${N}gpu.err
echo "rc=$?" >> $L
cat $L
