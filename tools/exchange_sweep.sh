#!/bin/bash
# NVLink exchange kernel sweep on N GPUs of one box: bash tools/exchange_sweep.sh (under gpurun --gpus N)
cd ${GRAFT_REPO_ROOT:-.}
N=$(nvidia-smi -L | wc -l)
for cfg in "4 256" "8 1024" "8 8192" "4 40000"; do
  set -- $cfg
  SNFFT_XCHG_UNROLL=$1 SNFFT_XCHG_GX=$2 timeout 300 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload s12 --steps 5 --warmup 3 2>/dev/null |
    python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line)['coset_split']; print('unroll $1 gx $2: step %.2f ms exchange/dir %.3f ms %.0f GB/s barrier %.2f' % (d['ms_per_step'], d['exchange_ms_per_direction'], d['exchange_GBps_per_rank'], d['per_kernel_ms'].get('barrier',0)))
"
done
