"""The column-shard levels of ONE rank of a virtual world on one GPU (for ncu launch lists):
python tools/shard_rank_step.py N WORLD [RANK]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snfft_b200.sharded import ShardedPlan

n, world = int(sys.argv[1]), int(sys.argv[2])
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sp = ShardedPlan(n, rank, world, device=0)
x = torch.randn(sp.nblocks * sp.sizes[rank][sp.m], dtype=torch.float64, device="cuda")
print("n", n, "world", world, "m", sp.m, "column-shard elements", x.numel())
for rep in range(2):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    spec = sp.forward_finish_columns(x.clone())
    back = sp.inverse_local_columns(spec)
    t1.record()
    torch.cuda.synchronize()
    print("levels m+1..n forward + inverse ms", t0.elapsed_time(t1), "rel", float((back - x).norm() / x.norm()))
