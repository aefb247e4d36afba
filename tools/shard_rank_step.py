"""The column-shard levels of ONE rank of a virtual world on one GPU (for ncu launch lists):
python tools/shard_rank_step.py N WORLD [RANK] [SPLIT]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snfft_b200.sharded import ShardedPlan

n, world = int(sys.argv[1]), int(sys.argv[2])
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
split = int(sys.argv[4]) if len(sys.argv) > 4 else None
sp = ShardedPlan(n, rank, world, device=0, split=split)
x = torch.randn(sp.nblocks * sp.sizes[rank][sp.m], dtype=torch.float64, device="cuda")
print("n", n, "world", world, "m", sp.m, "column-shard elements", x.numel(), "blocks", sp.block_range[rank])
# the local part of this rank: levels 2..m on its blocks
b, e = sp.block_range[rank]
f = torch.randn((e - b) * sp.mfact, dtype=torch.float64, device="cuda")
for rep in range(2):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    loc = sp.forward_local_spectra(f)
    bk = sp._local_inverse(loc)
    t1.record()
    torch.cuda.synchronize()
    print("levels 2..m forward + inverse ms", t0.elapsed_time(t1))
del loc, bk, f
for rep in range(2):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    spec = sp.forward_finish_columns(x.clone())
    back = sp.inverse_local_columns(spec)
    t1.record()
    torch.cuda.synchronize()
    print("levels m+1..n forward + inverse ms", t0.elapsed_time(t1), "rel", float((back - x).norm() / x.norm()))
