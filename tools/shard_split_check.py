"""Under torchrun: forward + inverse of one S_n function with a given split level, wall time per
step vs summed kernel time (python -m torch.distributed.run ... tools/shard_split_check.py N SPLIT)"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import snfft_b200 as sb
from snfft_b200.sharded import ShardedPlan

n, split = int(sys.argv[1]), int(sys.argv[2])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sp = ShardedPlan(n, rank, world, device=local, split=split)
p2p = sp.enable_p2p()
b, e = sp.local_elements()
x = torch.randn(e - b, dtype=torch.float64, device=f"cuda:{local}")
for _ in range(3):
    back = sp.inverse(sp.forward(x))
sb.timing_enable(True)
times = []
for _ in range(6):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    back = sp.inverse(sp.forward(x))
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
kinds = sb.timing_read()
err = float((back - x).norm() / x.norm())
print(f"rank {rank} blocks {sp.block_range[rank]} m {sp.m} p2p {p2p} wall ms {[round(t, 2) for t in times]} "
      f"kernel ms/step {sum(v[0] for v in kinds.values()) / 6:.2f} err {err:.1e}", flush=True)
dist.destroy_process_group()
