"""One forward + inverse transform at a large n (for ncu launch lists): python tools/large_n_step.py N BATCH"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snfft_b200 as sb

n, batch = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
plan = sb.get_plan(n)
f = torch.randn(batch, plan.size, dtype=torch.float64, device="cuda")
for _ in range(reps):
    fhat = plan.forward(f)
    back = plan.inverse(fhat)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
fhat = plan.forward(f)
back = plan.inverse(fhat)
t1.record()
torch.cuda.synchronize()
print("n", n, "batch", batch, "ms", t0.elapsed_time(t1), "rel", float((back - f).norm() / f.norm()))
