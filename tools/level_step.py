"""Forward + inverse of one S_n function in coset-major order, timed with CUDA events, plus the
per-kind kernel times of the library: python tools/level_step.py N [BATCH] [REPS].
Run under `SNFFT_WINDOWS=...` to try other phase windows (see plan.hpp: phase_ranges)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snfft_b200 as sb

n = int(sys.argv[1])
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
plan = sb.get_plan(n)
f = torch.randn(batch, plan.size, dtype=torch.float64, device="cuda")
spec, back, work = torch.empty_like(f), torch.empty_like(f), torch.empty_like(f)
for _ in range(2):
    plan.forward(f, 'coset', out=spec, work=work)
    plan.inverse(spec, 'coset', out=back, work=work)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
fw, iv = [], []
for _ in range(reps):
    ev[0].record()
    plan.forward(f, 'coset', out=spec, work=work)
    ev[1].record()
    plan.inverse(spec, 'coset', out=back, work=work)
    ev[2].record()
    torch.cuda.synchronize()
    fw.append(ev[0].elapsed_time(ev[1]))
    iv.append(ev[1].elapsed_time(ev[2]))
sb.timing_enable(True)
sb.timing_read()
plan.forward(f, 'coset', out=spec, work=work)
plan.inverse(spec, 'coset', out=back, work=work)
kinds = {k: round(v[0], 3) for k, v in sb.timing_read().items() if v[1]}
sb.timing_enable(False)
rel = float((back - f).norm() / f.norm())
pl = float(((spec.double() ** 2).sum()))
print("n", n, "batch", batch, "windows", os.environ.get("SNFFT_WINDOWS", "default"),
      "kb", os.environ.get("SNFFT_BUNDLE_KB", "default"),
      "fwd_ms %.3f inv_ms %.3f total %.3f" % (min(fw), min(iv), min(fw) + min(iv)), "roundtrip %.2e" % rel, kinds)
