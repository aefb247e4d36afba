"""Does running the S_8 x 1024 transform in batch chunks help (the intermediate of a chunk stays in
the 126 MB L2 between the fused kernel and the level-8 kernel)?  Times forward + inverse with the
batch cut into 1, 2, 4, 8, 16 pieces."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snfft_b200 as sb

plan = sb.get_plan(8)
B = 1024
x = torch.randn(B, plan.size, dtype=torch.float64, device='cuda')
spec, back, work = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)


def step(chunks):
    per = B // chunks
    for c in range(chunks):
        s = slice(c * per, (c + 1) * per)
        plan.forward(x[s], 'coset', out=spec[s], work=work[:per])
    for c in range(chunks):
        s = slice(c * per, (c + 1) * per)
        plan.inverse(spec[s], 'coset', out=back[s], work=work[:per])


for chunks in (1, 2, 4, 8, 16, 1):
    for _ in range(3):
        step(chunks)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step(chunks)
    e1.record()
    torch.cuda.synchronize()
    print('chunks', chunks, 'ms/step %.4f' % (e0.elapsed_time(e1) / 20), 'roundtrip %.1e' % float((back - x).norm() / x.norm()))
