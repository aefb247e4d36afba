"""A/B of the inverse fused kernel's stage-in: 2 520 16-byte cp.async (LDGSTS) vs ONE cp.async.bulk
(1-D TMA, UBLKCP) + mbarrier per S_7 spectrum.  S_8 x 1024, per-kind kernel times of the library."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snfft_b200 as sb
from snfft_b200 import _native as N

plan = sb.get_plan(8)
x = torch.randn(1024, plan.size, dtype=torch.float64, device='cuda')
spec, back, work = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
plan.forward(x, 'coset', out=spec, work=work)
ref = None
for rep in range(2):
    for tma in (0, 1):
        N.check(N.lib.snfft_debug_fused_tma(tma))
        for _ in range(3):
            plan.inverse(spec, 'coset', out=back, work=work)
        torch.cuda.synchronize()
        sb.timing_enable(True)
        sb.timing_read()
        for _ in range(20):
            plan.inverse(spec, 'coset', out=back, work=work)
        kinds = sb.timing_read()
        sb.timing_enable(False)
        if ref is None:
            ref = back.clone()
        same = bool(torch.equal(back, ref))
        print('tma' if tma else 'ldgsts', 'fused_inv ms/launch %.4f' % (kinds['fused_inv'][0] / kinds['fused_inv'][1]),
              'identical' if same else 'DIFFERENT', 'roundtrip %.2e' % float((back - x).norm() / x.norm()))
N.check(N.lib.snfft_debug_fused_tma(0))
