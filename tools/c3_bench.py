"""Stage 1 of the wreath transform alone (snfft_c3_dft forward / adjoint on preallocated planes) and
the whole C3 wr S8 forward / inverse at (2,3,3)/((1,1),(1,1,1),(2,1)): python tools/c3_bench.py [REPS].
SNFFT_B200_LIB selects the library (A/B of kernel variants)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snfft_b200 import _native as N
from snfft_b200.wreath import WreathPlan

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
wp = WreathPlan((2, 3, 3), ((1, 1), (1, 1, 1), (2, 1)), device=0)
f = torch.randn((40320, 2187), dtype=torch.float64, device='cuda')
wp.character_sums(f)                                     # builds wp.freq
Fre = torch.empty((40320, wp.N), dtype=torch.float64, device='cuda')
Fim = torch.empty_like(Fre)
gre, gim = torch.empty_like(f), torch.empty_like(f)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dft(inverse):
    if inverse:
        N.check(N.lib.snfft_c3_dft(wp.dft_m, N.ptr(Fre), N.ptr(Fim), N.ptr(wp.freq), wp.N, 40320, N.ptr(gre), N.ptr(gim), 1, stream))
    else:
        N.check(N.lib.snfft_c3_dft(wp.dft_m, N.ptr(f), None, N.ptr(wp.freq), wp.N, 40320, N.ptr(Fre), N.ptr(Fim), 0, stream))


def timed(fn):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


fh = wp.forward(f)
res = {'lib': os.path.basename(N.LIB_PATH),
       'c3_forward_ms': timed(lambda: dft(False)), 'c3_adjoint_ms': timed(lambda: dft(True)),
       'wreath_forward_ms': timed(lambda: wp.forward(f)), 'wreath_inverse_ms': timed(lambda: wp.inverse(fh))}
print({k: (tuple(round(x, 4) for x in v) if isinstance(v, tuple) else v) for k, v in res.items()})
