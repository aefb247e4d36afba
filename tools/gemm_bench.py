"""fp64 DMMA GEMM (snfft_dgemm) at the shapes of the path: direct sums [B, n!] x [n!, n!] (n = 7) and a
square reference shape; prints TFLOP/s.  Run under ncu for the pipe utilisation (profiles/README.md)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snfft_b200 import _native as N


def gemm(a, b):
    m, k = a.shape
    n = b.shape[1]
    out = torch.empty((m, n), dtype=torch.float64, device=a.device)
    N.check(N.lib.snfft_dgemm(N.ptr(a), N.ptr(b), N.ptr(out), m, n, k,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


for (m, n, k) in [(1024, 5040, 5040), (4096, 4096, 4096), (1024, 720, 720), (313600, 16, 72)]:
    a = torch.randn(m, k, dtype=torch.float64, device='cuda')
    b = torch.randn(k, n, dtype=torch.float64, device='cuda')
    c = gemm(a, b)
    err = float((c - a @ b).abs().max() / (a @ b).abs().max())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gemm(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    e0.record()
    for _ in range(5):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    ms_cublas = e0.elapsed_time(e1) / 5
    print(f'[{m},{k}]x[{k},{n}] snfft_dgemm {ms:.3f} ms {2.0 * m * n * k / ms / 1e9:.2f} TFLOP/s | '
          f'cuBLAS {ms_cublas:.3f} ms {2.0 * m * n * k / ms_cublas / 1e9:.2f} TFLOP/s | max rel diff {err:.1e}')
