"""Per-phase cycles of the fused S_7 kernel (debug build with -DSNFFT_PHASE_CLOCKS).

SNFFT_B200_LIB=$PWD/snfft_b200/_lib/variants/clocks.so python tools/phase_clocks.py
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snfft_b200 as sb
from snfft_b200 import _native as N

plan = sb.get_plan(8)
f = torch.randn(1024, plan.size, dtype=torch.float64, device="cuda")
buf = (ctypes.c_ulonglong * 32)()
N.lib.snfft_debug_phase_cycles.argtypes = [ctypes.c_void_p]
base = None
for rep in range(3):
    fhat = plan.forward(f, "coset")
    back = plan.inverse(fhat, "coset")
    torch.cuda.synchronize()
    N.lib.snfft_debug_phase_cycles(buf)
    cur = list(buf)
    if rep == 2:
        groups = 1024 * 8
        d = [(c - b) / groups for c, b in zip(cur, base)]
        names_f = {0: "stage-in", 1: "S4", 2: "L5", 3: "L6", 4: "L7 phase 1", 12: "L7 phase 2 + end barrier"}
        names_i = {5: "stage-in", 6: "L7 phase 2^T", 7: "L7 phase 1^T", 8: "L6^T", 9: "L5^T", 11: "S4^T", 12: "copy-out"}
        print("forward, cycles per S_7 group:")
        for k, v in names_f.items():
            print(f"  {v:28s} {d[k]:9.0f}")
        print("  total", sum(d[k] for k in names_f))
        print("inverse, cycles per S_7 group:")
        for k, v in names_i.items():
            print(f"  {v:28s} {d[16 + k]:9.0f}")
        print("  total", sum(d[16 + k] for k in names_i))
    base = cur
