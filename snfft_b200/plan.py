"""Python handle on a native plan (``snfft_plan``) and the tensor-level transform API.

This is the additive "batched tensor" entry point of SURVEY.md section 8(b); the reference-shaped
functions in ``snfft_b200/fft.py`` are thin adaptors over it.  PyTorch is used for device memory
and streams only; all arithmetic happens in the CUDA library behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
import math
import threading

import numpy as np

from . import _native as N

_PLANS = {}
_LOCK = threading.Lock()


class Plan:
    """Tables of the S_n transform, resident on one device (``device=-1``: host tables only)."""

    def __init__(self, n: int, device: int = 0):
        handle = C.c_void_p()
        N.check(N.lib.snfft_plan_create(int(n), int(device), C.byref(handle)))
        self._h = handle
        self.n = int(n)
        self.device = int(device)
        self.size = math.factorial(self.n)
        self._layouts = {}

    def close(self):
        if getattr(self, "_h", None):
            N.lib.snfft_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host-side tables
    def layout(self, k: int | None = None):
        """(partitions, dims, offsets) of S_k in ``utils.partitions(k)`` order."""
        k = self.n if k is None else int(k)
        if k not in self._layouts:
            cnt = N.lib.snfft_num_irreps(self._h, k)
            if cnt < 0:
                raise N.SnfftError(f"level {k} is outside 1..{self.n}")
            parts = np.zeros((cnt, k), dtype=np.int32)
            dims = np.zeros(cnt, dtype=np.int32)
            offs = np.zeros(cnt, dtype=np.int64)
            N.check(N.lib.snfft_layout(self._h, k, N.ptr(parts), N.ptr(dims), N.ptr(offs)))
            plist = [tuple(int(v) for v in row if v > 0) for row in parts]
            self._layouts[k] = (plist, [int(d) for d in dims], [int(o) for o in offs])
        return self._layouts[k]

    def irrep_index(self, partition, k: int | None = None) -> int:
        k = sum(partition) if k is None else k
        return self.layout(k)[0].index(tuple(partition))

    def branch_down(self, k: int, irrep: int):
        idx = np.zeros(k + 1, dtype=np.int32)
        off = np.zeros(k + 1, dtype=np.int32)
        cnt = C.c_int32()
        N.check(N.lib.snfft_branch_down(self._h, k, irrep, N.ptr(idx), N.ptr(off), C.byref(cnt)))
        return [int(i) for i in idx[:cnt.value]], [int(o) for o in off[:cnt.value]]

    def yor_trans(self, k: int, irrep: int, j: int) -> np.ndarray:
        d = self.layout(k)[1][irrep]
        out = np.empty((d, d), dtype=np.float64)
        N.check(N.lib.snfft_yor_trans(self._h, k, irrep, j, N.ptr(out)))
        return out

    def coset_to_lex(self) -> np.ndarray:
        out = np.empty(self.size, dtype=np.int64)
        N.check(N.lib.snfft_coset_to_lex_host(self._h, N.ptr(out)))
        return out

    def macs_per_element(self, k: int) -> float:
        return float(N.lib.snfft_level_macs_per_element(self._h, k))

    def device_bytes(self) -> int:
        return int(N.lib.snfft_plan_device_bytes(self._h))

    def export_panel(self, k: int, panel: int):
        """Host copy of one panel program (see ``snfft_export_panel``)."""
        nb = C.c_int32()
        N.check(N.lib.snfft_export_panel(self._h, k, panel, C.byref(nb), None, None, None, None,
                                         None, None, None, None))
        nb = nb.value
        d = nb // (k - 1)
        m = np.zeros(nb, dtype=np.int32)
        stage = np.zeros(nb, dtype=np.int32)
        slots = np.zeros((nb, 5), dtype=np.int32)
        final = np.zeros(nb, dtype=np.int32)
        coef_f = np.zeros((nb, 25), dtype=np.float64)
        coef_i = np.zeros((nb, 25), dtype=np.float64)
        dest = np.zeros(k * d, dtype=np.int64)
        rowbase = np.zeros(k * d, dtype=np.int64)
        N.check(N.lib.snfft_export_panel(self._h, k, panel, C.byref(C.c_int32()), N.ptr(m),
                                         N.ptr(stage), N.ptr(slots), N.ptr(final), N.ptr(coef_f),
                                         N.ptr(coef_i), N.ptr(dest), N.ptr(rowbase)))
        return dict(d=d, m=m, stage=stage, slots=slots, final=final, coef_fwd=coef_f,
                    coef_inv=coef_i, dest=dest, rowbase=rowbase)

    def export_coef(self, k: int) -> np.ndarray:
        """Coefficient pool of level k (the matrices that block and phase tables index)."""
        cnt = C.c_int64()
        N.check(N.lib.snfft_export_coef(self._h, k, None, C.byref(cnt)))
        coef = np.zeros(cnt.value, dtype=np.float64)
        N.check(N.lib.snfft_export_coef(self._h, k, N.ptr(coef), C.byref(cnt)))
        return coef

    def export_phase(self, k: int, phase: int):
        """Host copy of the device tables of one row-group phase of level k >= 9 (see
        ``snfft_export_phase``): what ``phase_kernel`` executes, for CPU tests."""
        sizes = np.zeros(8, dtype=np.int32)
        N.check(N.lib.snfft_export_phase(self._h, k, phase, N.ptr(sizes), None, None, None, None))
        npan, ngrp, words = (int(v) for v in sizes[:3])
        panels = np.zeros((npan, 5), dtype=np.int32)
        groups = np.zeros((ngrp, 4), dtype=np.int64)
        blob_f = np.zeros(words * 16, dtype=np.uint8)
        blob_i = np.zeros(words * 16, dtype=np.uint8)
        N.check(N.lib.snfft_export_phase(self._h, k, phase, N.ptr(sizes), N.ptr(panels), N.ptr(groups),
                                         N.ptr(blob_f), N.ptr(blob_i)))
        return dict(panels=panels, groups=groups, blob_fwd=blob_f, blob_inv=blob_i, max_rows=int(sizes[3]),
                    max_blob=int(sizes[4]), nstages=int(sizes[5]), stage_lo=int(sizes[6]), nphases=int(sizes[7]))

    # ------------------------------------------------------------------ device transforms
    def _check(self, t, name, last=None):
        import torch
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float64:
            raise TypeError(f"{name} must be a float64 CUDA tensor")
        if t.device.index != self.device:
            raise ValueError(f"{name} lives on cuda:{t.device.index}, the plan on cuda:{self.device}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        if last is not None and (t.dim() == 0 or t.shape[-1] != last):
            raise ValueError(f"{name} must have last dimension {last}")

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def forward(self, f, order: str = "lex", out=None, work=None):
        """Packed spectrum [..., n!] of f [..., n!] (float64 CUDA tensor)."""
        import torch
        self._check(f, "f", self.size)
        out = torch.empty_like(f) if out is None else out
        work = torch.empty_like(f) if work is None else work
        self._check(out, "out", self.size)
        self._check(work, "work", self.size)
        batch = f.numel() // self.size
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_forward(self._h, N.ptr(f), N.ptr(out), N.ptr(work), batch,
                                        _order_flag(order), self._stream()))
        return out

    def inverse(self, fhat, order: str = "lex", out=None, work=None):
        """f [..., n!] from a packed spectrum [..., n!]."""
        import torch
        self._check(fhat, "fhat", self.size)
        out = torch.empty_like(fhat) if out is None else out
        work = torch.empty_like(fhat) if work is None else work
        self._check(out, "out", self.size)
        self._check(work, "work", self.size)
        batch = fhat.numel() // self.size
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_inverse(self._h, N.ptr(fhat), N.ptr(out), N.ptr(work), batch,
                                        _order_flag(order), self._stream()))
        return out

    def level(self, k: int, x, inverse: bool = False):
        """One level of the recursion: [groups, k, (k-1)!] -> [groups, k!] (or back)."""
        import torch
        self._check(x, "x")
        kf = math.factorial(k)
        if x.numel() % kf:
            raise ValueError("x must hold whole groups of k! elements")
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_level(self._h, k, N.ptr(x), N.ptr(out), x.numel() // kf,
                                      int(bool(inverse)), self._stream()))
        return out

    def reorder(self, x, to_coset: bool):
        import torch
        self._check(x, "x", self.size)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_reorder(self._h, N.ptr(x), N.ptr(out), x.numel() // self.size,
                                        int(bool(to_coset)), self._stream()))
        return out

    def yor(self, irrep: int, perms, seminormal: bool = False):
        """rho_lambda(sigma) for int32 CUDA tensor perms [count, n] -> [count, d, d] in Young's
        orthogonal form, or (``seminormal=True``) in the seminormal form of yor.py:119-182."""
        import torch
        if perms.dtype != torch.int32 or not perms.is_cuda or not perms.is_contiguous():
            raise TypeError("perms must be a contiguous int32 CUDA tensor")
        if perms.dim() != 2 or perms.shape[1] != self.n:
            raise ValueError(f"perms must have shape [count, {self.n}]")
        d = self.layout()[1][irrep]
        out = torch.empty((perms.shape[0], d, d), dtype=torch.float64, device=perms.device)
        with torch.cuda.device(self.device):
            fn = N.lib.snfft_ysemi if seminormal else N.lib.snfft_yor
            N.check(fn(self._h, irrep, N.ptr(perms), perms.shape[0], N.ptr(out), self._stream()))
        return out

    def inverse_at(self, fhat, perms):
        """f(sigma_j) for int32 CUDA tensor perms [count, n] from ONE packed spectrum fhat [n!]
        (``snfft_inverse_at``: all irreps in one launch)."""
        import torch
        self._check(fhat, "fhat", self.size)
        if fhat.numel() != self.size:
            raise ValueError("inverse_at takes a single packed spectrum [n!]")
        if perms.dtype != torch.int32 or not perms.is_cuda or not perms.is_contiguous():
            raise TypeError("perms must be a contiguous int32 CUDA tensor")
        if perms.dim() != 2 or perms.shape[1] != self.n:
            raise ValueError(f"perms must have shape [count, {self.n}]")
        out = torch.empty(perms.shape[0], dtype=torch.float64, device=perms.device)
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_inverse_at(self._h, N.ptr(fhat), N.ptr(perms), perms.shape[0], N.ptr(out),
                                           self._stream()))
        return out

    def dft_direct(self, f):
        """Direct-sum transform as one DMMA GEMM (n <= 8); f in lexicographic order.  The [n!, n!]
        matrix of all rho_lambda(sigma) is built on first use and stays in the plan (13 GB at
        n = 8) until ``dft_release()``."""
        import torch
        self._check(f, "f", self.size)
        out = torch.empty_like(f)
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_dft_direct(self._h, N.ptr(f), N.ptr(out), f.numel() // self.size,
                                           self._stream()))
        return out

    def dft_release(self):
        """Free the direct-sum matrix that ``dft_direct`` cached in this plan."""
        N.check(N.lib.snfft_dft_release(self._h))

    def roundtrip_host(self, f_host, fhat_host=None, back_host=None, order: str = "coset",
                       chunks: int = 8, streams: int = 4):
        """Forward (and, if ``back_host`` is given, inverse) transform of a batch that lives in
        pinned host memory: the batch is cut into ``chunks`` pieces that flow through
        H2D copy -> forward -> D2H copy [-> inverse -> D2H copy] on ``streams`` CUDA streams, so that
        the PCIe transfers of one piece overlap the kernels and the opposite-direction copies of the
        others.  f_host / fhat_host / back_host: float64 pinned CPU tensors [batch, n!]."""
        import torch
        batch = f_host.shape[0]
        if f_host.dtype != torch.float64 or f_host.dim() != 2 or f_host.shape[1] != self.size:
            raise TypeError("f_host must be a float64 CPU tensor [batch, n!]")
        chunks = max(1, min(chunks, batch))
        per = (batch + chunks - 1) // chunks
        dev = f"cuda:{self.device}"
        key = ("host_pipe", per, streams)
        if getattr(self, "_pipe_key", None) != key:
            self._pipe_key = key
            self._pipe_streams = [torch.cuda.Stream(device=dev) for _ in range(streams)]
            self._pipe_bufs = [[torch.empty((per, self.size), dtype=torch.float64, device=dev) for _ in range(4)]
                               for _ in range(streams)]
        start = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(start)
        for c in range(chunks):
            lo, hi = c * per, min(batch, (c + 1) * per)
            if lo >= hi:
                break
            st = self._pipe_streams[c % streams]
            x, spec, back, work = (b[:hi - lo] for b in self._pipe_bufs[c % streams])
            with torch.cuda.stream(st):
                st.wait_event(ready)
                x.copy_(f_host[lo:hi], non_blocking=True)
                self.forward(x, order, out=spec, work=work)
                if fhat_host is not None:
                    fhat_host[lo:hi].copy_(spec, non_blocking=True)
                if back_host is not None:
                    self.inverse(spec, order, out=back, work=work)
                    back_host[lo:hi].copy_(back, non_blocking=True)
        for st in self._pipe_streams:
            start.wait_stream(st)

    def unpack(self, packed):
        """{partition: view [..., d, d]} of a packed spectrum, in ``partitions(n)`` order."""
        parts, dims, offs = self.layout()
        lead = tuple(packed.shape[:-1])
        return {p: packed[..., o:o + d * d].reshape(lead + (d, d))
                for p, d, o in zip(parts, dims, offs)}


def _order_flag(order: str) -> int:
    if order == "lex":
        return N.ORDER_LEX
    if order == "coset":
        return N.ORDER_COSET
    raise ValueError("order must be 'lex' or 'coset'")


def get_plan(n: int, device: int = 0) -> Plan:
    """Process-wide cache of plans (one per (n, device))."""
    key = (int(n), int(device))
    with _LOCK:
        if key not in _PLANS:
            _PLANS[key] = Plan(n, device)
        return _PLANS[key]


def launch_count() -> int:
    return int(N.lib.snfft_launch_count())


KERNEL_KINDS = ("fused_fwd", "fused_inv", "level8_fwd", "level8_inv", "level_table", "p1", "p0", "phase_fwd",
                "phase_inv", "exchange", "barrier", "reorder", "yor", "dgemm", "other")


def timing_enable(on: bool) -> None:
    N.check(N.lib.snfft_timing_enable(int(bool(on))))


def timing_read():
    """{kind: (total ms, launches)} since the last read (synchronises the recorded events)."""
    ms = np.zeros(len(KERNEL_KINDS), dtype=np.float64)
    cnt = np.zeros(len(KERNEL_KINDS), dtype=np.int64)
    N.check(N.lib.snfft_timing_read(N.ptr(ms), N.ptr(cnt)))
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_KINDS)}
