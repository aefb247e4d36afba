"""Fourier transforms over S_n with the reference's entry points (reference ``fft.py``).

Signatures and return values are the reference's: ``f`` is a Python callable on permutation
objects, results are fresh ``float64[d, d]`` numpy arrays in the reference's tableau order.  The
callable is evaluated once per group element on the host; everything after that runs on the
B200 through the C ABI (``snfft_forward`` / ``snfft_dft_direct``).  No CPU fallback exists:
without the CUDA library or a device these functions raise.

Additive entry points (not in the reference): ``forward`` / ``inverse`` on dense arrays or
tensors, and ``ift_full`` -- the reference only has a per-element direct inverse
(tile/par_inv_ft.py:27-45).
"""
from __future__ import annotations

import hashlib
import math

import numpy as np

from . import perm as _perm
from . import perm2 as _perm2
from .plan import get_plan
from .utils import partitions
from .young_tableau import FerrersDiagram

_DEVICE = 0
_MEMO = {}          # (kind, n, digest of the sampled values) -> packed spectrum (host)
_MEMO_MAX = 8
# The direct sum is a GEMM with the [n!, n!] matrix of all rho_lambda(sigma): 203 MB at n = 7, 13 GB
# at n = 8 (supported by Plan.dft_direct, released with Plan.dft_release), 1 TB at n = 9.  The
# reference-shaped entry points keep the cached matrix small: n = 8 goes through the fast transform.
DIRECT_MAX_N = 7


def set_device(device: int) -> None:
    global _DEVICE
    _DEVICE = int(device)


def _sample(f, n, cls):
    """f evaluated on all of S_n in lexicographic order (the order of perm2.sn / perm.sn)."""
    group = _perm.sn(n) if cls is _perm.Perm else _perm2.sn(n)
    return np.array([f(p) for p in group], dtype=np.float64)


def _spectrum(values, n, kind):
    """Packed spectrum of a sampled function, memoised on the sampled values so that a loop over
    irreps (test.py:76-85 style) runs one device transform."""
    import torch
    key = (kind, n, hashlib.sha1(values.tobytes()).hexdigest())
    hit = _MEMO.get(key)
    if hit is not None:
        return hit
    plan = get_plan(n, _DEVICE)
    dev = torch.from_numpy(values).to(f'cuda:{_DEVICE}')
    if kind == 'direct' and n <= DIRECT_MAX_N:
        packed = plan.dft_direct(dev)
    else:
        packed = plan.forward(dev, order='lex')
    host = packed.cpu().numpy()
    if len(_MEMO) >= _MEMO_MAX:
        _MEMO.pop(next(iter(_MEMO)))
    _MEMO[key] = host
    return host


def _block(packed, n, partition):
    parts, dims, offs = get_plan(n, _DEVICE).layout()
    i = parts.index(tuple(partition))
    d, o = dims[i], offs[i]
    return packed[o:o + d * d].reshape(d, d).copy()


def fft(f, ferrers):
    """Clausen--Baum transform of ``f`` at the irrep ``ferrers`` (fft.py:72-113).
    ``f`` is called on ``perm.Perm`` objects, as in the reference."""
    n = ferrers.size
    return _block(_spectrum(_sample(f, n, _perm.Perm), n, 'fft'), n, ferrers.partition)


def fft2(f, ferrers):
    """Same transform with ``f`` called on ``perm2.Perm2`` objects (fft.py:39-70).  The
    reference's version raises for n >= 2 (perm2.py:141-142); this one returns what ``fft``
    returns."""
    n = ferrers.size
    return _block(_spectrum(_sample(f, n, _perm2.Perm2), n, 'fft'), n, ferrers.partition)


def fourier_transform(f, ferrers):
    """Direct sum over ``perm.sn`` (fft.py:115-132), as one DMMA GEMM on the device."""
    n = ferrers.size
    return _block(_spectrum(_sample(f, n, _perm.Perm), n, 'direct'), n, ferrers.partition)


def fourier_transform2(f, ferrers):
    """Direct sum over ``perm2.sn`` (fft.py:134-151), as one DMMA GEMM on the device."""
    n = ferrers.size
    return _block(_spectrum(_sample(f, n, _perm2.Perm2), n, 'direct'), n, ferrers.partition)


def ft_full(f, n, algo='fft'):
    """All irreps at once (fft.py:16-37): dict keyed by fresh FerrersDiagram objects in
    ``partitions(n)`` order.  algo: 'fft' | 'ft1' | 'ft2' as in the reference."""
    if algo == 'fft':
        packed = _spectrum(_sample(f, n, _perm2.Perm2), n, 'fft')
    elif algo == 'ft1':
        packed = _spectrum(_sample(f, n, _perm.Perm), n, 'direct')
    elif algo == 'ft2':
        packed = _spectrum(_sample(f, n, _perm2.Perm2), n, 'direct')
    else:
        return {}
    return {FerrersDiagram(p): _block(packed, n, p) for p in partitions(n)}


# ---- additive dense API -----------------------------------------------------------------------
def forward(values, n=None, order='lex'):
    """{partition: ndarray[..., d, d]} of a dense array of samples [..., n!]."""
    import torch
    values = np.ascontiguousarray(values, dtype=np.float64)
    n = _degree(values.shape[-1]) if n is None else n
    plan = get_plan(n, _DEVICE)
    packed = plan.forward(torch.from_numpy(values).to(f'cuda:{_DEVICE}'), order=order)
    return {p: v.cpu().numpy() for p, v in plan.unpack(packed).items()}


def inverse(fhat, n, order='lex'):
    """Samples [..., n!] from {partition: [..., d, d]} (any mapping keyed by partition tuples or
    FerrersDiagram objects)."""
    import torch
    plan = get_plan(n, _DEVICE)
    parts, dims, offs = plan.layout()
    by_part = {(k.partition if hasattr(k, 'partition') else tuple(k)): np.asarray(v, dtype=np.float64)
               for k, v in fhat.items()}
    lead = by_part[parts[0]].shape[:-2]
    packed = np.empty(lead + (plan.size,), dtype=np.float64)
    for p, d, o in zip(parts, dims, offs):
        packed[..., o:o + d * d] = by_part[p].reshape(lead + (d * d,))
    out = plan.inverse(torch.from_numpy(packed).to(f'cuda:{_DEVICE}'), order=order)
    return out.cpu().numpy()


def ift_full(fhat, n):
    """Inverse of ``ft_full``: the sampled function on S_n in ``sn(n)`` order."""
    return inverse(fhat, n, order='lex')


def inverse_at(fhat, n, perms):
    """f(sigma_j) for a list of permutation tuples -- the sampled inverse that the reference's
    consumers run per state (tile/par_inv_ft.py:27-45, compute_policy.py:66-88):
    f(sigma) = (1/n!) sum_lambda d_lambda sum(f_hat(lambda) o rho_lambda(sigma)).
    One device launch for all irreps and all states (``snfft_inverse_at``)."""
    import torch
    plan = get_plan(n, _DEVICE)
    parts, dims, offs = plan.layout()
    by_part = {(k.partition if hasattr(k, 'partition') else tuple(k)): v for k, v in fhat.items()}
    dev = f'cuda:{_DEVICE}'
    packed = np.empty(plan.size, dtype=np.float64)
    for p, d, o in zip(parts, dims, offs):
        packed[o:o + d * d] = np.asarray(by_part[p], dtype=np.float64).reshape(d * d)
    tuples = [tuple(p.tup_rep) if hasattr(p, 'tup_rep') else tuple(p) for p in perms]
    pt = torch.tensor([t + tuple(range(len(t) + 1, n + 1)) for t in tuples], dtype=torch.int32,
                      device=dev).reshape(-1, n)
    return plan.inverse_at(torch.from_numpy(packed).to(dev), pt).cpu().numpy()


def cube2_inv_fft_part(irrep_dict, fourier_mat, coset_reps, cyclic_irrep_func, cyc_tup, perm_tup):
    """One term of the wreath inverse at g = (cyc_tup, perm_tup) in C3 wr S_n (fft.py:158-172):
    dim * trace(rho(g^-1) f_hat), with rho(g^-1) assembled from the block dictionary of
    ``wreath.wreath_yor`` exactly as the reference's callers hold it (test_inv_fft.py:135).  The
    caller sums over irreps and divides by the group order.  All elements at once on the device:
    ``wreath.WreathPlan.inverse``."""
    from .wreath import WreathCycSn, wreath_rep
    cyc_inv, perm_inv = WreathCycSn.from_tup(cyc_tup, perm_tup, order=3).inv_tup_rep()
    rho_inv = wreath_rep(cyc_inv, perm_inv, irrep_dict, coset_reps, cyclic_irrep_func)
    return fourier_mat.shape[0] * np.sum(rho_inv.T * fourier_mat)


def cube2_inv_fft_func(irrep_dict, fourier_mat, coset_reps, cyclic_irrep_func):
    """(cyc_tup, perm_tup) -> ``cube2_inv_fft_part`` with the irrep fixed (fft.py:153-156)."""
    def evaluate(cyc_tup, perm_tup):
        return cube2_inv_fft_part(irrep_dict, fourier_mat, coset_reps, cyclic_irrep_func, cyc_tup, perm_tup)
    return evaluate


def _degree(size):
    n = 1
    while math.factorial(n) < size:
        n += 1
    if math.factorial(n) != size:
        raise ValueError('last dimension must be n!')
    return n
