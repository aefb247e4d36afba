"""Young's orthogonal representation with the reference's interface (reference ``yor.py``).

``yor_trans`` reads rho_lambda((k k+1)) off the plan's chain-indexed coefficient tables
(``snfft_yor_trans``); ``yor`` evaluates rho_lambda(sigma) on the device by applying the
bubble-sort word of sigma to the identity (``snfft_yor``).  There is no host implementation of
the matrix product: without a B200 ``yor`` raises.
"""
from __future__ import annotations

import numpy as np

from .plan import Plan, get_plan
from .young_tableau import FerrersDiagram

YOR_CACHE = {}     # (partition, tup_rep) -> matrix, filled when use_cache=True (yor.py:15)
YOR_T_CACHE = {}   # (partition, transposition) -> matrix (yor.py:16)
CACHE = {'hit': 0, 'sparse_hit': 0}

_HOST_PLANS = {}


def _host_plan(n):
    """Host-only plan: enough for the generator tables (no device needed)."""
    if n not in _HOST_PLANS:
        _HOST_PLANS[n] = Plan(n, -1)
    return _HOST_PLANS[n]


def cycle_to_adj_transpositions(cyc, n):
    """Word of adjacent transpositions whose product is the cycle (yor.py:19-38): the swaps of
    a bubble sort of the one-line form, last swap first."""
    cyc = tuple(cyc)
    line = [cyc[(cyc.index(x) + 1) % len(cyc)] if x in cyc else x for x in range(1, n + 1)]
    swaps = []
    for done in range(n):
        for pos in range(n - 1, done, -1):
            if line[pos] < line[pos - 1]:
                line[pos], line[pos - 1] = line[pos - 1], line[pos]
                swaps.append((pos, pos + 1))
    swaps.reverse()
    return swaps


def perm_to_adj_transpositions(perm, n):
    """Concatenated words of the cycles of a permutation (yor.py:40-50)."""
    word = []
    for cyc in perm:
        word.extend(cycle_to_adj_transpositions(cyc, n))
    return word


def _full_tuple(permutation, n):
    tup = tuple(permutation.tup_rep) if hasattr(permutation, 'tup_rep') else tuple(permutation)
    return tup + tuple(range(len(tup) + 1, n + 1))


def yor_trans(ferrers, transposition):
    """rho_lambda((k k+1)) as a dense d x d matrix (yor.py:87-117)."""
    a, b = transposition
    assert a < b
    if b != a + 1:
        raise ValueError('yor_trans needs an adjacent transposition (k, k+1)')
    key = (ferrers.partition, (a, b))
    if key in YOR_T_CACHE:
        CACHE['hit'] += 1
        return YOR_T_CACHE[key]
    n = ferrers.size
    plan = _host_plan(n)
    mat = plan.yor_trans(n, plan.irrep_index(ferrers.partition), a)
    YOR_T_CACHE[key] = mat
    return mat


def yor(ferrers, permutation, use_cache=True):
    """rho_lambda(sigma) (yor.py:52-85).  ``permutation`` is a Perm/Perm2 (anything with
    ``tup_rep``); shorter tuples are padded with fixed points."""
    n = ferrers.size
    tup = _full_tuple(permutation, n)
    key = (ferrers.partition, tup)
    if key in YOR_CACHE:
        CACHE['hit'] += 1
        return YOR_CACHE[key]
    mat = yor_batch(ferrers, [tup])[0]
    if use_cache:
        YOR_CACHE[key] = mat
    return mat


def yor_batch(ferrers, tuples, device=0):
    """rho_lambda(sigma) for a list of permutation tuples -> ndarray [count, d, d]
    (the batched form used by exp/yor_dataset.py:26-49 style consumers)."""
    import torch
    n = ferrers.size
    plan = get_plan(n, device)
    perms = torch.tensor([_full_tuple(t, n) for t in tuples], dtype=torch.int32,
                         device=f'cuda:{device}').reshape(-1, n)
    out = plan.yor(plan.irrep_index(ferrers.partition), perms)
    return out.cpu().numpy()


def ysemi_batch(ferrers, tuples, device=0):
    """Seminormal-form matrices of a list of permutation tuples on the device (``snfft_ysemi``)
    -> ndarray [count, d, d]; the host functions below are the reference-shaped entry points."""
    import torch
    n = ferrers.size
    plan = get_plan(n, device)
    perms = torch.tensor([_full_tuple(t, n) for t in tuples], dtype=torch.int32,
                         device=f'cuda:{device}').reshape(-1, n)
    return plan.yor(plan.irrep_index(ferrers.partition), perms, seminormal=True).cpu().numpy()


# ---- Young's seminormal form (reference yor.py:119-182); small host-side helper, not on the hot path
def ysemi_t(f, transposition):
    """Seminormal matrix of the adjacent transposition (k, k+1): for a tableau t whose swap t' is
    standard and comes later in the list, rows (t, t') carry [[1/d, 1 - 1/d^2], [1, -1/d]] with d the
    axial distance; otherwise +1 (same row) or -1 (same column)."""
    tabs = f.tableaux
    rep = np.zeros((len(tabs), len(tabs)))
    i0, i1 = transposition
    for t in tabs:
        other = t.transpose(transposition)
        if other is None:
            if t.get_row(i0) == t.get_row(i1):
                rep[t.idx, t.idx] = 1.
            elif t.get_col(i0) == t.get_col(i1):
                rep[t.idx, t.idx] = -1.
            continue
        if t.idx < other.idx:
            dist = t.ax_dist(i0, i1)
            rep[t.idx, t.idx] = 1. / dist
            rep[t.idx, other.idx] = 1. - 1. / dist ** 2
            rep[other.idx, t.idx] = 1.
            rep[other.idx, other.idx] = -1. / dist
    return rep


def ysemi(ferrers, permutation):
    """Seminormal-form matrix of a permutation: generators along the bubble-sort word of each cycle,
    accumulated from the left (yor.py:119-146)."""
    cycles = [c for c in permutation.cycle_decomposition if len(c) > 1]
    if not cycles:
        return np.eye(ferrers.n_tabs())
    res = None
    for cycle in cycles:
        for t in reversed(cycle_to_adj_transpositions(cycle, ferrers.size)):
            y = ysemi_t(ferrers, (min(t), max(t)))
            res = y if res is None else y.dot(res)
    return res
