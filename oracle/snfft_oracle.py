"""CPU oracle for the S_n Fourier-transform hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's (horacepan/SnFFT) algorithm for the path
named in BASELINE.json: Young's orthogonal representation, the direct Fourier sums, the
Clausen--Baum recursion and the direct inverse.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``snfft_b200/`` imports it; the product path is CUDA only.

Parity pin: every function here is checked in ``tests/test_oracle.py`` against golden vectors
produced by importing the UNMODIFIED reference from ``/root/reference`` in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) and against the known-answer
matrices in the reference's own ``test.py:53-85``.  For n >= 9 the reference itself cannot run
(``young_tableau.py:109`` scans (n-1)! fillings), so there the oracle is "parity unpinned" and
stands on the chain of trust  reference (n<=8) -> this file (same code, any n) -> CUDA.

Each function cites the reference file:line it follows.  Conventions:
permutations are 1-based tuples ``t`` with ``t[i-1] = sigma(i)``; ``(g*h)(i) = g(h(i))``;
``sn(n)`` enumerates in ``itertools.permutations`` (lexicographic) order; functions on S_n are
float64 vectors indexed by that lexicographic rank.
"""
from __future__ import annotations

import itertools
import math
from functools import lru_cache

import numpy as np


# --------------------------------------------------------------------------------------------
# combinatorics
# --------------------------------------------------------------------------------------------
def partitions(n, start=1):
    """All partitions of n in the reference's irrep order (utils.py:52-67)."""
    if n == 0:
        return [()]
    parts = [(n,)]
    for i in range(start, n // 2 + 1):
        for p in partitions(n - i, i):
            parts.append(p + (i,))
    return parts


def branch_down(partition):
    """Shapes obtained by removing one corner, rows top->bottom (young_tableau.py:62-81)."""
    if sum(partition) == 1:
        return []
    out = []
    for idx in range(len(partition)):
        nxt = partition[idx + 1] if idx != len(partition) - 1 else 0
        if partition[idx] - 1 >= nxt:
            # young_tableau.py:22-39 get_minus_partition
            out.append(tuple(v - 1 if j == idx else v
                             for j, v in enumerate(partition) if not (j == idx and v == 1)))
    return out


@lru_cache(maxsize=None)
def tableaux(partition):
    """Standard tableaux of a shape in the reference's sorted (last-letter) order.

    The reference brute-forces all fillings and sorts with ``YoungTableau.__lt__``
    (young_tableau.py:98-125, :193-206).  Last-letter order means: compare the row of n first,
    then n-1, ... -- i.e. the list is the concatenation, over ``branch_down`` order, of the
    tableaux of each sub-shape with n appended in the removed corner.  Returned
    as tuples of row tuples (the reference's ``YoungTableau.contents``).
    """
    n = sum(partition)
    if n == 1:
        return (((1,),),)
    out = []
    for idx in range(len(partition)):
        nxt = partition[idx + 1] if idx != len(partition) - 1 else 0
        if partition[idx] - 1 < nxt:
            continue
        sub = tuple(v - 1 if j == idx else v
                    for j, v in enumerate(partition) if not (j == idx and v == 1))
        for t in tableaux(sub):
            rows = [list(r) for r in t]
            if idx < len(rows):
                rows[idx].append(n)
            else:
                rows.append([n])
            out.append(tuple(tuple(r) for r in rows))
    return tuple(out)


def _content_maps(tab):
    """content(x) = col - row for every letter (young_tableau.py:329-339)."""
    c = {}
    for r, row in enumerate(tab):
        for k, x in enumerate(row):
            c[x] = k - r
    return c


@lru_cache(maxsize=None)
def yor_trans(partition, k):
    """rho_lambda((k k+1)) in Young's orthogonal form (yor.py:87-117).

    rep[i,i] = 1/dist with dist = content(k+1) - content(k) (young_tableau.py:326-327); when
    swapping k,k+1 gives a standard tableau j: rep[i,j] = rep[j,i] = sqrt(1 - 1/dist^2).
    """
    tabs = tableaux(partition)
    index = {t: i for i, t in enumerate(tabs)}
    rep = np.zeros((len(tabs), len(tabs)))
    for i, t in enumerate(tabs):
        c = _content_maps(t)
        dist = c[k + 1] - c[k]
        rep[i, i] = 1.0 / dist
        swapped = tuple(tuple(k + 1 if x == k else (k if x == k + 1 else x) for x in row)
                        for row in t)
        j = index.get(swapped)          # young_tableau.py:305-312 (None if not standard)
        if j is not None:
            rep[i, j] = np.sqrt(1 - (1.0 / dist) ** 2)
            rep[j, i] = rep[i, j]
    return rep


def cycle_to_adj_transpositions(cyc, n):
    """Bubble-sort word of a cycle (yor.py:19-38); returns [(j, j+1), ...]."""
    cyc = tuple(cyc)
    perm = [cyc[(cyc.index(x) + 1) % len(cyc)] if x in cyc else x for x in range(1, n + 1)]
    factors = []
    for i in range(n):
        for j in range(n - 1, i, -1):
            if perm[j] < perm[j - 1]:
                perm[j], perm[j - 1] = perm[j - 1], perm[j]
                factors.append((j, j + 1))
    return list(reversed(factors))


def cycle_decomposition(tup):
    """Non-trivial cycles, each starting at its smallest element (perm2.py:157-178)."""
    n = len(tup)
    seen, out = set(), []
    for i in range(1, n + 1):
        if i in seen:
            continue
        cur, cyc = i, []
        while cur not in seen:
            cyc.append(cur)
            seen.add(cur)
            cur = tup[cur - 1]
        if len(cyc) > 1:
            out.append(cyc)
    return out


def yor(partition, tup):
    """rho_lambda(sigma) as the product of generators along the bubble-sort word of each
    cycle, multiplied left to right (yor.py:52-85)."""
    n = sum(partition)
    res = None
    for cyc in cycle_decomposition(tup):
        for (a, _b) in cycle_to_adj_transpositions(cyc, n):
            y = yor_trans(partition, a)
            res = y if res is None else res.dot(y)
    if res is None:                                     # identity (yor.py:67-71)
        return np.eye(len(tableaux(partition)))
    return res


def sn(n):
    """S_n as tuples in the reference's enumeration order (perm2.py:214-233, perm.py:174-180)."""
    return list(itertools.permutations(range(1, n + 1)))


def perm_mul(g, h):
    """(g*h)(i) = g(h(i))  (perm2.py:138-145)."""
    return tuple(g[h[i] - 1] for i in range(len(g)))


def perm_inv(g):
    """perm2.py:188-201."""
    out = [0] * len(g)
    for i, v in enumerate(g):
        out[v - 1] = i + 1
    return tuple(out)


def cont_cycle(n, a, b):
    """The contiguous cycle a->a+1->...->b->a inside S_n (perm2.py:111-119; fft.py:97)."""
    t = list(range(1, n + 1))
    for i in range(a, b):
        t[i - 1] = i + 1
    t[b - 1] = a
    return tuple(t)


def lex_rank(tup):
    """Rank of a permutation tuple in ``sn(n)`` order."""
    n = len(tup)
    rank, items = 0, list(range(1, n + 1))
    for i, v in enumerate(tup):
        j = items.index(v)
        rank += j * math.factorial(n - 1 - i)
        items.pop(j)
    return rank


# --------------------------------------------------------------------------------------------
# transforms on dense vectors (f indexed by lexicographic rank)
# --------------------------------------------------------------------------------------------
def fourier_transform(f, partition):
    """Direct sum  sum_sigma f(sigma) rho_lambda(sigma)  (fft.py:115-151)."""
    n = sum(partition)
    res = None
    for r, p in enumerate(sn(n)):
        term = f[r] * yor(partition, p)
        res = term if res is None else res + term
    return res


def ft_full_direct(f, n):
    """fft.py:16-37 with algo='ft2': dict over partitions(n) order."""
    return {p: fourier_transform(f, p) for p in partitions(n)}


@lru_cache(maxsize=None)
def _coset_gather(k):
    """idx[i-1][rank_{k-1}(pi)] = rank_k(c_i * pi), c_i = (i i+1 ... k); this is the closure
    ``f_i = lambda pi: f(cyc * pi)`` of fft.py:97-98 written as an index table."""
    sub = sn(k - 1)
    out = np.empty((k, len(sub)), dtype=np.int64)
    for i in range(1, k + 1):
        cyc = cont_cycle(k, i, k)
        for r, p in enumerate(sub):
            out[i - 1, r] = lex_rank(perm_mul(cyc, p + (k,)))
    return out


@lru_cache(maxsize=None)
def _rho_cycle(partition, i):
    """rho_lambda((i i+1 ... n)) = yor(ferrers, cyc) of fft.py:101."""
    n = sum(partition)
    return yor(partition, cont_cycle(n, i, n))


def fft_all(f):
    """Clausen--Baum recursion (fft.py:72-113) for every irrep at once.

    ``f`` has shape [..., k!] (lexicographic).  The reference recomputes the sub-transforms for
    every lambda; here they are computed once per coset and shared, which changes no arithmetic:
    f_hat(lambda) = sum_i rho_lambda(c_i) . (direct sum over branch_down of f_hat_i(mu)),
    accumulated in i = 1..n order exactly as fft.py:96-112.
    Returns {partition: array[..., d, d]} in partitions(k) order.
    """
    f = np.asarray(f, dtype=np.float64)
    k = math.factorial
    n = 1
    while k(n) < f.shape[-1]:
        n += 1
    assert k(n) == f.shape[-1]
    if n == 1:
        return {(1,): f.reshape(f.shape[:-1] + (1, 1)).copy()}     # fft.py:86-87
    gather = _coset_gather(n)
    subs = [fft_all(f[..., gather[i]]) for i in range(n)]
    out = {}
    for lam in partitions(n):
        d = len(tableaux(lam))
        acc = np.zeros(f.shape[:-1] + (d, d))
        for i in range(n):
            res = np.zeros(f.shape[:-1] + (d, d))
            idx = 0
            for mu in branch_down(lam):                             # fft.py:105-110
                blk = subs[i][mu]
                dm = blk.shape[-1]
                res[..., idx:idx + dm, idx:idx + dm] += blk
                idx += dm
            acc += np.matmul(_rho_cycle(lam, i + 1), res)           # fft.py:112
        out[lam] = acc
    return out


def inverse_direct(fhat, n):
    """f(g) = (1/n!) sum_lambda d_lambda * sum(fhat(lambda) o rho_lambda(g))
    (tile/par_inv_ft.py:27-45, summed over irreps by the caller)."""
    out = np.zeros(math.factorial(n))
    for r, p in enumerate(sn(n)):
        acc = 0.0
        for lam, mat in fhat.items():
            acc += mat.shape[0] * (mat * yor(lam, p)).sum()
        out[r] = acc / math.factorial(n)
    return out


# --------------------------------------------------------------------------------------------
# data order used by the device path; restated here so tests can cross-check
# --------------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def coset_major_to_lex(n):
    """perm[flat] = lexicographic rank of sigma = c^(n)_{i_n} c^(n-1)_{i_{n-1}} ... c^(2)_{i_2},
    flat = sum_k (i_k - 1) (k-1)!  -- the order in which the recursion of fft.py:96-98 visits
    group elements (outermost coset slowest)."""
    out = np.empty(math.factorial(n), dtype=np.int64)
    ident = tuple(range(1, n + 1))

    def rec(k, prefix, base):
        if k == 1:
            out[base] = lex_rank(prefix)
            return
        for i in range(1, k + 1):
            rec(k - 1, perm_mul(prefix, cont_cycle(n, i, k)), base + (i - 1) * math.factorial(k - 1))

    rec(n, ident, 0)
    return out


def packed_layout(n):
    """Offsets of each irrep's row-major d x d block in the packed spectrum, partitions(n) order."""
    offs, dims, o = {}, {}, 0
    for p in partitions(n):
        d = len(tableaux(p))
        offs[p], dims[p] = o, d
        o += d * d
    assert o == math.factorial(n)
    return offs, dims


def pack(fhat, n):
    offs, dims = packed_layout(n)
    some = next(iter(fhat.values()))
    out = np.empty(some.shape[:-2] + (math.factorial(n),))
    for p in partitions(n):
        out[..., offs[p]:offs[p] + dims[p] ** 2] = fhat[p].reshape(some.shape[:-2] + (-1,))
    return out


def unpack(packed, n):
    offs, dims = packed_layout(n)
    return {p: packed[..., offs[p]:offs[p] + dims[p] ** 2].reshape(packed.shape[:-1] + (dims[p], dims[p]))
            for p in partitions(n)}


# --------------------------------------------------------------------------------------------
# sparse forms for large shapes (n >= 10): the same generators without the dense d x d matrix
# --------------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def yor_trans_sparse(partition, k):
    """rho_lambda((k k+1)) as (diag, partner, offdiag) vectors: row i has ``diag[i]`` on the
    diagonal and, when swapping k, k+1 in tableau i gives the standard tableau ``partner[i]``
    (else -1), ``offdiag[i]`` in that column.  Same construction as ``yor_trans``
    (yor.py:87-117, young_tableau.py:305-339), without materialising the matrix."""
    tabs = tableaux(partition)
    index = {t: i for i, t in enumerate(tabs)}
    d = len(tabs)
    diag = np.zeros(d)
    partner = np.full(d, -1, dtype=np.int64)
    off = np.zeros(d)
    for i, t in enumerate(tabs):
        c = _content_maps(t)
        dist = c[k + 1] - c[k]
        diag[i] = 1.0 / dist
        swapped = tuple(tuple(k + 1 if x == k else (k if x == k + 1 else x) for x in row)
                        for row in t)
        j = index.get(swapped)
        if j is not None:
            partner[i] = j
            off[i] = np.sqrt(1 - (1.0 / dist) ** 2)
    return diag, partner, off


def yor_word(tup):
    """The generator word of sigma exactly as ``yor`` multiplies it (yor.py:52-85): for every
    non-trivial cycle the reversed bubble-sort factors, cycles left to right.  Returns [a, ...]
    meaning rho(sigma) = prod_a rho((a a+1))."""
    n = len(tup)
    word = []
    for cyc in cycle_decomposition(tup):
        word.extend(a for (a, _b) in cycle_to_adj_transpositions(cyc, n))
    return word


def yor_columns(partition, tup, cols):
    """Columns ``cols`` of rho_lambda(sigma): the word of ``yor`` applied to unit vectors from
    the right-most factor inwards, each factor as a 2-sparse row operation."""
    d = len(tableaux(partition))
    v = np.zeros((d, len(cols)))
    v[list(cols), range(len(cols))] = 1.0
    for a in reversed(yor_word(tup)):
        diag, partner, off = yor_trans_sparse(partition, a)
        has = partner >= 0
        w = diag[:, None] * v
        w[has] += off[has, None] * v[partner[has]]
        v = w
    return v


def coset_index(tup):
    """Position of sigma in coset-major order (the order in which the
    recursion of fft.py:96-98 visits the group): sigma = c^(n)_{i_n} ... c^(2)_{i_2} with
    c^(k)_i = (i i+1 ... k), i_k = (current sigma)(k); flat = sum_k (i_k - 1)(k-1)!."""
    p = list(tup)
    n = len(p)
    flat = 0
    for k in range(n, 1, -1):
        i = p[k - 1]
        flat += (i - 1) * math.factorial(k - 1)
        # sigma <- (c^(k)_i)^-1 sigma: values above i move down by one, i goes to k (a fixed point)
        p = [(v - 1 if v > i else (k if v == i else v)) for v in p[:k - 1]]
    return flat
