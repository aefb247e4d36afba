"""CPU oracle for the wreath-product (C3 wr S_n, 2x2 cube) representations and their direct Fourier
transform -- TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing under snfft_b200/).

Numpy restatement, in complex128, of the reference's wreath path.  Each function cites the lines it
follows.  The reference computes the cyclic scalars in complex64 (wreath.py:53) and densifies in
complex64 (wreath.py:331), so its own results carry ~1e-7 relative error; the golden vectors made
from it (tests/golden/make_golden_wreath.py) are compared at 1e-6 and the 1e-10 parity of the CUDA
path is against this complex128 restatement.
"""
from __future__ import annotations

import itertools
import math
from functools import reduce

import numpy as np

import snfft_oracle as O


def young_subgroup(alpha):
    """S_alpha as tuples of segment permutations of the segments' own letters, in
    itertools.product order (coset_utils.py:31-44 ``young_subgroup_canonical``)."""
    groups, start = [], 1
    for p in alpha:
        if p == 0:
            continue
        groups.append(list(itertools.permutations(range(start, start + p))))
        start += p
    return list(itertools.product(*groups))


def young_subgroup_perm(alpha):
    """The same elements as permutations of 1..n (coset_utils.py:6-29)."""
    return [tuple(x for seg in t for x in seg) for t in young_subgroup(alpha)]


def coset_reps(n, alpha):
    """Left-coset representatives of S_alpha in S_n: scanning S_n in lexicographic order, the first
    element of every coset not seen yet (coset_utils.py:72-100)."""
    H = young_subgroup_perm(alpha)
    seen, reps = set(), []
    for g in O.sn(n):
        if g in seen:
            continue
        reps.append(g)
        for h in H:
            seen.add(O.perm_mul(g, h))
    return reps


def cyclic_irrep(alpha, tup):
    """exp(2 pi i (0*p1 + 1*p2 + 2*p3) / 3) with p_k the sum of the k-th alpha-segment of tup
    (wreath.py:123-143).  For a weak partition into TWO parts this is the C_2 character
    (+1 if p2 is even else -1) of the pyraminx group C2 wr S6 (pyraminx/px_wreath.py:17-23)."""
    m, start, expo = len(alpha), 0, 0
    for k, a in enumerate(alpha):
        expo += k * sum(tup[start:start + a])
        start += a
    if m == 2:
        return complex(1.0 if expo % 2 == 0 else -1.0)
    return np.exp(2j * np.pi * expo / float(m))


def block_cyclic_irreps(alpha, otup, reps):
    """One scalar per coset representative: the character of tup o t_i (wreath.py:34-58)."""
    return np.array([cyclic_irrep(alpha, tuple(otup[t[p] - 1] for p in range(len(otup)))) for t in reps],
                    dtype=np.complex128)


def young_yor(alpha, parts):
    """{h in S_alpha (as a permutation of 1..n): kron of the YOR matrices of the factors}
    (wreath.py:191-224)."""
    out = {}
    nz = [(a, p) for a, p in zip(alpha, parts) if a > 0]
    for segs in young_subgroup(alpha):
        mats, start = [], 1
        for (a, p), seg in zip(nz, segs):
            local = tuple(x - start + 1 for x in seg)      # the factor as a permutation of 1..a
            mats.append(O.yor(tuple(p), local))
            start += a
        out[tuple(x for seg in segs for x in seg)] = reduce(np.kron, mats)
    return out


def wreath_blocks(alpha, parts, g, reps=None, yy=None):
    """{(i, j): block} of the induced representation at the permutation g: block (i, j) is present
    iff t_i^-1 g t_j lies in S_alpha (wreath.py:281-315; literal double scan over representatives)."""
    n = sum(alpha)
    reps = coset_reps(n, alpha) if reps is None else reps
    yy = young_yor(alpha, parts) if yy is None else yy
    out = {}
    for i, ti in enumerate(reps):
        ti_inv = O.perm_inv(ti)
        for j, tj in enumerate(reps):
            h = O.perm_mul(O.perm_mul(ti_inv, g), tj)
            if h in yy:
                out[(i, j)] = yy[h]
                break
    return out


def wreath_rep(alpha, parts, otup, g, reps=None, yy=None):
    """Dense D(o) P(g) (wreath.py:317-343, :387-394) in complex128."""
    n = sum(alpha)
    reps = coset_reps(n, alpha) if reps is None else reps
    blocks = wreath_blocks(alpha, parts, g, reps, yy)
    b = next(iter(blocks.values())).shape[0]
    scal = block_cyclic_irreps(alpha, otup, reps)
    mat = np.zeros((len(reps) * b, len(reps) * b), dtype=np.complex128)
    for (i, j), blk in blocks.items():
        mat[i * b:(i + 1) * b, j * b:(j + 1) * b] = scal[i] * blk
    return mat


def orientations(n, order=3, zero_sum=True):
    """Orientation tuples in product order; zero_sum keeps sum = 0 mod order (utils.py:206-209)."""
    return [t for t in itertools.product(range(order), repeat=n) if not zero_sum or sum(t) % order == 0]


def wreath_ft_direct(alpha, parts, f, oris):
    """f_hat = sum_{(o, g)} f[g][o] rho(o, g) (multi.py:25-46, cube_ft.py:53-61); f is indexed
    [rank of g in sn(n)][index of o in oris]."""
    n = sum(alpha)
    reps = coset_reps(n, alpha)
    yy = young_yor(alpha, parts)
    out = None
    for r, g in enumerate(O.sn(n)):
        blocks = wreath_blocks(alpha, parts, g, reps, yy)
        b = next(iter(blocks.values())).shape[0]
        if out is None:
            out = np.zeros((len(reps) * b, len(reps) * b), dtype=np.complex128)
        for q, o in enumerate(oris):
            scal = block_cyclic_irreps(alpha, o, reps)
            for (i, j), blk in blocks.items():
                out[i * b:(i + 1) * b, j * b:(j + 1) * b] += f[r, q] * scal[i] * blk
    return out


def wreath_inverse_direct(alpha, parts, fhat, oris, order_total):
    """d * tr(rho(g)^H f_hat) / |G| for every group element (cube_ift.py:56-67, fft.py:158-172;
    the caller of the reference divides by the group order, test_inv_fft.py:140)."""
    n = sum(alpha)
    reps = coset_reps(n, alpha)
    yy = young_yor(alpha, parts)
    out = np.zeros((math.factorial(n), len(oris)), dtype=np.complex128)
    for r, g in enumerate(O.sn(n)):
        for q, o in enumerate(oris):
            rho = wreath_rep(alpha, parts, o, g, reps, yy)
            out[r, q] = fhat.shape[0] * np.vdot(rho, fhat) / order_total
    return out
