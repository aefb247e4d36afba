"""On-disk formats either side of the transform.

Inputs: text tables ``perm,dist`` (s8fft.py:16-27, one permutation per line as a digit string) and
``otup,ptup,dist`` (multi.py:18-23, cube_ft.py:24-27).  Outputs: one ``{partition}.npy`` float64
matrix per irrep (s8fft.py:62, tile/par_ft.py:42-51) or ``{alpha}/{parts}.npy`` complex128
(cube_ft.py:110-113), the files that exp/fourier_policy.py:98-107 and dqn.py:145-156 load.
"""
from __future__ import annotations

import math
import os

import numpy as np


def stotup(s):
    return tuple(int(c) for c in s)


def read_perm_dist(fname):
    """[(perm tuple, value)] from ``perm,dist`` lines (s8fft.py:16-27)."""
    perms, vals = [], []
    with open(fname) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            stup, dist = line.split(',')
            perms.append(stotup(stup))
            vals.append(float(dist))
    return perms, vals


def clean_line(line):
    """(otup, ptup, dist) from one ``otup,ptup,dist`` line (multi.py:18-23)."""
    ostr, pstr, dist = line.strip().split(',')
    return stotup(ostr), stotup(pstr), int(dist)


def _lex_rank(tup):
    n, rank, items = len(tup), 0, list(range(1, len(tup) + 1))
    for i, v in enumerate(tup):
        j = items.index(v)
        rank += j * math.factorial(n - 1 - i)
        items.pop(j)
    return rank


def dense_from_perm_dist(perms, vals, n):
    """Dense float64[n!] in ``sn(n)`` (lexicographic) order; permutations that are absent stay 0."""
    out = np.zeros(math.factorial(n))
    for p, v in zip(perms, vals):
        out[_lex_rank(p)] += v
    return out


def save_fourier(savedir, fhat):
    """One ``{partition}.npy`` per irrep, named like s8fft.py:62 (``str(partition)``)."""
    os.makedirs(savedir, exist_ok=True)
    for key, mat in fhat.items():
        part = key.partition if hasattr(key, 'partition') else tuple(key)
        np.save(os.path.join(savedir, '{}.npy'.format(part)), np.asarray(mat))


def load_fourier(savedir, n):
    from .utils import partitions
    return {p: np.load(os.path.join(savedir, '{}.npy'.format(p))) for p in partitions(n)}


def save_wreath_fourier(savedir, alpha, parts, fhat):
    """``{savedir}/{alpha}/{parts}.npy`` complex128, as cube_ft.py:110-113 writes it."""
    sub = os.path.join(savedir, str(tuple(alpha)))
    os.makedirs(sub, exist_ok=True)
    np.save(os.path.join(sub, '{}.npy'.format(tuple(tuple(p) for p in parts))), np.asarray(fhat))


def fft_from_file(fname, n, savedir=None):
    """The job of s8fft.do_all_ffts: read ``perm,dist``, transform every irrep on the device,
    optionally save; returns {partition: ndarray}."""
    from . import fft
    perms, vals = read_perm_dist(fname)
    spec = fft.forward(dense_from_perm_dist(perms, vals, n), n, order='lex')
    if savedir is not None:
        save_fourier(savedir, spec)
    return spec
