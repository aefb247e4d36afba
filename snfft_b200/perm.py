"""Cycle-list permutations with the reference's interface (reference ``perm.py``).

``Perm`` is the class the reference's working ``fft`` hands to the user's function
(fft.py:86-98).  Products concatenate cycle lists and are evaluated right to left
(perm.py:71-88, :137-140), i.e. the same ``(g*h)(i) = g(h(i))`` convention as ``Perm2``.
Unlike ``Perm2`` the degree is implicit: it is the largest letter that appears.
"""
from __future__ import annotations

from itertools import permutations

import numpy as np

from .utils import canonicalize

SN_CACHE = {}


class Perm:
    def __init__(self, cyc_decomp):
        cyc_decomp = [tuple(c) for c in cyc_decomp]
        self._max = max(max(c) for c in cyc_decomp)
        # evaluate the (possibly overlapping) cycles right to left on every letter
        image = {}
        for letter in range(1, self._max + 1):
            at = letter
            for cyc in reversed(cyc_decomp):
                if at in cyc:
                    at = cyc[(cyc.index(at) + 1) % len(cyc)]
            image[letter] = at
        self._map = image
        self.decomp = self._disjoint_cycles()
        self.tup_rep = tuple(image[i] for i in range(1, self._max + 1))
        self._str = None

    def _disjoint_cycles(self):
        todo = set(range(1, self._max + 1))
        cycles = []
        while todo:
            start = cur = min(todo)
            cyc = []
            while True:
                cyc.append(cur)
                todo.discard(cur)
                cur = self._map[cur]
                if cur == start:
                    break
            cycles.append(cyc)
        return canonicalize(cycles)

    @staticmethod
    def from_dict(_dict):
        return Perm.from_lst([_dict[i] for i in range(1, len(_dict) + 1)])

    @staticmethod
    def from_lst(lst):
        lst = list(lst)
        seen, cycles = set(), []
        for start in range(1, len(lst) + 1):
            if start in seen:
                continue
            cyc, cur = [], start
            while cur not in seen:
                seen.add(cur)
                cyc.append(cur)
                cur = lst[cur - 1]
            cycles.append(tuple(cyc))
        return Perm(cycles)

    @property
    def cycle_decomposition(self):
        return self.decomp

    def get_tup_rep(self):
        return self.tup_rep

    def __getitem__(self, x):
        return self._map.get(x, x)

    __call__ = __getitem__

    def __len__(self):
        return len(self.decomp)

    def __mul__(self, other):
        return Perm(list(self.cycle_decomposition) + list(other.cycle_decomposition))

    def inverse(self):
        return Perm(canonicalize([c[::-1] for c in self.decomp]))

    def order(self):
        out = 1
        for c in self.decomp:
            out = out * len(c) // np.gcd(out, len(c))
        return int(out)

    def matrix(self, n=None):
        n = self._max if n is None else n
        mat = np.zeros((n, n))
        for src, dst in self._map.items():
            mat[dst - 1, src - 1] = 1
        return mat

    def __repr__(self):
        if self._str is None:
            self._str = ', '.join(str(c) for c in self.decomp if len(c) > 1)
        return self._str or '()'


def sn(n):
    """S_n as Perm objects in lexicographic order (perm.py:174-180)."""
    if n not in SN_CACHE:
        SN_CACHE[n] = [Perm.from_lst(t) for t in permutations(range(1, n + 1))]
    return SN_CACHE[n]
