"""Tuple-backed permutations with the reference's interface (reference ``perm2.py``).

Convention (perm2.py:138-145, test_perm.py:21-36): ``tup_rep[i-1] = sigma(i)`` and
``(g * h)(i) = g(h(i))``.  ``sn(n)`` enumerates in ``itertools.permutations`` order, which is the
lexicographic order the device path's ``order='lex'`` refers to.
"""
from __future__ import annotations

from itertools import permutations

SN_CACHE = {}  # n -> {tuple: Perm2}; interning table, same role as perm2.py:8


class Perm2:
    __slots__ = ("size", "tup_rep", "_cycles", "_id")

    def __init__(self, p_map, n, tup_rep=None, cyc_decomp=None):
        """p_map: dict i -> sigma(i) (missing keys are fixed points), n: degree."""
        self.size = n
        self.tup_rep = tuple(tup_rep) if tup_rep is not None else tuple(
            p_map.get(i, i) for i in range(1, n + 1))
        self._cycles = cyc_decomp
        self._id = None
        SN_CACHE.setdefault(n, {}).setdefault(self.tup_rep, self)

    # ---- constructors (perm2.py:70-127) ---------------------------------------------------
    @staticmethod
    def from_tup(tup):
        tup = tuple(tup)
        hit = SN_CACHE.get(len(tup), {}).get(tup)
        return hit if hit is not None else Perm2(None, len(tup), tup)

    @staticmethod
    def eye(size):
        return Perm2.from_tup(range(1, size + 1))

    @staticmethod
    def from_cycle_decomp(decomp_lst):
        n = sum(len(c) for c in decomp_lst)
        image = list(range(1, n + 1))
        for cyc in decomp_lst:
            for pos, val in enumerate(cyc):
                image[val - 1] = cyc[(pos + 1) % len(cyc)]
        return Perm2.from_tup(image)

    @staticmethod
    def swap(n, a, b):
        return Perm2.from_trans((a, b), n)

    @staticmethod
    def from_trans(trans, n):
        image = list(range(1, n + 1))
        a, b = trans
        image[a - 1], image[b - 1] = b, a
        return Perm2.from_tup(image)

    @staticmethod
    def cont_cycle(n, a, b):
        """a -> a+1 -> ... -> b -> a inside S_n (the coset representative of fft.py:97)."""
        image = list(range(1, n + 1))
        for i in range(a, b):
            image[i - 1] = i + 1
        image[b - 1] = a
        return Perm2.from_tup(image)

    # ---- group structure ---------------------------------------------------------------------
    @property
    def _map(self):
        return {i + 1: v for i, v in enumerate(self.tup_rep)}

    @property
    def cycle_decomposition(self):
        """Non-trivial cycles as lists, each led by its smallest element (perm2.py:157-178)."""
        if self._cycles is None:
            seen, cycles = set(), []
            for start in range(1, self.size + 1):
                if start in seen:
                    continue
                cyc, cur = [], start
                while cur not in seen:
                    seen.add(cur)
                    cyc.append(cur)
                    cur = self.tup_rep[cur - 1]
                if len(cyc) > 1:
                    cycles.append(cyc)
            self._cycles = cycles
        return self._cycles

    def __call__(self, x):
        return self.tup_rep[x - 1] if 1 <= x <= self.size else x

    __getitem__ = __call__

    def __mul__(self, other):
        if self.size != other.size:
            # the reference raises here too (perm2.py:141-142); fft2 no longer relies on it
            raise Exception('Currently cant mult two perms of diff sizes!')
        g, h = self.tup_rep, other.tup_rep
        return Perm2.from_tup(g[h[i] - 1] for i in range(self.size))

    def inv(self):
        image = [0] * self.size
        for i, v in enumerate(self.tup_rep):
            image[v - 1] = i + 1
        return Perm2.from_tup(image)

    def to_tup(self):
        return self.tup_rep

    def get_tup_rep(self):
        return self.tup_rep

    def __len__(self):
        return self.size

    def __hash__(self):
        return hash(self.tup_rep)

    def __eq__(self, other):
        return isinstance(other, Perm2) and self.tup_rep == other.tup_rep

    def __repr__(self):
        return str(self.cycle_decomposition)

    @property
    def id(self):
        return self._id

    def set_id(self, _id):
        self._id = _id


class ProdPerm:
    """Element of a direct product of symmetric groups (perm2.py:21-41)."""

    def __init__(self, *perms):
        self.perms = perms
        self.tup_rep = tuple(p.tup_rep for p in perms)

    def __mul__(self, other):
        return ProdPerm(*[a * b for a, b in zip(self.perms, other.perms)])

    def inv(self):
        return ProdPerm(*[p.inv() for p in self.perms])

    def get_tup_rep(self):
        return self.tup_rep

    def __repr__(self):
        return str(self.tup_rep)


def conjugate(x, g):
    return g.inv() * x * g


def sn(n, prefix=None):
    """All of S_n as Perm2 objects in lexicographic tuple order, ids = ranks (perm2.py:214-233)."""
    perms = [Perm2.from_tup(t) for t in permutations(range(1, n + 1))]
    for rank, p in enumerate(perms):
        p.set_id(rank)
    return perms
