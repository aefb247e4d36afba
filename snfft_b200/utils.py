"""Combinatorial helpers with the reference's names and orderings (reference ``utils.py``).

Only the helpers that sit on the transform path are mirrored: partition enumeration (which
*defines* the irrep order of every layout, utils.py:52-67), weak partitions and the 2x2-cube
irrep census (utils.py:69-132, :206-209), and ``chunk`` (utils.py:35-47).
"""
from __future__ import annotations

from itertools import product

CUBE2_SIZE = 88179840  # 3^7 * 8!  (utils.py:12)


def partitions(n, start=1):
    """Partitions of n, largest part first; order: (n,), then by smallest part i = start.. with
    the partitions of n-i (parts >= i) in the same order (utils.py:52-67)."""
    if n == 0:
        return [()]
    found = [(n,)]
    smallest = start
    while 2 * smallest <= n:
        found.extend(head + (smallest,) for head in partitions(n - smallest, smallest))
        smallest += 1
    return found


def weak_partitions(n, k):
    """k-tuples of non-negative ints summing to n, first coordinate ascending (utils.py:69-91)."""
    if n < 0 or k == 0:
        return []
    if k == 1:
        return [(n,)]
    return [(first,) + rest for first in range(n + 1) for rest in weak_partitions(n - first, k - 1)]


def partition_parts(partition):
    """All ways to pick one partition of each entry (utils.py:93-95)."""
    return product(*[partitions(p) for p in partition])


def size_partition(weak_parts):
    total = 1
    for p in weak_parts:
        total *= len(partitions(p))
    return total


def cube2_alphas():
    """Representatives of the 15 cyclic-shift classes of weak partitions of 8 into 3, in the
    reference's listed order (utils.py:104-124)."""
    return [(2, 3, 3), (4, 2, 2), (3, 1, 4), (3, 4, 1), (1, 2, 5), (1, 5, 2), (0, 4, 4), (5, 0, 3),
            (5, 3, 0), (6, 1, 1), (2, 6, 0), (2, 0, 6), (0, 1, 7), (0, 7, 1), (8, 0, 0)]


def cube2_irreps():
    """(alpha, parts) for the 270 irreps of the 2x2 cube group (utils.py:126-132)."""
    for alpha in cube2_alphas():
        for parts in partition_parts(alpha):
            yield (alpha, parts)


def cube2_orientations():
    """Orientation 8-tuples over Z_3 with zero sum, in product order (utils.py:206-209)."""
    for tup in product((0, 1, 2), repeat=8):
        if sum(tup) % 3 == 0:
            yield tup


def chunk(lst, n):
    """n nearly equal chunks; the remainder is dealt from the end of the list (utils.py:35-47)."""
    size = len(lst) // n
    pieces = [list(lst[i * size:(i + 1) * size]) for i in range(n)]
    for extra in range(len(lst) - size * n):
        pieces[extra].append(lst[-(extra + 1)])
    return pieces


def canonicalize(cycle_decomp):
    """Rotate every cycle so that its smallest element comes first (utils.py:154-168)."""
    out = []
    for cyc in cycle_decomp:
        cyc = list(cyc)
        at = cyc.index(min(cyc))
        out.append(cyc[at:] + cyc[:at])
    return out
