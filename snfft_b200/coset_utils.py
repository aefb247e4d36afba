"""Young subgroups and their left cosets, with the reference's names (reference ``coset_utils.py``).

``coset_reps`` keeps the reference's contract -- scanning G in order, the first element of every
new coset (coset_utils.py:72-100) -- and ``young_coset_reps`` builds the same list directly for
G = S_n, H = S_alpha: the lexicographically first element of a left coset g S_alpha is the one
whose values increase inside every alpha-segment, so the representatives are those tuples in
lexicographic order (the scan is O(|G| |H|) permutation products and needs hours at alpha=(2,3,3)).
"""
from __future__ import annotations

import itertools

from . import perm2


def perm_from_young_tuple(cyc_tup):
    return perm2.Perm2.from_tup(tuple(x for seg in cyc_tup for x in seg))


def young_subgroup(weak_partition):
    """Product of the symmetric groups on 1..alpha_k (each factor on its own letters 1..alpha_k)."""
    return itertools.product(*[itertools.permutations(range(1, p + 1)) for p in weak_partition if p > 0])


def young_subgroup_canonical(weak_partition):
    """The same product with the k-th factor acting on its own segment of 1..n."""
    factors, start = [], 1
    for p in weak_partition:
        if p == 0:
            continue
        factors.append(itertools.permutations(range(start, start + p)))
        start += p
    return itertools.product(*factors)


def young_subgroup_perm(weak_partition):
    return [perm_from_young_tuple(t) for t in young_subgroup_canonical(weak_partition)]


def tup_set(perms):
    return {p.tup_rep for p in perms}


def left_coset(g, H):
    return [g * h for h in H]


def right_coset(g, H):
    return [h * g for h in H]


def check_coset(g1, g2, H):
    return g2 in set(left_coset(g1, H))


def check_equal(g1, g2):
    return set(g1) == set(g2)


def coset_reps(G, H):
    """Left-coset representatives of H in G in the order G is given: the first element of every
    coset that has not been seen yet."""
    H = list(H)
    seen, reps = set(), []
    want = len(G) // len(H)
    for g in G:
        if g.tup_rep in seen:
            continue
        reps.append(g)
        seen.update((g * h).tup_rep for h in H)
        if len(reps) == want:
            break
    return reps


def young_coset_reps(weak_partition):
    """coset_reps(sn(n), young_subgroup_perm(alpha)) as tuples, built directly."""
    n = sum(weak_partition)
    bounds, start = [], 0
    for p in weak_partition:
        bounds.append((start, start + p))
        start += p

    def fill(k, remaining):
        if k == len(bounds):
            yield ()
            return
        lo, hi = bounds[k]
        for choice in itertools.combinations(remaining, hi - lo):
            rest = [x for x in remaining if x not in choice]
            for tail in fill(k + 1, rest):
                yield choice + tail

    return sorted(fill(0, list(range(1, n + 1))))
