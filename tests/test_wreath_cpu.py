"""Wreath-product oracle (oracle/wreath_oracle.py) against golden vectors made by the unmodified
reference (tests/golden/make_golden_wreath.py), and the host-side coset helpers.  CPU only.
The reference computes in complex64 (wreath.py:53, :331): agreement is to ~1e-6."""
import itertools
import math
import os

import numpy as np
import pytest

import snfft_oracle as O
import wreath_oracle as W
from conftest import GOLDEN
from snfft_b200 import coset_utils as cu
from snfft_b200 import perm2

CASES = [((0, 3, 2), ((), (2, 1), (1, 1))), ((1, 2, 2), ((1,), (1, 1), (2,))), ((2, 2, 1), ((2,), (1, 1), (1,)))]


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(GOLDEN, 'reference_wreath.npz'))


@pytest.mark.parametrize('ci', range(3))
def test_coset_reps_and_young_subgroup_order(golden, ci):
    alpha, _ = CASES[ci]
    n = sum(alpha)
    assert [tuple(r) for r in golden[f'case{ci}_reps']] == W.coset_reps(n, alpha) == cu.young_coset_reps(alpha)
    assert [tuple(r) for r in golden[f'case{ci}_young']] == W.young_subgroup_perm(alpha)
    assert [p.tup_rep for p in cu.young_subgroup_perm(alpha)] == W.young_subgroup_perm(alpha)
    assert [p.tup_rep for p in cu.coset_reps(perm2.sn(n), cu.young_subgroup_perm(alpha))] == W.coset_reps(n, alpha)


def test_reference_coset_test():
    # test_coset.py:26-30: alpha = (2,2,2,1) in S_7: count and coset equality
    alpha = (2, 2, 2, 1)
    reps = cu.young_coset_reps(alpha)
    H = cu.young_subgroup_perm(alpha)
    assert len(reps) == math.factorial(7) // len(H) == 630
    seen = set()
    for t in reps[::37]:
        coset = {(perm2.Perm2.from_tup(t) * h).tup_rep for h in H}
        assert min(coset) == t and not (coset & seen)
        seen |= coset


@pytest.mark.parametrize('ci', range(3))
def test_representation_matrices(golden, ci):
    alpha, parts = CASES[ci]
    for g, o, ref in zip(golden[f'case{ci}_g'], golden[f'case{ci}_o'], golden[f'case{ci}_rho']):
        mine = W.wreath_rep(alpha, parts, tuple(int(x) for x in o), tuple(int(x) for x in g))
        assert np.abs(mine - ref).max() < 1e-6


@pytest.mark.parametrize('ci', range(3))
def test_direct_transform(golden, ci):
    alpha, parts = CASES[ci]
    n = sum(alpha)
    oris = W.orientations(n)
    f = np.random.default_rng(100 + ci).standard_normal((math.factorial(n), len(oris)))
    mine = W.wreath_ft_direct(alpha, parts, f, oris)
    ref = golden[f'case{ci}_fhat']
    assert np.linalg.norm(mine - ref) / np.linalg.norm(ref) < 1e-6


def test_homomorphism_with_orientations():
    # test_wreath.py:16-35 extended to the full wreath multiplication (o, s)(o', s') = (o + s.o', s s')
    alpha, parts = CASES[0]
    rng = np.random.default_rng(4)
    for _ in range(5):
        g = tuple(int(x) + 1 for x in rng.permutation(5))
        h = tuple(int(x) + 1 for x in rng.permutation(5))
        og = tuple(int(x) for x in rng.integers(0, 3, 5))
        oh = tuple(int(x) for x in rng.integers(0, 3, 5))
        ginv = O.perm_inv(g)
        moved = tuple(oh[ginv[i] - 1] for i in range(5))                  # g . oh  (wreath.py:14-24)
        o_gh = tuple((a + b) % 3 for a, b in zip(og, moved))
        lhs = W.wreath_rep(alpha, parts, og, g) @ W.wreath_rep(alpha, parts, oh, h)
        assert np.allclose(lhs, W.wreath_rep(alpha, parts, o_gh, O.perm_mul(g, h)))
    ident = W.wreath_rep(alpha, parts, (0,) * 5, (1, 2, 3, 4, 5))
    assert np.allclose(ident, np.eye(ident.shape[0]))                      # test_wreath.py:76-85


def test_cube2_inv_fft_part_matches_oracle_inverse():
    # fft.py:153-172 / test_inv_fft.py:128-141: dim * tr(rho(g^-1) f_hat) per element.  The block
    # dictionary comes from the oracle here (the mirror of wreath_yor needs the device), so this
    # checks the host assembly: group inverse, cyclic scalars per coset, dense matrix, trace.
    from snfft_b200 import fft as F
    from snfft_b200 import wreath as MW
    alpha, parts = CASES[1]
    n = sum(alpha)
    reps = W.coset_reps(n, alpha)
    yy = W.young_yor(alpha, parts)
    irrep_dict = {g: W.wreath_blocks(alpha, parts, g, reps, yy) for g in O.sn(n)}
    oris = W.orientations(n)
    rng = np.random.default_rng(3)
    dim = len(reps) * next(iter(yy.values())).shape[0]
    fhat = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    want = W.wreath_inverse_direct(alpha, parts, fhat, oris, 1)
    rep_objs = [perm2.Perm2.from_tup(t) for t in reps]   # the reference holds Perm2 objects here
    ift = F.cube2_inv_fft_func(irrep_dict, fhat, rep_objs, MW.cyclic_irreps(alpha))
    perms = O.sn(n)
    for r in (0, 7, 53, 119):
        for q in (0, 5, len(oris) - 1):
            got = ift(tuple(oris[q]), perms[r])
            assert abs(got - want[r, q]) < 1e-9 * max(1.0, abs(want[r, q]))

