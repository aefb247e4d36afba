"""The oracle (oracle/snfft_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and against the known answers in the reference's own
test.py.  CPU only."""
import math

import numpy as np
import pytest

import snfft_oracle as O
from conftest import rel_fro

TOL = 1e-12


def key(p):
    return '_'.join(map(str, p))


@pytest.mark.parametrize('n', range(1, 8))
def test_partitions_and_tableaux_order(golden_small, n):
    parts = O.partitions(n)
    ref = [tuple(int(v) for v in row if v > 0) for row in golden_small[f'partitions_{n}']]
    assert parts == ref
    vals = [tuple(x for row in t for x in row) for p in parts for t in O.tableaux(p)]
    assert vals == [tuple(int(v) for v in row) for row in golden_small[f'tableaux_{n}']]


def test_partitions_5_reference_test():
    # test_utils.py:15-22
    assert set(O.partitions(5)) == {(1, 1, 1, 1, 1), (2, 1, 1, 1), (2, 2, 1), (3, 1, 1), (3, 2), (4, 1), (5,)}


def test_yor_known_answers():
    # test.py:66-74
    p12 = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, 1]])
    p23 = np.array([[0.5, np.sqrt(3) / 2., 0], [np.sqrt(3) / 2., -0.5, 0], [0, 0, 1]])
    p34 = np.array([[1, 0, 0], [0, 1. / 3., np.sqrt(8) / 3], [0, np.sqrt(8) / 3, -1. / 3.]])
    assert np.allclose(O.yor((3, 1), (2, 1, 3, 4)), p12)
    assert np.allclose(O.yor((3, 1), (1, 3, 2, 4)), p23)
    assert np.allclose(O.yor((3, 1), (1, 2, 4, 3)), p34)


def test_tableau_counts_n4():
    # test.py:40-51
    for p, cnt in {(4,): 1, (3, 1): 3, (2, 2): 2, (2, 1, 1): 3, (1, 1, 1, 1): 1}.items():
        assert len(O.tableaux(p)) == cnt


@pytest.mark.parametrize('n', range(2, 8))
def test_yor_trans_bit_exact(golden_small, n):
    for p in O.partitions(n):
        for k in range(1, n):
            assert np.array_equal(O.yor_trans(p, k), golden_small[f'yortrans_{key(p)}_{k}'])


def test_yor_random_perms_bit_exact(golden_small):
    names = [k for k in golden_small.files if k.startswith('yor_')]
    assert len(names) >= 12
    for name in names:
        fields = name.split('_')[1:]
        # partition then permutation: the permutation is the trailing n entries, n = sum(partition)
        nums = list(map(int, fields))
        for cut in range(1, len(nums)):
            if sum(nums[:cut]) == len(nums) - cut:
                p, tup = tuple(nums[:cut]), tuple(nums[cut:])
                break
        assert np.array_equal(O.yor(p, tup), golden_small[name])


@pytest.mark.parametrize('n', range(2, 8))
def test_direct_sum_bit_exact(golden_small, n):
    f = np.random.default_rng(n).standard_normal(math.factorial(n))
    got = O.pack(O.ft_full_direct(f, n), n) if n <= 6 else None
    if got is not None:
        assert np.array_equal(got, golden_small[f'direct_{n}'])
    # the shared-subtransform Clausen recursion agrees with the direct sum per irrep
    fa = O.fft_all(f)
    ref = O.unpack(golden_small[f'direct_{n}'], n)
    for p in O.partitions(n):
        assert rel_fro(fa[p], ref[p]) < TOL


@pytest.mark.parametrize('n', range(2, 7))
def test_clausen_bit_exact(golden_small, n):
    f = np.random.default_rng(n).standard_normal(math.factorial(n))
    assert np.array_equal(O.pack(O.fft_all(f), n), golden_small[f'clausen_{n}'])


def test_reference_test_function(golden_small):
    # test.py:76-85: f(p) = 1 if p[1] == 2 else 1.5 on S_6
    f = np.array([1.0 if t[0] == 2 else 1.5 for t in O.sn(6)])
    assert np.array_equal(O.pack(O.fft_all(f), 6), golden_small['testpy_fft_6'])
    assert np.allclose(O.pack(O.fft_all(f), 6), golden_small['testpy_ft2_6'], rtol=1e-12, atol=1e-12)


def test_inverse_direct(golden_small):
    f = np.random.default_rng(5).standard_normal(120)
    rec = O.inverse_direct(O.fft_all(f), 5)
    assert np.allclose(rec, golden_small['inverse_direct_5'], rtol=0, atol=1e-13)
    assert np.allclose(rec, f, rtol=0, atol=1e-13)


def test_yor_homomorphism_n7():
    # test.py:87-101 (n = 9 there; 7 keeps the CPU suite short)
    rng = np.random.default_rng(0)
    for p in O.partitions(7):
        g = tuple(int(x) + 1 for x in rng.permutation(7))
        h = tuple(int(x) + 1 for x in rng.permutation(7))
        assert np.allclose(O.yor(p, g) @ O.yor(p, h), O.yor(p, O.perm_mul(g, h)))


def test_perm_conventions():
    # test_perm.py:21-36
    assert O.perm_mul((2, 3, 4, 1, 5), (5, 2, 3, 4, 1)) == (5, 3, 4, 1, 2)
    assert O.perm_mul((5, 2, 3, 4, 1), (2, 3, 4, 1, 5)) == (2, 3, 4, 5, 1)
    assert O.perm_mul((2, 3, 4, 1), (4, 1, 2, 3)) == O.perm_mul((4, 1, 2, 3), (2, 3, 4, 1))


def test_coset_major_order_is_a_bijection():
    for n in range(1, 8):
        m = O.coset_major_to_lex(n)
        assert sorted(m.tolist()) == list(range(math.factorial(n)))
    # (2,4,1,5,3) <-> (i5,i4,i3,i2) = (3,4,1,2)
    flat = (3 - 1) * 24 + (4 - 1) * 6 + (1 - 1) * 2 + (2 - 1) * 1
    assert O.sn(5)[O.coset_major_to_lex(5)[flat]] == (2, 4, 1, 5, 3)


def test_sparse_generator_forms_match_the_dense_ones():
    """The helpers that let the checker reach n = 12 (single columns of rho_lambda(sigma), the
    coset-major index of a permutation) against the pinned dense functions at small n."""
    for n in (3, 4, 5, 6):
        c2l = O.coset_major_to_lex(n)
        group = O.sn(n)
        for t in range(0, len(c2l), max(1, len(c2l) // 97)):
            assert O.coset_index(group[c2l[t]]) == t
    rng = np.random.default_rng(0)
    for p in [(3, 2, 1), (4, 1, 1), (2, 2, 2), (5, 1)]:
        for k in range(1, 6):
            diag, partner, off = O.yor_trans_sparse(p, k)
            dense = np.diag(diag)
            rows = np.nonzero(partner >= 0)[0]
            dense[rows, partner[rows]] = off[rows]
            assert np.array_equal(dense, O.yor_trans(p, k))
        for _ in range(5):
            tup = tuple(int(v) + 1 for v in rng.permutation(6))
            cols = [0, len(O.tableaux(p)) - 1, len(O.tableaux(p)) // 2]
            assert np.allclose(O.yor_columns(p, tup, cols), O.yor(p, tup)[:, cols], atol=1e-14, rtol=0)
