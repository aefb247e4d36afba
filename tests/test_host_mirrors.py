"""Host-side mirrors of the reference's modules (perm2, perm, utils, young_tableau, yor words)
against the golden vectors and the reference's own unit tests.  CPU only."""
import math
import re
import os

import numpy as np
import pytest

import snfft_oracle as O
from snfft_b200 import perm, perm2, utils, yor
from snfft_b200.young_tableau import FerrersDiagram, YoungTableau, wreath_dim


def test_perm2_known_answers():
    # test_perm.py:6-36
    assert perm2.Perm2.from_tup((1, 2, 3)).tup_rep == (1, 2, 3)
    p1 = perm2.Perm2.from_trans((3, 5), 7)
    p2 = perm2.Perm2.from_trans((5, 3), 7)
    assert p1.tup_rep == p2.tup_rep == (1, 2, 5, 4, 3, 6, 7)
    a, b = perm2.Perm2.from_tup((2, 3, 4, 1)), perm2.Perm2.from_tup((4, 1, 2, 3))
    assert (a * b).tup_rep == (b * a).tup_rep
    g = perm2.Perm2.from_tup((2, 3, 4, 1, 5))
    t = perm2.Perm2.from_trans((1, 5), 5)
    assert (g * t).tup_rep == (5, 3, 4, 1, 2)
    assert (t * g).tup_rep == (2, 3, 4, 5, 1)


def test_perm2_group_axioms_and_enumeration():
    s4 = perm2.sn(4)
    assert [p.tup_rep for p in s4] == O.sn(4)
    assert [p.id for p in s4] == list(range(24))
    e = perm2.Perm2.eye(4)
    for p in s4:
        assert (p * p.inv()) == e and (p.inv() * p) == e
        assert p.cycle_decomposition == O.cycle_decomposition(p.tup_rep)
    with pytest.raises(Exception):
        perm2.Perm2.eye(3) * perm2.Perm2.eye(4)
    assert perm2.Perm2.cont_cycle(5, 2, 4).tup_rep == O.cont_cycle(5, 2, 4) == (1, 3, 4, 2, 5)
    assert perm2.Perm2.from_cycle_decomp([(1, 3), (2,), (4, 5)]).tup_rep == (3, 2, 1, 5, 4)


def test_perm_cycle_class_matches_perm2():
    for t in O.sn(4):
        p = perm.Perm.from_lst(t)
        assert p.tup_rep == t
        for u in O.sn(4)[::5]:
            q = perm.Perm.from_lst(u)
            assert (p * q).tup_rep == O.perm_mul(t, u)
        assert (p * p.inverse()).tup_rep == (1, 2, 3, 4)
    # the coset representative of fft.py:97 and its short tuple
    c = perm.Perm([tuple(range(2, 5))])
    assert c.tup_rep == (1, 3, 4, 2)
    assert (c * perm.Perm([(1,)])).tup_rep == (1, 3, 4, 2)
    assert [p.tup_rep for p in perm.sn(3)] == O.sn(3)


@pytest.mark.parametrize('n', range(0, 11))
def test_partitions_order(n):
    assert utils.partitions(n) == O.partitions(n)


def test_utils_reference_tests():
    # test_utils.py:15-36
    assert set(utils.partitions(5)) == {(1,) * 5, (2, 1, 1, 1), (2, 2, 1), (3, 1, 1), (3, 2), (4, 1), (5,)}
    assert set(utils.weak_partitions(6, 2)) == {(6, 0), (5, 1), (2, 4), (3, 3), (0, 6), (1, 5), (4, 2)}
    assert utils.weak_partitions(3, 2) == [(0, 3), (1, 2), (2, 1), (3, 0)]
    irreps = list(utils.cube2_irreps())
    assert len(irreps) == 270
    total = 0
    for alpha, parts in irreps:
        cosets = math.factorial(8) // np.prod([math.factorial(a) for a in alpha])
        total += int(cosets * wreath_dim(parts)) ** 2
    assert total == utils.CUBE2_SIZE            # sum D^2 = |G|
    assert sum(1 for _ in utils.cube2_orientations()) == 3 ** 7
    assert utils.chunk(list(range(7)), 3) == [[0, 1, 6], [2, 3], [4, 5]]


def test_young_tableau_reference_tests():
    # test.py:13-51
    contents = [(1, 3, 5), (2,), (4,)]
    y = YoungTableau(contents, (1, 3, 5, 2, 4))
    for r, row in enumerate(contents):
        for c, val in enumerate(row):
            assert y.get_row(val) == r + 1 and y.get_col(val) == c + 1
    assert YoungTableau([(1, 3, 4), (2,), (5,)], (1, 3, 4, 2, 5)).valid()
    assert not YoungTableau([(1, 2, 4), (5,), (3,)], (1, 2, 4, 5, 3)).valid()
    assert YoungTableau([(1, 3, 5), (2,), (4,)], (1, 3, 5, 2, 4)) < YoungTableau([(1, 3, 4), (2,), (5,)], (1, 3, 4, 2, 5))
    for p, cnt in {(4,): 1, (3, 1): 3, (2, 2): 2, (2, 1, 1): 3, (1, 1, 1, 1): 1}.items():
        assert len(FerrersDiagram(p).gen_tableaux()) == cnt


@pytest.mark.parametrize('n', range(1, 8))
def test_tableaux_match_reference(golden_small, n):
    vals = []
    for p in utils.partitions(n):
        fd = FerrersDiagram(p)
        assert [t.idx for t in fd.tableaux] == list(range(fd.n_tabs()))
        assert fd.tableaux == sorted(fd.tableaux)           # last-letter order
        assert all(YoungTableau.valid_static(p, t.vals) and t.valid() for t in fd.tableaux)
        assert [b.partition for b in fd.branch_down()] == O.branch_down(p)
        vals.extend(t.vals for t in fd.tableaux)
    assert vals == [tuple(int(v) for v in row) for row in golden_small[f'tableaux_{n}']]


def test_gen_tableaux_with_explicit_fillings():
    import itertools
    fd = FerrersDiagram((3, 2))
    perms = [(1,) + p for p in itertools.permutations(range(2, 6))]
    assert [t.vals for t in fd.gen_tableaux(perms)] == [t.vals for t in FerrersDiagram((3, 2)).gen_tableaux()]


def test_transpose_and_distances():
    fd = FerrersDiagram((3, 1))
    t0, t1, t2 = fd.tableaux
    assert t0.transpose((2, 3)) is t1 and t1.transpose((3, 4)) is t2 and t2.transpose((1, 2)) is None
    assert t1.dist(3, 4) == 3 and t1.ax_dist(3, 4) == 3 and t0.content(2) == -1


def test_words_match_reference_algorithm():
    for n in range(2, 7):
        for i in range(1, n):
            cyc = tuple(range(i, n + 1))
            assert yor.cycle_to_adj_transpositions(cyc, n) == [(j, j + 1) for j in range(i, n)]
            assert yor.cycle_to_adj_transpositions(cyc, n) == O.cycle_to_adj_transpositions(cyc, n)
    assert yor.perm_to_adj_transpositions([(1, 3), (4, 5)], 5) == \
        O.cycle_to_adj_transpositions((1, 3), 5) + O.cycle_to_adj_transpositions((4, 5), 5)


def test_no_product_module_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'snfft_b200')
    for dirpath, _dirs, files in os.walk(root):
        for name in files:
            if name.endswith(('.py', '.cu', '.cpp', '.hpp', '.cuh')):
                src = open(os.path.join(dirpath, name)).read()
                assert not re.search(r'^\s*(import|from)\s+(oracle|snfft_oracle)', src, re.M), name


def test_ysemi_known_answers():
    # test.py:53-64
    from snfft_b200.yor import ysemi
    ferr = FerrersDiagram((3, 1))
    p12 = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, 1]])
    p23 = np.array([[0.5, 0.75, 0], [1, -0.5, 0], [0, 0, 1]])
    p34 = np.array([[1, 0, 0], [0, 1. / 3., 8. / 9.], [0, 1, -1. / 3.]])
    assert np.allclose(p12, ysemi(ferr, perm2.Perm2({1: 2, 2: 1}, 4)))
    assert np.allclose(p23, ysemi(ferr, perm2.Perm2({2: 3, 3: 2}, 4)))
    assert np.allclose(p34, ysemi(ferr, perm2.Perm2({3: 4, 4: 3}, 4)))
    assert np.allclose(np.eye(3), ysemi(ferr, perm2.Perm2.eye(4)))
    # a representation: homomorphism on S_4
    g, h = perm2.Perm2.from_tup((2, 3, 1, 4)), perm2.Perm2.from_tup((1, 4, 3, 2))
    assert np.allclose(ysemi(ferr, g) @ ysemi(ferr, h), ysemi(ferr, g * h))


def test_file_formats(tmp_path):
    from snfft_b200 import io
    txt = tmp_path / 'dists.txt'
    txt.write_text('123,0\n213,1\n321,3\n')
    perms, vals = io.read_perm_dist(str(txt))
    assert perms == [(1, 2, 3), (2, 1, 3), (3, 2, 1)] and vals == [0.0, 1.0, 3.0]
    dense = io.dense_from_perm_dist(perms, vals, 3)
    assert dense.tolist() == [0.0, 0.0, 1.0, 0.0, 0.0, 3.0]
    assert io.clean_line('00120,21345,7\n') == ((0, 0, 1, 2, 0), (2, 1, 3, 4, 5), 7)
    spec = {(3,): np.ones((1, 1)), (2, 1): np.eye(2), (1, 1, 1): -np.ones((1, 1))}
    io.save_fourier(str(tmp_path / 'fourier'), spec)
    back = io.load_fourier(str(tmp_path / 'fourier'), 3)
    assert list(back) == utils.partitions(3) and np.array_equal(back[(2, 1)], np.eye(2))
    io.save_wreath_fourier(str(tmp_path), (0, 1, 2), ((), (1,), (2,)), np.eye(3, dtype=np.complex128))
    assert os.path.exists(tmp_path / '(0, 1, 2)' / '((), (1,), (2,)).npy')
