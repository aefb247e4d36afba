"""The reference-shaped Python entry points (snfft_b200.fft / yor) on the GPU, written the way
the reference's own test.py exercises them."""
import math

import numpy as np
import pytest

import snfft_oracle as O
from conftest import rel_fro
from snfft_b200 import fft as sfft
from snfft_b200 import perm2
from snfft_b200.perm import Perm
from snfft_b200.perm2 import Perm2
from snfft_b200.utils import partitions
from snfft_b200.yor import yor, yor_trans
from snfft_b200.young_tableau import FerrersDiagram

pytestmark = pytest.mark.gpu


def test_yor_known_answers():
    # test.py:66-74
    ferr = FerrersDiagram((3, 1))
    p12 = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, 1]])
    p23 = np.array([[0.5, np.sqrt(3) / 2., 0], [np.sqrt(3) / 2., -0.5, 0], [0, 0, 1]])
    p34 = np.array([[1, 0, 0], [0, 1. / 3., np.sqrt(8) / 3], [0, np.sqrt(8) / 3, -1. / 3.]])
    assert np.allclose(p12, yor(ferr, Perm2({1: 2, 2: 1}, 4)))
    assert np.allclose(p23, yor(ferr, Perm2({2: 3, 3: 2}, 4)))
    assert np.allclose(p34, yor(ferr, Perm2({3: 4, 4: 3}, 4)))
    assert np.allclose(p34, yor_trans(ferr, (3, 4)))
    assert np.allclose(yor(ferr, Perm2.eye(4)), np.eye(3))
    assert np.allclose(yor(ferr, Perm([(3, 4)])), p34)          # short tuple is padded


def test_fft_reference_test(golden_small):
    # test.py:76-85
    f = lambda p: 1 if p[1] == 2 else 1.5   # noqa: E731
    ref_fft = O.unpack(golden_small['testpy_fft_6'], 6)
    # f depends on sigma(1) only, so most irreps are exactly zero in exact arithmetic: measure the
    # error against the norm of the whole spectrum there
    scale = np.linalg.norm(golden_small['testpy_fft_6'])
    for partition in partitions(6):
        ferrers = FerrersDiagram(partition)
        fft_res = sfft.fft(f, ferrers)
        ft = sfft.fourier_transform(f, ferrers)
        ft2 = sfft.fourier_transform2(f, ferrers)
        assert np.allclose(ft, ft2)
        assert np.allclose(fft_res, ft)
        assert fft_res.shape == (ferrers.n_tabs(),) * 2 and fft_res.dtype == np.float64
        assert np.linalg.norm(fft_res - ref_fft[partition]) < 1e-10 * scale
        assert np.linalg.norm(sfft.fft2(f, ferrers) - ref_fft[partition]) < 1e-10 * scale   # raises in the reference


def test_ft_full_all_algos(golden_small):
    n = 5
    values = np.random.default_rng(5).standard_normal(120)
    lut = {p.tup_rep: values[i] for i, p in enumerate(perm2.sn(n))}

    def f(p):
        t = tuple(p.tup_rep)
        return lut[t + tuple(range(len(t) + 1, n + 1))]
    ref = O.unpack(golden_small['direct_5'], 5)
    for algo in ('fft', 'ft1', 'ft2'):
        res = sfft.ft_full(f, n, algo)
        assert [k.partition for k in res] == partitions(n)
        for k, v in res.items():
            assert rel_fro(v, ref[k.partition]) < 1e-10
    back = sfft.ift_full(sfft.ft_full(f, n), n)
    assert rel_fro(back, values) < 1e-10


def test_dense_api_round_trip():
    values = np.random.default_rng(1).standard_normal((4, math.factorial(6)))
    spec = sfft.forward(values)
    assert list(spec) == partitions(6)
    assert rel_fro(sfft.inverse(spec, 6), values) < 1e-10


def test_yor_random_homomorphism_n7():
    # test.py:87-101 through the public API
    rng = np.random.default_rng(3)
    for p in partitions(7)[::3]:
        fd = FerrersDiagram(p)
        g = Perm2.from_tup(tuple(int(x) + 1 for x in rng.permutation(7)))
        h = Perm2.from_tup(tuple(int(x) + 1 for x in rng.permutation(7)))
        assert np.allclose(yor(fd, g) @ yor(fd, h), yor(fd, g * h))


@pytest.mark.parametrize('n', [2, 4, 5, 8, 9])
def test_sampled_inverse_against_the_oracle(n):
    """snfft_inverse_at (all irreps, all states in one launch) against the oracle's restatement of
    tile/par_inv_ft.py:27-45 on a RANDOM spectrum (not the transform of a function: nothing of the
    device forward path enters the expectation)."""
    import torch
    import snfft_oracle as O
    from snfft_b200 import get_plan
    rng = np.random.default_rng(40 + n)
    plan = get_plan(n)
    parts, dims, offs = plan.layout()
    fhat = {p: rng.standard_normal((d, d)) for p, d in zip(parts, dims)}
    if n <= 5:
        perms = list(O.sn(n))
    else:
        perms = [tuple(int(v) + 1 for v in rng.permutation(n)) for _ in range(6)] + [tuple(range(1, n + 1))]
    got = sfft.inverse_at(fhat, n, perms)
    if n <= 5:
        want = O.inverse_direct(fhat, n)
    else:
        nf = math.factorial(n)
        want = np.array([sum(d * (fhat[p] * O.yor(p, g)).sum() for p, d in zip(parts, dims)) / nf for g in perms])
    assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max())


def test_device_seminormal_form_matches_host_and_reference_known_answers():
    """snfft_ysemi (Young's seminormal form on the device, yor.py:119-182) against the host mirror
    (pinned to the reference's own known answers, test.py:53-64) on random permutations, plus the
    homomorphism property at n = 7."""
    from snfft_b200.yor import ysemi, ysemi_batch
    rng = np.random.default_rng(12)
    for p in [(3, 1), (2, 2), (3, 2, 1), (4, 1, 1), (2, 2, 1, 1)]:
        n = sum(p)
        fd = FerrersDiagram(p)
        tups = [tuple(int(x) + 1 for x in rng.permutation(n)) for _ in range(6)] + [tuple(range(1, n + 1))]
        got = ysemi_batch(fd, tups)
        for t, m in zip(tups, got):
            assert np.allclose(m, ysemi(fd, Perm2.from_tup(t)), atol=1e-13, rtol=0)
    fd = FerrersDiagram((3, 2, 2))
    g = tuple(int(x) + 1 for x in rng.permutation(7))
    h = tuple(int(x) + 1 for x in rng.permutation(7))
    gh = (Perm2.from_tup(g) * Perm2.from_tup(h)).tup_rep
    mg, mh, mgh = ysemi_batch(fd, [g, h, gh])
    assert np.allclose(mg @ mh, mgh, atol=1e-12, rtol=0)


def test_sampled_inverse_and_file_job(tmp_path):
    # tile/par_inv_ft.py:27-45 per state, summed over irreps; s8fft.do_all_ffts from a text table
    from snfft_b200 import io
    n = 6
    values = np.random.default_rng(2).standard_normal(math.factorial(n))
    spec = sfft.forward(values, n)
    group = perm2.sn(n)
    pick = [3, 100, 719, 0]
    got = sfft.inverse_at(spec, n, [group[i] for i in pick])
    assert np.allclose(got, values[pick], atol=1e-12)
    txt = tmp_path / 'table.txt'
    txt.write_text(''.join('{},{}\n'.format(''.join(map(str, p.tup_rep)), i % 7) for i, p in enumerate(group)))
    out = io.fft_from_file(str(txt), n, savedir=str(tmp_path / 'fourier'))
    ref = sfft.forward(np.array([i % 7 for i in range(720)], dtype=np.float64), n)
    loaded = io.load_fourier(str(tmp_path / 'fourier'), n)
    for p in partitions(n):
        assert np.array_equal(out[p], ref[p]) and np.array_equal(loaded[p], ref[p])
