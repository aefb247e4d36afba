"""Wreath-product (C3 wr S_n) transform on the GPU against the complex128 oracle (1e-10) and the
reference's golden vectors (1e-6: the reference works in complex64)."""
import itertools
import math
import os
import random

import numpy as np
import pytest
import torch

import snfft_oracle as O
import wreath_oracle as W
from conftest import GOLDEN
from snfft_b200 import perm2
from snfft_b200.wreath import (WreathCycSn, WreathPlan, cyclic_irreps, get_mat, wreath_rep, wreath_yor)
from snfft_b200.coset_utils import coset_reps, young_subgroup_perm

pytestmark = pytest.mark.gpu
CASES = [((0, 3, 2), ((), (2, 1), (1, 1))), ((1, 2, 2), ((1,), (1, 1), (2,))), ((2, 2, 1), ((2,), (1, 1), (1,)))]


@pytest.mark.parametrize('ci', range(3))
def test_forward_and_inverse_match_oracle(ci):
    alpha, parts = CASES[ci]
    n = sum(alpha)
    golden = np.load(os.path.join(GOLDEN, 'reference_wreath.npz'))
    oris = W.orientations(n)
    f = np.random.default_rng(100 + ci).standard_normal((math.factorial(n), len(oris)))
    wp = WreathPlan(alpha, parts)
    assert [tuple(o) for o in wp.oris.tolist()] == oris
    got = wp.forward(torch.from_numpy(f).cuda())
    ref = W.wreath_ft_direct(alpha, parts, f, oris)
    assert np.linalg.norm(got.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-10
    assert np.linalg.norm(got.cpu().numpy() - golden[f'case{ci}_fhat']) / np.linalg.norm(ref) < 1e-6
    inv = wp.inverse(got).cpu().numpy()
    want = W.wreath_inverse_direct(alpha, parts, ref, oris, math.factorial(n) * len(oris))
    assert np.abs(inv - want).max() < 1e-10 * max(1.0, np.abs(want).max())


def test_wreath_yor_homomorphism():
    # test_wreath.py:16-35 with the reference-shaped helpers (dictionary of blocks per permutation)
    alpha, parts = (0, 3, 2), ((), (2, 1), (1, 1))
    ydict = wreath_yor(alpha, parts)
    keys = list(ydict.keys())
    rng = random.Random(0)
    for _ in range(5):
        g, h = rng.choice(keys), rng.choice(keys)
        gh = perm2.Perm2.from_tup(g) * perm2.Perm2.from_tup(h)
        assert np.allclose(get_mat(g, ydict) @ get_mat(h, ydict), get_mat(gh, ydict))
    eye = perm2.Perm2.eye(5)
    assert np.allclose(get_mat(eye, ydict), np.eye(20))
    # same blocks as the oracle's literal scan, including the cyclic scalars
    reps = coset_reps(perm2.sn(5), young_subgroup_perm(alpha))
    o = (1, 0, 2, 2, 1)
    mine = wreath_rep(o, keys[17], ydict, reps, cyclic_irreps(alpha))
    assert np.allclose(mine, W.wreath_rep(alpha, parts, o, keys[17]))
    w = WreathCycSn.from_tup((1, 0, 2, 2, 1), keys[17], 3)
    assert (w * w.inv()).tup_rep == ((0,) * 5, (1, 2, 3, 4, 5))


def test_plancherel_over_all_irreps_of_c3_wr_s4():
    """sum over all irreps (alpha weak partition of 4 into 3, parts) of D ||f_hat||_F^2 = |G| ||f||^2
    on the FULL group C_3^4 x S_4 (all 81 orientations)."""
    n = 4
    oris = W.orientations(n, zero_sum=False)
    f = np.random.default_rng(9).standard_normal((24, len(oris)))
    fd = torch.from_numpy(f).cuda()
    from snfft_b200.utils import partition_parts, weak_partitions
    total, dims = 0.0, 0
    for alpha in weak_partitions(n, 3):
        for parts in partition_parts(alpha):
            wp = WreathPlan(alpha, parts, oris=oris)
            fh = wp.forward(fd)
            total += wp.D * (fh.abs() ** 2).sum().item()
            dims += wp.D ** 2
    assert dims == 24 * 81
    assert abs(total - 24 * 81 * (f ** 2).sum()) / total < 1e-12


@pytest.mark.parametrize('alpha,parts', [((0, 1, 7), ((), (1,), (4, 3))),           # multi.py:177, D = 112
                                         ((8, 0, 0), ((7, 1), (), ())),             # cube_ft.py:127-128, D = 7
                                         ((1, 2, 5), ((1,), (1, 1), (3, 2))),       # test_wreath.py:77-78, D = 840
                                         ((2, 3, 3), ((2,), (1, 1, 1), (1, 1, 1))), # cube_irrep.py:179-180, D = 560
                                         ((4, 2, 2), ((4,), (2,), (2,))),           # gen_sparse.py:135-137, D = 420
                                         ((2, 3, 3), ((1, 1), (1, 1, 1), (2, 1))),  # test_wreath.py:53-54, D = 1120
                                         ((0, 4, 4), ((), (2, 2), (3, 1)))])        # cube_sp_irrep.py:79-80, D = 420
def test_cube_irreps_sparse_support(alpha, parts):
    """BASELINE configs[3]: the 2x2 cube group (3^7 * 8! elements).  f supported on a few group
    elements: f_hat must equal sum_j c_j rho(g_j) built by the oracle."""
    wp = WreathPlan(alpha, parts)
    assert wp.order == 88179840 and len(wp.oris) == 2187
    rng = np.random.default_rng(38)
    f = torch.zeros((40320, 2187), dtype=torch.float64, device='cuda')
    reps, yy = W.coset_reps(8, alpha) if wp.N <= 60 else wp.reps, None
    want = np.zeros((wp.D, wp.D), dtype=np.complex128)
    ori_index = {tuple(o): q for q, o in enumerate(wp.oris.tolist())}
    for _ in range(2 if wp.N > 100 else 4):
        g = tuple(int(x) + 1 for x in rng.permutation(8))
        o = tuple(int(x) for x in wp.oris[int(rng.integers(2187))])
        c = float(rng.standard_normal())
        f[O.lex_rank(g), ori_index[o]] += c
        want += c * sparse_rho(wp, alpha, parts, o, g)
    got = wp.forward(f).cpu().numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-10


def sparse_rho(wp, alpha, parts, o, g):
    """rho(o, g) via the coset of g t_j (same result as the oracle's double scan, checked in
    test_wreath_yor_homomorphism; the scan needs N^2 products per element)."""
    index = {t: i for i, t in enumerate(wp.reps)}
    b = wp.block
    scal = W.block_cyclic_irreps(alpha, o, wp.reps)
    mat = np.zeros((wp.D, wp.D), dtype=np.complex128)
    nz = [(a, p) for a, p in zip(alpha, parts) if a > 0]
    for j, tj in enumerate(wp.reps):
        gt = O.perm_mul(g, tj)
        rep, start = [], 0
        for a in alpha:
            rep.extend(sorted(gt[start:start + a]))
            start += a
        i = index[tuple(rep)]
        h = O.perm_mul(O.perm_inv(wp.reps[i]), gt)
        mats, start = [], 1
        for a, p in nz:
            mats.append(O.yor(tuple(p), tuple(x - start + 1 for x in h[start - 1:start - 1 + a])))
            start += a
        blk = mats[0]
        for m in mats[1:]:
            blk = np.kron(blk, m)
        mat[i * b:(i + 1) * b, j * b:(j + 1) * b] = scal[i] * blk
    return mat


def test_plancherel_over_all_270_cube_irreps():
    """sum over the 270 irreps of the 2x2 cube group of D ||f_hat||_F^2 = |G| ||f||^2
    (utils.cube2_irreps: 15 alphas up to cyclic shift x the partition triples), sum D^2 = |G|."""
    from snfft_b200.utils import cube2_irreps
    irreps = list(cube2_irreps())
    assert len(irreps) == 270
    f = torch.randn((40320, 2187), dtype=torch.float64, device='cuda',
                    generator=torch.Generator(device='cuda').manual_seed(38))
    total, dims = 0.0, 0
    for alpha, parts in sorted(irreps, key=lambda ap: ap[0]):       # one alpha after the other: cached tables
        wp = WreathPlan(alpha, parts)
        fh = wp.forward(f)
        total += wp.D * (fh.abs() ** 2).sum().item()
        dims += wp.D ** 2
        del fh, wp
    assert dims == 88179840
    want = 88179840 * (f ** 2).sum().item()
    assert abs(total - want) / want < 1e-11


@pytest.mark.parametrize('ci', range(3))
def test_sampled_wreath_inverse_against_oracle(ci):
    """WreathPlan.inverse_at (d tr(rho(g)^H f_hat) / |G| per state, cube_ift.py:56-67) on a RANDOM
    complex f_hat against the oracle's direct evaluation, C3 wr S5."""
    alpha, parts = CASES[ci]
    n = sum(alpha)
    oris = W.orientations(n)
    wp = WreathPlan(alpha, parts)
    rng = np.random.default_rng(300 + ci)
    fhat = rng.standard_normal((wp.D, wp.D)) + 1j * rng.standard_normal((wp.D, wp.D))
    want = W.wreath_inverse_direct(alpha, parts, fhat, oris, math.factorial(n) * len(oris))   # [n!, n_ori]
    group = list(itertools.permutations(range(1, n + 1)))
    picks = [(int(rng.integers(len(group))), int(rng.integers(len(oris)))) for _ in range(12)]
    got = wp.inverse_at(torch.from_numpy(fhat).cuda(), [(oris[q], group[s]) for s, q in picks]).cpu().numpy()
    ref = np.array([want[s, q] for s, q in picks])
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    host = wp.inverse_at_host(torch.from_numpy(fhat).cuda(), [(oris[q], group[s]) for s, q in picks]).cpu().numpy()
    assert np.abs(host - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    # and the dense device inverse agrees with the sampled one on the cube-sized path too
    dense = wp.inverse(torch.from_numpy(fhat).cuda()).cpu().numpy()
    assert np.abs(dense - want).max() < 1e-10 * max(1.0, np.abs(want).max())


def test_pyraminx_group_c2_wr_sn():
    """Row f4: the same transform for the pyraminx group C2 wr S_n (pyraminx/px_wreath.py): irreps
    indexed by a weak partition into TWO parts, characters +-1.  Parity against the oracle on
    C2 wr S5 and Plancherel over all irreps of C2 wr S4 (sum D^2 = 2^4 4!)."""
    from snfft_b200.utils import partition_parts, weak_partitions
    n = 5
    oris = W.orientations(n, order=2, zero_sum=False)
    f = np.random.default_rng(26).standard_normal((math.factorial(n), len(oris)))
    for alpha, parts in [((2, 3), ((1, 1), (2, 1))), ((4, 1), ((2, 2), (1,))), ((0, 5), ((), (3, 2)))]:
        wp = WreathPlan(alpha, parts, oris=oris)
        got = wp.forward(torch.from_numpy(f).cuda()).cpu().numpy()
        ref = W.wreath_ft_direct(alpha, parts, f, oris)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10
        assert np.abs(got.imag).max() < 1e-12
        inv = wp.inverse(torch.from_numpy(ref).cuda()).cpu().numpy()
        want = W.wreath_inverse_direct(alpha, parts, ref, oris, math.factorial(n) * len(oris))
        assert np.abs(inv - want).max() < 1e-10 * max(1.0, np.abs(want).max())
    n = 4
    oris = W.orientations(n, order=2, zero_sum=False)
    f = np.random.default_rng(27).standard_normal((24, len(oris)))
    fd = torch.from_numpy(f).cuda()
    total, dims = 0.0, 0
    for alpha in weak_partitions(n, 2):
        for parts in partition_parts(alpha):
            wp = WreathPlan(alpha, parts, oris=oris)
            total += wp.D * (wp.forward(fd).abs() ** 2).sum().item()
            dims += wp.D ** 2
    assert dims == 24 * 16
    assert abs(total - 24 * 16 * (f ** 2).sum()) / total < 1e-12


def _c_wreath(alpha, parts, n, all_orientations=0):
    import ctypes as C
    from snfft_b200 import _native as N
    a = np.array(alpha, dtype=np.int32)
    p = np.zeros((3, n), dtype=np.int32)
    for k, part in enumerate(parts):
        p[k, :len(part)] = part
    h = C.c_void_p()
    N.check(N.lib.snfft_wreath_create(n, N.ptr(a), N.ptr(p), all_orientations, 0, C.byref(h)))
    dims = np.zeros(6, dtype=np.int64)
    N.check(N.lib.snfft_wreath_dims(h, N.ptr(dims)))
    return h, [int(v) for v in dims]


def test_c_abi_wreath_plan_without_the_python_classes():
    """snfft_wreath_create / _forward / _inverse driven through ctypes alone: the C
    library builds coset representatives, characters, Young-subgroup matrices and the gather index
    itself.  Against the oracle on C3 wr S5, against WreathPlan on the cube group."""
    import ctypes as C
    from snfft_b200 import _native as N
    for ci, (alpha, parts) in enumerate(CASES):
        n = sum(alpha)
        oris = W.orientations(n)
        f = np.random.default_rng(500 + ci).standard_normal((math.factorial(n), len(oris)))
        h, dims = _c_wreath(alpha, parts, n)
        D = dims[0]
        assert dims[4] == math.factorial(n) and dims[5] == len(oris)
        fd = torch.from_numpy(f).cuda()
        re, im = torch.empty((D, D), dtype=torch.float64, device='cuda'), torch.empty((D, D), dtype=torch.float64, device='cuda')
        N.check(N.lib.snfft_wreath_forward(h, N.ptr(fd), N.ptr(re), N.ptr(im), None))
        got = torch.complex(re, im).cpu().numpy()
        ref = W.wreath_ft_direct(alpha, parts, f, oris)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10
        ore, oim = torch.empty_like(fd), torch.empty_like(fd)
        N.check(N.lib.snfft_wreath_inverse(h, N.ptr(re), N.ptr(im), N.ptr(ore), N.ptr(oim), None))
        want = W.wreath_inverse_direct(alpha, parts, ref, oris, math.factorial(n) * len(oris))
        inv = torch.complex(ore, oim).cpu().numpy()
        assert np.abs(inv - want).max() < 1e-10 * max(1.0, np.abs(want).max())
        N.check(N.lib.snfft_wreath_destroy(h))
    # the cube group: same numbers as the Python plan (same kernels underneath, tables built in C++)
    for alpha, parts in [((2, 3, 3), ((1, 1), (1, 1, 1), (2, 1))), ((8, 0, 0), ((7, 1), (), ())), ((0, 1, 7), ((), (1,), (4, 3)))]:
        f = torch.randn((40320, 2187), dtype=torch.float64, device='cuda',
                        generator=torch.Generator(device='cuda').manual_seed(77))
        h, dims = _c_wreath(alpha, parts, 8)
        D = dims[0]
        re, im = torch.empty((D, D), dtype=torch.float64, device='cuda'), torch.empty((D, D), dtype=torch.float64, device='cuda')
        N.check(N.lib.snfft_wreath_forward(h, N.ptr(f), N.ptr(re), N.ptr(im), None))
        want = WreathPlan(alpha, parts).forward(f)
        got = torch.complex(re, im)
        assert ((got - want).abs().max() / want.abs().max()).item() < 1e-12
        N.check(N.lib.snfft_wreath_destroy(h))
        del f, re, im, want, got
        torch.cuda.empty_cache()
    # bad arguments come back as statuses, not crashes
    bad = np.array([1, 1, 1], dtype=np.int32)
    assert N.lib.snfft_wreath_create(5, N.ptr(bad), N.ptr(np.zeros((3, 5), dtype=np.int32)), 0, 0,
                                     C.byref(C.c_void_p())) == 1


@pytest.mark.parametrize('m,count', [(1, 1), (3, 5), (4, 8), (5, 3)])
def test_c3_dft_building_block_against_numpy(m, count):
    """snfft_c3_dft alone: the forward shares one complex transform between two REAL rows (odd
    row counts leave the last one alone), the adjoint writes planes or interleaved complex128."""
    import ctypes as C
    from snfft_b200 import _native as N
    size = 3 ** m
    rng = np.random.default_rng(10 * m + count)
    f = rng.standard_normal((count, size))
    freq = rng.choice(size, size=min(size, 7), replace=False).astype(np.int32)
    digits = np.array([[(e // 3 ** (m - 1 - a)) % 3 for a in range(m)] for e in range(size)])   # first letter most significant
    E = np.exp(2j * np.pi * (digits @ digits.T % 3) / 3.0)                                        # omega^{k.o}
    fd, qd = torch.from_numpy(f).cuda(), torch.from_numpy(freq).cuda()
    re = torch.full((count, len(freq)), float('nan'), dtype=torch.float64, device='cuda')
    im = torch.full_like(re, float('nan'))
    N.check(N.lib.snfft_c3_dft(m, N.ptr(fd), None, N.ptr(qd), len(freq), count, N.ptr(re), N.ptr(im), 0, None))
    want = f @ E[:, freq]
    got = torch.complex(re, im).cpu().numpy()
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
    # adjoint: out[o] = sum_k in[k] omega^{-k.o}
    wantT = want @ E[freq, :].conj()
    ore = torch.empty((count, size), dtype=torch.float64, device='cuda')
    oim = torch.empty_like(ore)
    N.check(N.lib.snfft_c3_dft(m, N.ptr(re), N.ptr(im), N.ptr(qd), len(freq), count, N.ptr(ore), N.ptr(oim), 1, None))
    assert np.abs(torch.complex(ore, oim).cpu().numpy() - wantT).max() < 1e-12 * np.abs(wantT).max()
    both = torch.empty((count, size), dtype=torch.complex128, device='cuda')
    N.check(N.lib.snfft_c3_dft(m, N.ptr(re), N.ptr(im), N.ptr(qd), len(freq), count, N.ptr(both), None, 1, None))
    assert torch.equal(both, torch.complex(ore, oim))
    # interleaved complex128 on both sides
    spec = torch.empty((count, len(freq)), dtype=torch.complex128, device='cuda')
    N.check(N.lib.snfft_c3_dft(m, N.ptr(fd), None, N.ptr(qd), len(freq), count, N.ptr(spec), None, 0, None))
    assert torch.equal(spec, torch.complex(re, im))
    N.check(N.lib.snfft_c3_dft(m, N.ptr(spec), None, N.ptr(qd), len(freq), count, N.ptr(both), None, 1, None))
    assert torch.equal(both, torch.complex(ore, oim))


@pytest.mark.parametrize('alpha,parts', [((0, 4, 1), ((), (2, 2), (1,))), ((4, 0, 1), ((3, 1), (), (1,))),
                                         ((1, 0, 4), ((1,), (), (2, 1, 1)))])
def test_factor_recursion_against_oracle(alpha, parts):
    """Stage 2 through the level recursion of the Young-subgroup factors (a factor S_4 here, several
    cosets, blocks too large for the fused kernel): forward and inverse against the oracle's direct
    evaluation on C3 wr S5."""
    n = sum(alpha)
    oris = W.orientations(n)
    f = np.random.default_rng(700 + alpha[0]).standard_normal((math.factorial(n), len(oris)))
    import snfft_b200 as sb
    wp = WreathPlan(alpha, parts)
    sb.timing_enable(True)
    sb.timing_read()
    got = wp.forward(torch.from_numpy(f).cuda()).cpu().numpy()
    kinds = sb.timing_read()
    sb.timing_enable(False)
    # the S_4 factor really went through the level kernels (levels 2..4 fused), once per plane
    assert kinds['fused_fwd'][1] == 2 and kinds['dgemm'][1] == 0
    ref = W.wreath_ft_direct(alpha, parts, f, oris)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10
    rng = np.random.default_rng(701)
    fhat = rng.standard_normal((wp.D, wp.D)) + 1j * rng.standard_normal((wp.D, wp.D))
    inv = wp.inverse(torch.from_numpy(fhat).cuda()).cpu().numpy()
    want = W.wreath_inverse_direct(alpha, parts, fhat, oris, math.factorial(n) * len(oris))
    assert np.abs(inv - want).max() < 1e-10 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize('alpha,parts', [((2, 3, 3), ((1, 1), (1, 1, 1), (2, 1))),      # fused small blocks
                                         ((0, 1, 7), ((), (1,), (4, 3))),             # one large factor, 8 cosets
                                         ((0, 4, 4), ((), (2, 2), (3, 1))),           # two factors S_4 x S_4
                                         ((8, 0, 0), ((6, 2), (), ()))])              # one coset: S_8 itself
def test_cube_inverse_against_sampled_inverse(alpha, parts):
    """The wreath inverse at cube irreps (small blocks: F-major adjoint of stage 2, interleaved F;
    large factors: the inverse level recursion factor by factor) against the sampled inverse, an
    independent evaluation of d tr(rho(g)^H f_hat) / |G| that the oracle tests pin."""
    wp = WreathPlan(alpha, parts)
    f = torch.randn((40320, 2187), dtype=torch.float64, device='cuda',
                    generator=torch.Generator(device='cuda').manual_seed(91))
    fh = wp.forward(f)
    fused = wp.inverse(fh)
    rng = random.Random(5)
    sn = list(itertools.islice(itertools.permutations(range(1, 9)), 0, 40320, 4001))
    elements = [(tuple(int(v) for v in wp.oris[rng.randrange(2187)]), rng.choice(sn)) for _ in range(6)]
    want = wp.inverse_at_host(fh, elements)
    kernel = wp.inverse_at(fh, elements)                      # snfft_wreath_inverse_at: one launch
    assert ((kernel - want).abs().max() / want.abs().max()).item() < 1e-12
    lex = {p: i for i, p in enumerate(itertools.permutations(range(1, 9)))}
    ori_index = {tuple(int(v) for v in o): i for i, o in enumerate(wp.oris.tolist())}
    got = torch.stack([fused[lex[s], ori_index[o]] for o, s in elements])
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-12
