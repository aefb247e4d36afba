"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden
vectors.  Tolerance (BASELINE.json north_star): relative Frobenius error <= 1e-10 per irrep,
fp64.  Sizes the oracle cannot reach are covered by size-independent properties: round trip,
Plancherel, translation covariance, sparse-support inputs."""
import math

import numpy as np
import pytest
import torch

import snfft_oracle as O
from conftest import rel_fro
from snfft_b200 import get_plan, launch_count

pytestmark = pytest.mark.gpu
TOL = 1e-10   # relative Frobenius per irrep (north_star)


def per_irrep_err(plan, got, ref):
    """max over irreps (and batch rows) of the relative Frobenius error."""
    parts, dims, offs = plan.layout()
    got = np.atleast_2d(got)
    ref = np.atleast_2d(ref)
    worst = 0.0
    for d, o in zip(dims, offs):
        for b in range(got.shape[0]):
            worst = max(worst, rel_fro(got[b, o:o + d * d], ref[b, o:o + d * d]))
    return worst


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize('n', range(1, 9))
def test_forward_matches_oracle(n):
    plan = get_plan(n)
    f = np.random.default_rng(n).standard_normal((3, math.factorial(n)))
    ref = O.pack(O.fft_all(f), n)
    before = launch_count()
    got = plan.forward(dev(f), 'lex').cpu().numpy()
    assert launch_count() > before or n == 1            # our kernels ran (n = 1 is a device copy)
    assert per_irrep_err(plan, got, ref) < TOL
    # coset-major input order gives the same spectrum
    got2 = plan.forward(dev(f[:, O.coset_major_to_lex(n)]), 'coset').cpu().numpy()
    assert np.array_equal(got, got2)


@pytest.mark.parametrize('n', range(2, 8))
def test_forward_matches_reference_golden(golden_small, n):
    plan = get_plan(n)
    f = np.random.default_rng(n).standard_normal(math.factorial(n))
    got = plan.forward(dev(f), 'lex').cpu().numpy()
    assert per_irrep_err(plan, got, golden_small[f'direct_{n}']) < TOL
    if n <= 6:
        assert per_irrep_err(plan, got, golden_small[f'clausen_{n}']) < TOL


def test_config1_s5_single_function(golden_small):
    # BASELINE configs[0]: S_5, one random function (seed 5), all 7 irreps vs the direct sum
    plan = get_plan(5)
    f = np.random.default_rng(5).standard_normal(120)
    got = plan.unpack(plan.forward(dev(f), 'lex'))
    ref = O.unpack(golden_small['direct_5'], 5)
    assert list(got) == O.partitions(5)
    assert [v.shape[-1] for v in got.values()] == [1, 4, 6, 4, 1, 5, 5]
    for p in got:
        assert rel_fro(got[p].cpu().numpy(), ref[p]) < TOL


@pytest.mark.parametrize('n', range(1, 9))
def test_inverse_matches_oracle_and_round_trip(n):
    plan = get_plan(n)
    f = np.random.default_rng(50 + n).standard_normal((2, math.factorial(n)))
    spec = O.pack(O.fft_all(f), n)
    back = plan.inverse(dev(spec), 'lex').cpu().numpy()
    assert rel_fro(back, f) < TOL
    if n <= 5:
        direct = np.stack([O.inverse_direct(O.unpack(s, n), n) for s in spec])
        assert rel_fro(back, direct) < TOL
    x = dev(f)
    assert rel_fro(plan.inverse(plan.forward(x, 'coset'), 'coset').cpu().numpy(), f) < TOL


@pytest.fixture
def table_driven():
    """Route level 8 through the table-driven streaming kernel (the one that serves k >= 9)."""
    from snfft_b200 import _native as N
    N.lib.snfft_debug_force_table_driven(1)
    yield
    N.lib.snfft_debug_force_table_driven(0)


@pytest.mark.parametrize('n', range(2, 9))
def test_streaming_level_kernel_every_level(n, table_driven):
    """snfft_level (the table-driven streaming kernel) applied level by level equals the fused path."""
    plan = get_plan(n)
    f = np.random.default_rng(7 * n).standard_normal((2, math.factorial(n)))
    ref = O.pack(O.fft_all(f), n)
    x = dev(f[:, O.coset_major_to_lex(n)])
    for k in range(2, n + 1):
        x = plan.level(k, x)
    assert per_irrep_err(plan, x.cpu().numpy(), ref) < TOL
    y = dev(ref)
    for k in range(n, 1, -1):
        y = plan.level(k, y, inverse=True)
    assert rel_fro(y.cpu().numpy(), f[:, O.coset_major_to_lex(n)]) < TOL


def test_config2_s8_batch_1024(golden_s8):
    # BASELINE configs[1]: S_8 forward + inverse, 1024 random functions (seed 8)
    n, B = 8, 1024
    plan = get_plan(n)
    f = np.random.default_rng(8).standard_normal((B, math.factorial(n)))
    x = dev(f)
    spec = plan.forward(x, 'lex')
    rows = golden_s8['rows']
    got = spec[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert per_irrep_err(plan, got, golden_s8['direct_8']) < TOL     # reference fourier_transform2
    back = plan.inverse(spec, 'lex')
    err = (back - x).norm(dim=1) / x.norm(dim=1)
    assert err.max().item() < TOL
    # Plancherel: sum_lambda d_lambda ||fhat(lambda)||_F^2 = n! ||f||^2
    parts, dims, offs = plan.layout()
    lhs = sum(d * spec[:, o:o + d * d].pow(2).sum(dim=1) for d, o in zip(dims, offs))
    rhs = math.factorial(n) * x.pow(2).sum(dim=1)
    assert ((lhs - rhs).abs() / rhs).max().item() < 1e-12


def test_edge_cases():
    plan = get_plan(4)
    empty = torch.empty((0, 24), dtype=torch.float64, device='cuda')
    assert plan.forward(empty).shape == (0, 24)
    assert plan.inverse(empty).shape == (0, 24)
    one = torch.zeros(24, dtype=torch.float64, device='cuda')
    one[0] = 1.0                                       # delta at the identity: fhat = identity
    for p, m in plan.unpack(plan.forward(one, 'lex')).items():
        assert torch.allclose(m, torch.eye(m.shape[-1], dtype=torch.float64, device='cuda'))
    const = torch.ones(24, dtype=torch.float64, device='cuda')
    spec = plan.unpack(plan.forward(const, 'lex'))
    assert abs(spec[(4,)].item() - 24.0) < 1e-12       # trivial irrep sums f
    assert all(m.abs().max().item() < 1e-12 for p, m in spec.items() if p != (4,))
    with pytest.raises(TypeError):
        plan.forward(torch.zeros(24, dtype=torch.float32, device='cuda'))
    with pytest.raises(ValueError):
        plan.forward(torch.zeros(25, dtype=torch.float64, device='cuda'))
    with pytest.raises(TypeError):
        plan.forward(torch.zeros(24, dtype=torch.float64))          # host tensor: no CPU fallback
    x = torch.randn(5, 24, dtype=torch.float64, device='cuda')
    from snfft_b200 import _native as N
    assert N.lib.snfft_forward(plan._h, N.ptr(x), N.ptr(x), N.ptr(torch.empty_like(x)), 5, 1, None) == 1


def test_linearity_and_ragged_batches():
    plan = get_plan(7)
    for B in (1, 2, 5, 149, 300):
        x = torch.randn(B, 5040, dtype=torch.float64, device='cuda')
        y = torch.randn(B, 5040, dtype=torch.float64, device='cuda')
        lhs = plan.forward(2.5 * x - y, 'lex')
        rhs = 2.5 * plan.forward(x, 'lex') - plan.forward(y, 'lex')
        assert ((lhs - rhs).norm() / rhs.norm()).item() < 1e-13


def test_reorder_kernels():
    for n in (1, 2, 5, 8):
        plan = get_plan(n)
        c2l = O.coset_major_to_lex(n)
        x = np.random.default_rng(1).standard_normal((3, math.factorial(n)))
        to_coset = plan.reorder(dev(x), True).cpu().numpy()
        assert np.array_equal(to_coset, x[:, c2l])
        assert np.array_equal(plan.reorder(dev(to_coset), False).cpu().numpy(), x)


@pytest.mark.parametrize('n', [9, 10, 11])
def test_reorder_tiled_kernel(n):
    # n >= 9 uses the tiled transpose kernel; the host index table is the checker
    plan = get_plan(n)
    c2l = plan.coset_to_lex()
    size = math.factorial(n)
    batch = 3 if n < 11 else 1
    x = torch.arange(batch * size, dtype=torch.float64, device='cuda').reshape(batch, size)
    to_coset = plan.reorder(x, True)
    want = torch.from_numpy(c2l.astype(np.float64)).cuda()
    for b in range(batch):
        assert torch.equal(to_coset[b], want + float(b * size))
    assert torch.equal(plan.reorder(to_coset, False), x)
    del x, to_coset, want
    torch.cuda.empty_cache()


def test_reorder_tiled_matches_elementwise_s12():
    from snfft_b200 import _native as N
    plan = get_plan(12)
    x = torch.arange(plan.size, dtype=torch.float64, device='cuda').reshape(1, -1)
    tiled = plan.reorder(x, True)
    N.lib.snfft_debug_force_table_driven(1)
    try:
        plain = plan.reorder(x, True)
    finally:
        N.lib.snfft_debug_force_table_driven(0)
    assert torch.equal(tiled, plain)
    del plain
    assert torch.equal(plan.reorder(tiled, False), x)
    del x, tiled
    torch.cuda.empty_cache()


def test_yor_device_matches_reference_golden(golden_small):
    names = [k for k in golden_small.files if k.startswith('yor_')]
    for name in names:
        nums = list(map(int, name.split('_')[1:]))
        for cut in range(1, len(nums)):
            if sum(nums[:cut]) == len(nums) - cut:
                p, tup = tuple(nums[:cut]), tuple(nums[cut:])
                break
        plan = get_plan(len(tup))
        perms = torch.tensor([tup], dtype=torch.int32, device='cuda')
        got = plan.yor(plan.irrep_index(p), perms)[0].cpu().numpy()
        assert rel_fro(got, golden_small[name]) < 1e-13


def test_yor_homomorphism_n9():
    # test.py:87-101: rho(g) rho(h) = rho(gh) for every shape of 9
    n = 9
    plan = get_plan(n)
    rng = np.random.default_rng(9)
    g = [tuple(int(x) + 1 for x in rng.permutation(n)) for _ in range(4)]
    h = [tuple(int(x) + 1 for x in rng.permutation(n)) for _ in range(4)]
    gh = [O.perm_mul(a, b) for a, b in zip(g, h)]
    perms = torch.tensor(g + h + gh, dtype=torch.int32, device='cuda')
    for i in range(len(plan.layout()[0])):
        m = plan.yor(i, perms)
        assert torch.allclose(torch.bmm(m[0:4], m[4:8]), m[8:12], atol=1e-12)


@pytest.mark.parametrize('n', [3, 5, 6, 7])
def test_direct_sum_gemm_matches_oracle(golden_small, n):
    plan = get_plan(n)
    f = np.random.default_rng(n).standard_normal(math.factorial(n))
    extra = np.random.default_rng(99).standard_normal((70, math.factorial(n)))
    got = plan.dft_direct(dev(np.vstack([f[None], extra]))).cpu().numpy()
    assert per_irrep_err(plan, got[:1], golden_small[f'direct_{n}'][None]) < TOL
    fast = plan.forward(dev(extra), 'lex').cpu().numpy()
    assert per_irrep_err(plan, got[1:], fast) < TOL


def test_direct_sum_gemm_s8_against_reference_golden(golden_s8):
    """snfft_dft_direct at n = 8 (the [8!, 8!] matrix of all rho_lambda(sigma) is 13 GB, built on the
    device) against the reference's own fourier_transform2 of functions 0, 511, 1023 of the seed-8
    batch (tests/golden/reference_s8.npz)."""
    n = 8
    plan = get_plan(n)
    f = np.random.default_rng(8).standard_normal((1024, math.factorial(n)))
    rows = golden_s8['rows']
    try:
        got = plan.dft_direct(dev(np.ascontiguousarray(f[rows]))).cpu().numpy()
    finally:
        plan.dft_release()
    assert per_irrep_err(plan, got, golden_s8['direct_8']) < TOL
    assert plan.device_bytes() < 1 << 30          # the matrix is gone again


def test_misaligned_buffers_are_rejected():
    """The fused kernel stages its input with 16-byte cp.async: a float64 view at an odd element
    offset must be refused with SNFFT_ERR_INVALID instead of faulting."""
    from snfft_b200 import _native as N
    n = 7
    plan = get_plan(n)
    nf = math.factorial(n)
    base = torch.zeros(2 * nf + 2, dtype=torch.float64, device='cuda')
    out, work = torch.empty(nf, dtype=torch.float64, device='cuda'), torch.empty(nf, dtype=torch.float64, device='cuda')
    odd = base[1:1 + nf]
    assert odd.data_ptr() % 16 == 8
    with pytest.raises(N.SnfftError):
        plan.forward(odd, 'coset', out=out, work=work)
    with pytest.raises(N.SnfftError):
        plan.inverse(odd, 'coset', out=out, work=work)
    torch.cuda.synchronize()                      # no sticky error: the context is still usable
    assert plan.forward(base[:nf], 'coset', out=out, work=work).abs().max().item() == 0.0


def test_largest_irreps_columns_against_oracle_s10_s12():
    """n = 10 and 12: columns of the LARGEST irreps (d = 768 / d = 7700, 5632, 5775: the 48-row
    groups of the phase kernel) for a sparse-support input against columns of rho_lambda(sigma)
    built by the oracle's restatement of yor.py - not by the device's own yor kernel."""
    import parity_tools as PT
    for n in (10, 12):
        plan = get_plan(n)
        parts, dims, offs = plan.layout()
        perms, coefs = PT.sparse_support(n, count=3, seed=n)
        x = torch.zeros(math.factorial(n), dtype=torch.float64, device='cuda')
        for p_, c_ in zip(perms, coefs):
            x[PT.O.coset_index(p_)] += float(c_)
        spec = plan.forward(x, 'coset')
        del x

        def get_columns(i, cols):
            d = dims[i]
            return spec[offs[i]:offs[i] + d * d].reshape(d, d)[:, cols].cpu().numpy()
        irreps = PT.pick_irreps(parts, dims, big=4, small=4)
        worst, checked = PT.check_columns(get_columns, parts, dims, perms, coefs, irreps)
        assert worst < TOL, checked
        assert max(d for _, d, _, _ in checked) == max(dims)
        del spec
        torch.cuda.empty_cache()


@pytest.mark.parametrize('n', [9, 10])
def test_large_n_properties(n):
    """n = 9, 10: the reference cannot run; the oracle port runs n = 9 only on a sample.  Check
    round trip, Plancherel, sparse-support inputs and translation covariance."""
    plan = get_plan(n)
    nf = math.factorial(n)
    gen = torch.Generator(device='cuda').manual_seed(n)
    x = torch.randn(2, nf, dtype=torch.float64, device='cuda', generator=gen)
    spec = plan.forward(x, 'lex')
    back = plan.inverse(spec, 'lex')
    assert ((back - x).norm() / x.norm()).item() < TOL
    parts, dims, offs = plan.layout()
    lhs = sum(d * spec[:, o:o + d * d].pow(2).sum(dim=1) for d, o in zip(dims, offs))
    rhs = nf * x.pow(2).sum(dim=1)
    assert ((lhs - rhs).abs() / rhs).max().item() < 1e-12
    # sparse support: f = sum_j c_j delta_{sigma_j}  =>  fhat(lambda) = sum_j c_j rho_lambda(sigma_j)
    rng = np.random.default_rng(n)
    sig = [tuple(int(v) + 1 for v in rng.permutation(n)) for _ in range(3)]
    coef = rng.standard_normal(3)
    sparse = torch.zeros(nf, dtype=torch.float64, device='cuda')
    for s, c in zip(sig, coef):
        sparse[O.lex_rank(s)] += c
    sp = plan.unpack(plan.forward(sparse, 'lex'))
    perms = torch.tensor(sig, dtype=torch.int32, device='cuda')
    cdev = torch.from_numpy(coef).cuda()
    for i, p in enumerate(parts):
        want = (cdev[:, None, None] * plan.yor(i, perms)).sum(dim=0)
        assert ((sp[p] - want).norm() / want.norm()).item() < TOL, p
        if dims[i] <= 48:   # the device rho against the oracle's rho (restated yor.py)
            assert rel_fro(plan.yor(i, perms[:1])[0].cpu().numpy(), O.yor(p, sig[0])) < 1e-12
    # translation covariance: (L_tau f)(sigma) = f(tau^-1 sigma)  =>  hat = rho(tau) fhat
    tau = sig[0]
    tau_inv = O.perm_inv(tau)
    support = [tuple(int(v) + 1 for v in rng.permutation(n)) for _ in range(5)]
    vals = rng.standard_normal(5)
    f0 = torch.zeros(nf, dtype=torch.float64, device='cuda')
    f1 = torch.zeros(nf, dtype=torch.float64, device='cuda')
    for s, v in zip(support, vals):
        f0[O.lex_rank(s)] += v
        f1[O.lex_rank(O.perm_mul(tau, s))] += v          # f1(tau s) = f0(s)
    s0 = plan.unpack(plan.forward(f0, 'lex'))
    s1 = plan.unpack(plan.forward(f1, 'lex'))
    for i, p in enumerate(parts):
        rho = plan.yor(i, perms[:1])[0]
        assert ((s1[p] - rho @ s0[p]).norm() / s1[p].norm()).item() < TOL


def test_s9_against_oracle_port():
    """n = 9 full transform against the numpy oracle (reference infeasible; oracle port runs in
    a few seconds thanks to shared sub-transforms)."""
    n = 9
    plan = get_plan(n)
    f = np.random.default_rng(9).standard_normal(math.factorial(n))
    ref = O.pack(O.fft_all(f), n)
    got = plan.forward(dev(f), 'lex').cpu().numpy()
    assert per_irrep_err(plan, got, ref) < TOL


def test_host_pipeline_matches_device_path():
    n, B = 7, 37
    plan = get_plan(n)
    f = torch.randn(B, 5040, dtype=torch.float64).pin_memory()
    spec = torch.empty_like(f).pin_memory()
    back = torch.empty_like(f).pin_memory()
    plan.roundtrip_host(f, spec, back, order='lex', chunks=5, streams=3)
    torch.cuda.synchronize()
    want = plan.forward(f.cuda(), 'lex').cpu()
    assert torch.equal(spec, want)
    assert ((back - f).norm() / f.norm()).item() < TOL


@pytest.mark.parametrize('n', [11, 12])
def test_s11_s12_properties(n):
    """BASELINE configs[4]: S_12 (479 M elements, 3.8 GB per buffer) on one GPU; the reference and
    the oracle cannot run here, so: round trip, Plancherel, and a sparse-support input checked per
    irrep against rho_lambda(sigma) from the device YOR kernel (irreps up to dimension 320)."""
    plan = get_plan(n)
    nf = math.factorial(n)
    gen = torch.Generator(device='cuda').manual_seed(n)
    x = torch.randn(nf, dtype=torch.float64, device='cuda', generator=gen)
    spec = plan.forward(x, 'coset')
    back = plan.inverse(spec, 'coset')
    assert ((back - x).norm() / x.norm()).item() < TOL
    del back
    parts, dims, offs = plan.layout()
    lhs = sum(d * spec[o:o + d * d].pow(2).sum().item() for d, o in zip(dims, offs))
    rhs = nf * x.pow(2).sum().item()
    assert abs(lhs - rhs) / rhs < 1e-12
    del x, spec
    torch.cuda.empty_cache()
    rng = np.random.default_rng(n)
    sig = [tuple(int(v) + 1 for v in rng.permutation(n)) for _ in range(2)]
    coef = rng.standard_normal(2)
    sparse = torch.zeros(nf, dtype=torch.float64, device='cuda')
    for s, c in zip(sig, coef):
        sparse[O.lex_rank(s)] += c
    sp = plan.forward(sparse, 'lex')
    del sparse
    perms = torch.tensor(sig, dtype=torch.int32, device='cuda')
    cdev = torch.from_numpy(coef).cuda()
    checked = 0
    for i, (p, d, o) in enumerate(zip(parts, dims, offs)):
        if d > 320:
            continue
        want = (cdev[:, None, None] * plan.yor(i, perms)).sum(dim=0)
        got = sp[o:o + d * d].reshape(d, d)
        assert ((got - want).norm() / want.norm()).item() < TOL, p
        checked += 1
    assert checked >= 10
