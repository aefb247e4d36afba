"""The C-ABI library on a CPU-only box: it loads, exports every symbol the header declares, and
its host-side plan builder (no device) reproduces the reference's tables and, when its block
programs are interpreted in numpy, the reference's level step.  No compute call is made on a
device here."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import snfft_oracle as O
from conftest import rel_fro
from snfft_b200 import Plan
from snfft_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'snfft_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(snfft_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 20
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    for name in declared:
        assert hasattr(N.lib, name)
    assert N.lib.snfft_abi_version() == 2


def test_plan_create_errors():
    h = C.c_void_p()
    assert N.lib.snfft_plan_create(0, -1, C.byref(h)) == 1
    assert N.lib.snfft_plan_create(13, -1, C.byref(h)) == 1
    assert b'1..12' in N.lib.snfft_last_error()
    assert N.lib.snfft_plan_create(4, -1, None) == 1
    with pytest.raises(N.SnfftError):
        Plan(4, 99)                                  # no such device: fails loudly, no fallback


def test_host_only_plan_refuses_device_work():
    plan = Plan(4, -1)
    buf = np.zeros(24)
    rc = N.lib.snfft_forward(plan._h, N.ptr(buf), N.ptr(buf.copy()), N.ptr(buf.copy()), 1, 1, None)
    assert rc == 2 and b'no CPU fallback' in N.lib.snfft_last_error()
    assert N.lib.snfft_inverse(plan._h, N.ptr(buf), N.ptr(buf.copy()), N.ptr(buf.copy()), 1, 1, None) == 2
    assert N.lib.snfft_level(plan._h, 3, N.ptr(buf), N.ptr(buf.copy()), 4, 0, None) == 2
    # the peer-memory exchange of a shard of a host-only plan: same refusal, and argument checks
    shard = C.c_void_p()
    assert N.lib.snfft_shard_create(plan._h, 0, 2, 3, C.byref(shard)) == 0
    peers = (C.c_void_p * 2)(N.ptr(buf), N.ptr(buf))
    assert N.lib.snfft_shard_exchange(shard, peers, N.ptr(buf), 0, None) == 2
    assert N.lib.snfft_shard_exchange(shard, None, N.ptr(buf), 0, None) == 1
    assert N.lib.snfft_shard_exchange(None, peers, N.ptr(buf), 0, None) == 1
    # the whole-schedule entry points, the sampled inverse, the seminormal form and the wreath plan
    # refuse the same way: no device, no result (and no CPU path behind them)
    ws = N.ShardWorkspace()
    ws.local_a = ws.local_b = ws.shard_a = ws.shard_b = ws.status = N.ptr(buf).value
    assert N.lib.snfft_forward_sharded(shard, N.ptr(buf), N.ptr(buf.copy()), C.byref(ws), None) == 2
    assert N.lib.snfft_inverse_sharded(shard, N.ptr(buf), N.ptr(buf.copy()), C.byref(ws), None) == 2
    assert N.lib.snfft_forward_sharded(shard, None, N.ptr(buf), C.byref(ws), None) == 1
    sizes = (C.c_int64 * 4)()
    assert N.lib.snfft_shard_ws_sizes(shard, sizes) == 0
    # 4 blocks of S_3, 2 per rank; rank 1 owns the larger column share (4 of the 6 entries of a spectrum)
    assert sizes[0] == 4 * 4 and sizes[1] == 2 * 6 and sizes[2] == 0 and sizes[3] == 32
    N.lib.snfft_shard_destroy(shard)
    perms = np.array([[1, 2, 3, 4]], dtype=np.int32)
    assert N.lib.snfft_inverse_at(plan._h, N.ptr(buf), N.ptr(perms), 1, N.ptr(buf.copy()), None) == 2
    assert N.lib.snfft_ysemi(plan._h, 1, N.ptr(perms), 1, N.ptr(buf.copy()), None) == 2
    alpha = np.array([1, 2, 1], dtype=np.int32)
    parts = np.zeros((3, 4), dtype=np.int32)
    parts[0, 0], parts[1, 0], parts[2, 0] = 1, 2, 1
    assert N.lib.snfft_wreath_create(4, N.ptr(alpha), N.ptr(parts), 0, -1, C.byref(C.c_void_p())) == 2
    assert b'no CPU fallback' in N.lib.snfft_last_error()
    assert N.lib.snfft_wreath_create(4, N.ptr(np.array([1, 1, 1], dtype=np.int32)), N.ptr(parts), 0, -1,
                                     C.byref(C.c_void_p())) == 1
    assert N.lib.snfft_c3_dft(9, N.ptr(buf), None, N.ptr(perms), 1, 1, N.ptr(buf), N.ptr(buf), 0, None) == 1
    # null plan / null state list: statuses, not crashes
    assert N.lib.snfft_wreath_inverse_at(None, N.ptr(buf), N.ptr(buf), None, None, 0, N.ptr(buf), N.ptr(buf), None) == 1
    assert N.lib.snfft_wreath_inverse(None, N.ptr(buf), N.ptr(buf), N.ptr(buf), None, None) == 1


@pytest.mark.parametrize('n', range(1, 11))
def test_layout_matches_oracle(n):
    plan = Plan(n, -1)
    parts, dims, offs = plan.layout()
    assert parts == O.partitions(n)
    assert sum(d * d for d in dims) == math.factorial(n)
    if n <= 8:
        o_offs, o_dims = O.packed_layout(n)
        assert dims == [o_dims[p] for p in parts] and offs == [o_offs[p] for p in parts]
    for i, p in enumerate(parts):
        idx, off = plan.branch_down(n, i)
        sub = plan.layout(n - 1)[0] if n > 1 else []
        assert [sub[j] for j in idx] == O.branch_down(p)
        assert off == list(np.cumsum([0] + [plan.layout(n - 1)[1][j] for j in idx])[:-1]) or n == 1


def test_hook_length_dimensions_n12():
    plan = Plan(12, -1)
    parts, dims, _ = plan.layout()
    assert len(parts) == 77 and max(dims) == 7700 and sum(dims) == 140152

    def hook(p):
        conj = [sum(1 for r in p if r > c) for c in range(p[0])]
        prod = 1
        for r, length in enumerate(p):
            for c in range(length):
                prod *= (length - c - 1) + (conj[c] - r - 1) + 1
        return math.factorial(sum(p)) // prod
    assert dims == [hook(p) for p in parts]


@pytest.mark.parametrize('n', range(2, 8))
def test_yor_trans_tables_bit_exact(golden_small, n):
    plan = Plan(n, -1)
    for i, p in enumerate(plan.layout()[0]):
        for k in range(1, n):
            ref = golden_small['yortrans_{}_{}'.format('_'.join(map(str, p)), k)]
            assert np.array_equal(plan.yor_trans(n, i, k), ref)
    h = plan._h
    assert N.lib.snfft_yor_trans(h, n, 0, n, N.ptr(np.zeros(1))) == 1     # j out of range


def interpret_level(plan, k, x, inverse=False):
    """Run the exported block programs of level k in numpy: [k, (k-1)!] <-> [k!].
    Forward: rows are read at `rowbase`, every finished output (final mask) lands at `dest`.
    Inverse: rows are read at `dest`, blocks run backwards with the inverse matrices."""
    parts, dims, offs = plan.layout(k - 1)
    kf = math.factorial(k)
    flat_in = x.reshape(-1)
    out = np.zeros(kf)
    for pi, (d, o) in enumerate(zip(dims, offs)):
        prog = plan.export_panel(k, pi)
        assert prog['d'] == d and len(prog['m']) == (k - 1) * d
        assert sorted(set(prog['slots'][prog['slots'] >= 0])) == list(range(k * d))
        assert sorted(prog['rowbase'].tolist()) == sorted(i * (kf // k) + o + u * d for i in range(k) for u in range(d))
        # every slot is finished exactly once
        finals = [prog['slots'][b, a] for b in range(len(prog['m'])) for a in range(prog['m'][b])
                  if (prog['final'][b] >> a) & 1]
        assert sorted(finals) == list(range(k * d))
        src = prog['dest'] if inverse else prog['rowbase']
        dst = prog['rowbase'] if inverse else prog['dest']
        buf = np.stack([flat_in[src[s]:src[s] + d] for s in range(k * d)])
        order = range(len(prog['m']))
        assert list(prog['stage']) == sorted(prog['stage'])
        for b in (reversed(order) if inverse else order):
            m = prog['m'][b]
            sl = prog['slots'][b, :m]
            cm = prog['coef_inv' if inverse else 'coef_fwd'][b, :m * m].reshape(m, m)
            buf[sl] = cm @ buf[sl]
        for s in range(k * d):
            out[dst[s]:dst[s] + d] = buf[s]
    return out if not inverse else out.reshape(k, kf // k)


@pytest.mark.parametrize('n', range(2, 8))
def test_block_programs_reproduce_the_level_step(n):
    plan = Plan(n, -1)
    f = np.random.default_rng(100 + n).standard_normal(math.factorial(n))
    gather = O._coset_gather(n)
    subs = np.stack([O.pack(O.fft_all(f[gather[i]]), n - 1) for i in range(n)])
    ref = O.pack(O.fft_all(f), n)
    got = interpret_level(plan, n, subs)
    for p, a in O.unpack(got, n).items():
        assert rel_fro(a, O.unpack(ref, n)[p]) < 1e-13, p
    back = interpret_level(plan, n, ref, inverse=True)
    assert np.abs(back - subs).max() < 1e-12


def test_whole_transform_from_block_programs_n6():
    n = 6
    plan = Plan(n, -1)
    f = np.random.default_rng(6).standard_normal(720)
    x = f[plan.coset_to_lex()]
    for k in range(2, n + 1):
        groups = x.reshape(-1, k, math.factorial(k - 1))
        x = np.concatenate([interpret_level(plan, k, g) for g in groups])
    assert rel_fro(x, O.pack(O.fft_all(f), n)) < 1e-13


@pytest.mark.parametrize('n', range(1, 9))
def test_coset_to_lex_map(n):
    assert np.array_equal(Plan(n, -1).coset_to_lex(), O.coset_major_to_lex(n))


def test_operation_counts():
    plan = Plan(12, -1)
    macs = [plan.macs_per_element(k) for k in range(2, 13)]
    assert abs(macs[0] - 2.0) < 1e-12 and all(b > a for a, b in zip(macs, macs[1:]))
    assert sum(macs) < 0.75 * 12 * 11          # below Maslen's (3/4) n (n-1) bound
    assert plan.macs_per_element(1) == 0.0 and plan.macs_per_element(13) == 0.0


def test_reorder_tile_factorisation():
    # The index identity behind reorder_tiled_kernel (kernels.cu), restated in Python and checked
    # against the host coset-major -> lexicographic table: with the top R = 5 coset digits fixing
    # the last 5 letters (an arrangement A of a value set V) and the first n-5 letters increasing,
    #   c2l[top(A) + mid * 4! + low] = base(V, mid, low) + A
    n, R, SF = 9, 5, 24
    m = n - R
    fact = [math.factorial(i) for i in range(n + 1)]
    c2l = Plan(n, -1).coset_to_lex()
    rng = np.random.default_rng(0)
    for _ in range(40):
        V = sorted(rng.choice(np.arange(1, n + 1), size=R, replace=False).tolist())
        rest = [x for x in range(1, n + 1) if x not in V]          # iota: x-th value outside V
        A = int(rng.integers(0, fact[R]))
        # arrangement A of V (Lehmer decoding) = letters sigma(m+1..n)
        pool, a, suf = list(V), A, []
        for j in range(R):
            q, a = divmod(a, fact[R - 1 - j])
            suf.append(pool.pop(q))
        # top coset digits: at level k the letter sits at (value - #extracted smaller values)
        top = 0
        for k in range(R - 1, -1, -1):
            pos = suf[k] - sum(1 for t in range(k + 1, R) if suf[t] < suf[k])
            top += (pos - 1) * fact[m + k]
        mid = int(rng.integers(0, fact[m] // SF))
        low = int(rng.integers(0, SF))
        flat = mid * SF + low
        q = list(range(n + 1))                                      # 1-based one-line notation
        for k in range(m, 1, -1):
            i = (flat // fact[k - 1]) % k + 1
            q[i:k + 1] = q[i + 1:k + 1] + [q[i]]
        base = 0
        for a_ in range(1, m + 1):
            smaller = rest[q[a_] - 1] - q[a_] + sum(1 for b in range(a_ + 1, m + 1) if q[b] < q[a_])
            base += smaller * fact[n - a_]
        assert c2l[top + flat] == base + A



def interpret_level_by_phases(plan, k, x, inverse=False):
    """Level k >= 9 the way the device runs it: stages 1..5 in place on the X side (export_panel
    blocks; on the device these are the generated programs plus the finished-row move), then every
    row-group phase from the exported program blobs (what phase_kernel executes): per group load
    its rows, interpret the block stream, store rows in place or - when finished - to the
    spectrum.  x: [k, (k-1)!] forward, [k!] inverse."""
    parts, dims, offs = plan.layout(k - 1)
    kf = math.factorial(k)
    nphases = plan.export_phase(k, 1)['nphases']
    phases = [plan.export_phase(k, ph) for ph in range(1, nphases)]
    p1_stages = phases[0]['stage_lo'] - 1          # stages 1..p1_stages: generated in-place programs
    xs = np.zeros(kf) if inverse else x.reshape(-1).copy()
    ss = x.reshape(-1).copy() if inverse else np.zeros(kf)

    def low_stages(inv):
        for pi, d in enumerate(dims):
            prog = plan.export_panel(k, pi)
            low = [b for b in range(len(prog['m'])) if prog['stage'][b] <= p1_stages]
            last = {}
            for b in range(len(prog['m'])):
                for a in range(prog['m'][b]):
                    last[prog['slots'][b, a]] = b
            done_low = [s for s in range(k * d) if prog['stage'][last[s]] <= p1_stages]
            touched = sorted({int(v) for b in low for v in prog['slots'][b, :prog['m'][b]]})
            buf = {}
            for s in touched:
                src = ss[prog['dest'][s]:prog['dest'][s] + d] if (inv and s in done_low) else \
                    xs[prog['rowbase'][s]:prog['rowbase'][s] + d]
                buf[s] = src.copy()
            for b in (reversed(low) if inv else low):
                m = prog['m'][b]
                sl = [int(v) for v in prog['slots'][b, :m]]
                cm = prog['coef_inv' if inv else 'coef_fwd'][b, :m * m].reshape(m, m)
                res = cm @ np.stack([buf[s] for s in sl])
                for a, s in enumerate(sl):
                    buf[s] = res[a]
            for s in touched:
                if not inv and s in done_low:
                    ss[prog['dest'][s]:prog['dest'][s] + d] = buf[s]
                else:
                    xs[prog['rowbase'][s]:prog['rowbase'][s] + d] = buf[s]

    def run_phase(T, inv):
        blob = T['blob_inv' if inv else 'blob_fwd']
        for pi, (d, tiles, g_off, ng, pack) in enumerate(T['panels']):
            assert d == dims[pi]
            for g in range(g_off, g_off + ng):
                word0, words, nrows, nlow = (int(v) for v in T['groups'][g])
                assert nrows <= T['max_rows'] and words <= T['max_blob']
                raw = blob[16 * word0:16 * (word0 + words)]
                table = raw[:8 * nrows].view(np.uint32)
                xoff, dest = table[:nrows].astype(np.int64), table[nrows:].astype(np.int64)
                # rows [0, nlow) live on the x side, rows [nlow, nrows) are finished in this phase
                buf = np.stack([(ss[dest[r]:dest[r] + d] if (inv and r >= nlow) else xs[xoff[r]:xoff[r] + d]).copy()
                                for r in range(nrows)])
                at = 16 * ((2 * nrows + 3) // 4)
                while True:
                    hdr = raw[at:at + 16].view(np.uint16)
                    m = int(hdr[5])
                    if m == 0:
                        break
                    assert 2 <= m <= 5 and all(int(v) % 256 == 0 for v in hdr[:m])
                    loc = [int(v) // 256 for v in hdr[:m]]
                    assert len(set(loc)) == m and max(loc) < nrows
                    cm = raw[at + 16:at + 16 + 8 * m * m].view(np.float64).reshape(m, m)
                    buf[loc] = cm @ buf[loc]
                    at += 16 + 16 * ((m * m + 1) // 2)
                assert at + 16 == 16 * words
                for r in range(nrows):
                    if not inv and r >= nlow:
                        ss[dest[r]:dest[r] + d] = buf[r]
                    else:
                        xs[xoff[r]:xoff[r] + d] = buf[r]

    if not inverse:
        low_stages(False)
        for T in phases:
            run_phase(T, False)
        return ss
    for T in reversed(phases):
        run_phase(T, True)
    low_stages(True)
    return xs.reshape(k, kf // k)


@pytest.mark.parametrize('windows', [None, '9:7,8'])
def test_phase_bundle_tables_reproduce_level_9(windows, monkeypatch):
    """The bundle tables of the row-group phases (device layout) against the plain block programs."""
    if windows:
        monkeypatch.setenv('SNFFT_WINDOWS', windows)
    k = 9
    plan = Plan(k, -1)
    x = np.random.default_rng(9).standard_normal((k, math.factorial(k - 1)))
    ref = interpret_level(plan, k, x)
    got = interpret_level_by_phases(plan, k, x)
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()
    back = interpret_level_by_phases(plan, k, ref, inverse=True)
    assert np.abs(back - x).max() < 1e-11
