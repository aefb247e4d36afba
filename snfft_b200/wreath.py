"""Wreath-product (C_3 wr S_n, 2x2 cube) representations and their Fourier transform, with the
reference's names (reference ``wreath.py``, ``multi.py``, ``cube_ft.py``, ``cube_ift.py``).

rho(o, sigma) = D(o) P(sigma): P = Ind_{S_alpha}^{S_n}(kron_k rho_{parts_k}) as an N x N block
matrix over the left cosets t_i S_alpha (block (i, j) present iff t_i^-1 sigma t_j in S_alpha,
wreath.py:281-315) and D(o) the block-scalar matrix of cyclic characters (wreath.py:43-58,
:123-143).  The transform  f_hat = sum_g f(g) rho(g)  (multi.py:25-46, cube_ft.py:53-61) factors
into (1) character sums over the orientations, F_i(sigma) = sum_o f(o, sigma) chi_i(o), and (2) for
every coset pair a sum over the Young subgroup, block(i, j) = sum_h F_i(t_i h t_j^-1) rho_parts(h)
(SURVEY.md 7.0(10)).  For the standard C_3 orientation layouts the whole transform lives behind the
C ABI (``snfft_wreath_create / _forward / _inverse / _inverse_at``): (1) is a DFT over C_3^m in
shared memory, (2) a fused gather kernel for small blocks or the S_a level recursion of the
Young-subgroup factors.  Other groups (C_2 wr S_n, arbitrary orientation lists) run the same two
stages from this module as dense fp64 products on the DMMA tensor pipe (``snfft_dgemm``), regrouped
by one index gather (``snfft_gather``).  All arithmetic is complex128 (the reference uses complex64
scalars).  ``WreathPlan`` needs a CUDA device; the dictionary-based helpers (``wreath_yor``,
``get_mat``, ``wreath_rep``) are host-side glue like the reference's.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
from functools import reduce

import numpy as np

from . import _native as N
from . import perm2
from .coset_utils import (coset_reps, tup_set, young_coset_reps, young_subgroup, young_subgroup_perm)
from .plan import get_plan
from .utils import partitions
from .yor import yor
from .young_tableau import FerrersDiagram, wreath_dim


# ---- group elements (wreath.py:14-121) -----------------------------------------------------------
def dot(perm, cyc):
    p_inv = perm.inv()
    return CyclicGroup(tuple(cyc[p_inv[i] - 1] for i in range(1, len(cyc) + 1)), cyc.order)


def dot_tup(perm, tup):
    p_inv = perm.inv()
    return tuple(tup[p_inv[i] - 1] for i in range(1, len(tup) + 1))


def dot_tup_inv(perm, tup):
    return tuple(tup[perm[i] - 1] for i in range(1, len(tup) + 1))


class CyclicGroup:
    def __init__(self, cyc, order):
        self.cyc, self.size, self.order = tuple(cyc), len(cyc), order

    @property
    def tup_rep(self):
        return self.cyc

    def inv(self):
        return CyclicGroup(tuple((self.order - a) % self.order for a in self.cyc), self.order)

    def __mul__(self, other):
        return CyclicGroup(tuple((a + b) % self.order for a, b in zip(self.cyc, other.cyc)), self.order)

    __add__ = __mul__

    def __len__(self):
        return len(self.cyc)

    def __getitem__(self, i):
        return self.cyc[i]

    def __repr__(self):
        return '{}/{}'.format(self.cyc, self.order)


class WreathCycSn:
    """(o, sigma) with (o, s)(o', s') = (o + s.o', s s') (wreath.py:92-121)."""

    def __init__(self, cyc, perm):
        assert perm.size == len(cyc)
        self.cyc, self.perm = cyc, perm

    @property
    def tup_rep(self):
        return self.cyc.tup_rep, self.perm.tup_rep

    @staticmethod
    def from_tup(cyc_tup, perm_tup, order):
        return WreathCycSn(CyclicGroup(cyc_tup, order), perm2.Perm2.from_tup(perm_tup))

    def __mul__(self, w):
        return WreathCycSn(self.cyc + dot(self.perm, w.cyc), self.perm * w.perm)

    def inv(self):
        p_inv = self.perm.inv()
        return WreathCycSn(dot(p_inv, self.cyc.inv()), p_inv)

    def inv_tup_rep(self):
        return self.inv().tup_rep

    def __repr__(self):
        return '{} | {}'.format(self.cyc, self.perm)


# ---- characters and induced blocks (wreath.py:43-58, :123-143, :175-224, :281-343, :387-394) -------
def cyclic_irreps(weak_partition):
    """tup -> exp(2 pi i (1*p2 + 2*p3)/3), p_k = sum of the k-th alpha-segment of tup
    (wreath.py:123-143).  A weak partition into two parts gives the C_2 characters (+-1) of the
    pyraminx group C2 wr S6 (pyraminx/px_wreath.py:17-23, ``cyc_irr_func``)."""
    m = len(weak_partition)
    seg = np.repeat(np.arange(m), weak_partition)

    def func(tup):
        expo = int(np.dot(seg, np.asarray(tup)))
        if m == 2:
            return 1.0 if expo % 2 == 0 else -1.0
        return np.exp(2j * np.pi * expo / float(m))
    return func


cyc_irr_func = cyclic_irreps   # the name pyraminx/px_wreath.py uses


def block_cyclic_irreps(tup, coset_reps, cyclic_irrep_func):
    """One scalar per coset representative (complex128; the reference stores complex64)."""
    return np.array([cyclic_irrep_func(dot_tup_inv(rep, tup)) for rep in coset_reps], dtype=np.complex128)


def canonical_order(tup_rep):
    out, shift = [], 0
    for lvl, tup in enumerate(tup_rep):
        out.append(tup if lvl == 0 else tuple(shift + i for i in tup))
        shift += len(tup)
    return tuple(out)


def young_subgroup_yor(alpha, _parts, prefix=None):
    """{h in S_alpha as a permutation of 1..n: kron of the factors' YOR matrices}."""
    nz = [(a, tuple(p)) for a, p in zip(alpha, _parts) if a > 0]
    tables = []
    for a, p in nz:
        fd = FerrersDiagram.from_partition(p)
        tables.append({t: yor(fd, perm2.Perm2.from_tup(t)) for t in itertools.permutations(range(1, a + 1))})
    out = {}
    for g in young_subgroup(alpha):
        key = tuple(i for t in canonical_order(g) for i in t)
        out[key] = reduce(np.kron, [tab[h] for tab, h in zip(tables, g)])
    return out


def wreath_yor(alpha, _parts, prefix=None):
    """{sigma: {(i, j): block}} for every sigma in S_n (wreath.py:281-315).  Instead of scanning all
    coset pairs per sigma, column j is found from the coset of sigma t_j."""
    n = sum(alpha)
    young_yor = young_subgroup_yor(alpha, _parts)
    reps = [perm2.Perm2.from_tup(t) for t in young_coset_reps(alpha)]
    index = {r.tup_rep: i for i, r in enumerate(reps)}
    rep_dict = {}
    for g in perm2.sn(n):
        g_rep = {}
        for j, t_j in enumerate(reps):
            i = index[_coset_rep((g * t_j).tup_rep, alpha)]
            g_rep[(i, j)] = young_yor[(reps[i].inv() * g * t_j).tup_rep]
        rep_dict[g.tup_rep] = g_rep
    return rep_dict


def _coset_rep(tup, alpha):
    """Representative of tup S_alpha: values sorted inside every alpha-segment."""
    out, start = [], 0
    for p in alpha:
        out.extend(sorted(tup[start:start + p]))
        start += p
    return tuple(out)


def get_mat(g, yor_dict, block_scalars=None):
    yg = yor_dict[g if isinstance(g, tuple) else g.tup_rep]
    block = next(iter(yg.values())).shape[0]
    size = len(yg) * block
    mat = np.zeros((size, size), dtype=np.complex128)
    for (i, j), v in yg.items():
        mat[block * i:block * (i + 1), block * j:block * (j + 1)] = (1 if block_scalars is None else block_scalars[i]) * v
    return mat


def wreath_rep(cyc_tup, perm, yor_dict, cos_reps, cyc_irrep_func=None, alpha=None):
    if cyc_irrep_func is None and alpha is None:
        raise ValueError('Need to supply either irrep func or alpha')
    if cyc_irrep_func is None:
        cyc_irrep_func = cyclic_irreps(alpha)
    return get_mat(perm, yor_dict, block_cyclic_irreps(cyc_tup, cos_reps, cyc_irrep_func))


# ---- the transform on the device ---------------------------------------------------------------------
def _lex_ranks(perms):
    """Lexicographic ranks of an int array of permutations [..., n] (1-based letters)."""
    n = perms.shape[-1]
    rank = np.zeros(perms.shape[:-1], dtype=np.int64)
    for a in range(n):
        smaller = (perms[..., a + 1:] < perms[..., a:a + 1]).sum(axis=-1)
        rank += smaller * math.factorial(n - 1 - a)
    return rank


def orientation_mode(oris, n):
    """How the orientation axis of a C_3 group is laid out: 'zero_sum' = the tuples of {0,1,2}^n with sum = 0 mod 3
    in itertools.product order (utils.py:206-209: the index is the base-3 number of the first n-1
    letters), 'full' = all 3^n tuples in product order, None = anything else."""
    oris = np.asarray(oris, dtype=np.int64).reshape(-1, n)
    pw = 3 ** np.arange(n - 1, -1, -1)
    if len(oris) == 3 ** n and np.array_equal(oris @ pw, np.arange(3 ** n)):
        return 'full'
    if len(oris) == 3 ** (n - 1) and np.array_equal(oris[:, :n - 1] @ pw[1:], np.arange(3 ** (n - 1))) \
            and not (oris.sum(axis=1) % 3).any():
        return 'zero_sum'
    return None


def dft_frequencies(alpha, reps, mode):
    """Index of chi_i on the frequency grid of the C_3^m DFT (``snfft_c3_dft``) for every coset
    representative t_i.  chi_i(o) = omega^{sum_q e_i(q) o_q} with e_i(t_i(p)) = k for p in the k-th
    alpha-segment (wreath.py:34-58, :123-143); on the zero-sum subgroup o_n = -(o_1 + .. + o_{n-1}),
    so chi_i is the character (e_i(q) - e_i(n)) mod 3 of C_3^{n-1}."""
    n = sum(alpha)
    T = np.array(reps, dtype=np.int64)
    seg = np.repeat(np.arange(3), alpha)                      # segment of position p
    e = np.empty_like(T)
    np.put_along_axis(e, T - 1, np.broadcast_to(seg, T.shape), axis=1)   # e[i, t_i(p) - 1] = seg(p)
    if mode == 'full':
        return (e @ (3 ** np.arange(n - 1, -1, -1))).astype(np.int32)
    k = (e[:, :n - 1] - e[:, n - 1:]) % 3
    return (k @ (3 ** np.arange(n - 2, -1, -1))).astype(np.int32)


_ALPHA_TABLES = {}


def _alpha_tables(alpha, device):
    """Tables that depend on alpha only (shared by all `parts`): coset representatives and the
    gather index sigma = t_i h t_j^-1 for every (i, j, h)."""
    import torch
    key = (tuple(alpha), device)
    if key not in _ALPHA_TABLES:
        reps = young_coset_reps(alpha)
        N_ = len(reps)
        T = np.array(reps, dtype=np.int64)
        H = np.array([p.tup_rep for p in young_subgroup_perm(alpha)], dtype=np.int64)   # [nH, n]
        Tinv = np.argsort(T, axis=1) + 1
        hT = H[:, Tinv - 1]                                           # h(t_j^-1(p)) : [nH, N, n]
        idx = np.empty((N_, N_, len(H)), dtype=np.int64)
        for i in range(N_):
            sig = T[i][hT - 1]                                        # t_i(h(t_j^-1(p)))
            idx[i] = _lex_ranks(sig).T * N_ + i
        flat = idx.reshape(-1)
        inv = np.empty_like(flat)
        inv[flat] = np.arange(flat.size)
        dev = f'cuda:{device}'
        _ALPHA_TABLES[key] = dict(reps=reps, T=T, nH=len(H), idx=torch.from_numpy(flat).to(dev),
                                  inv_idx=torch.from_numpy(inv).to(dev))
        if len(_ALPHA_TABLES) > 4:                                   # the big ones hold 360 MB of indices
            _ALPHA_TABLES.pop(next(iter(_ALPHA_TABLES)))
    return _ALPHA_TABLES[key]


class WreathPlan:
    """Fourier transform of functions on C_m^n x S_n (orientations restricted to ``oris``) at one
    irrep (alpha, parts); m = len(alpha): 3 for the 2x2 cube group C3 wr S8 (wreath.py), 2 for the
    pyraminx group C2 wr S6 (pyraminx/px_wreath.py).  f is a float64 CUDA tensor [n!, len(oris)]:
    sigma in ``sn(n)`` order (slow axis), orientation index (fast axis).

    Stage 1 (character sums over the orientations) is a radix-3 DFT over C_3^m in shared memory
    (``snfft_c3_dft``: one pass over f) whenever the orientation axis is the standard one
    (``orientation_mode``); an arbitrary orientation list falls back to the dense character GEMM.
    Stage 2 (sum over the Young subgroup per coset pair): behind the C plan a fused gather kernel for
    small blocks, else the S_a level recursion of the Young-subgroup factors; one gather plus one DMMA
    GEMM in this class's own fallback for general C_m."""

    def __init__(self, alpha, parts, oris=None, device: int = 0):
        import torch
        self.alpha, self.parts, self.device = tuple(alpha), tuple(tuple(p) for p in parts), device
        self.n = n = sum(alpha)
        dev = f'cuda:{device}'
        self.reps = young_coset_reps(self.alpha)
        self.N = len(self.reps)
        self.block = wreath_dim(self.parts)
        self.D = self.N * self.block
        self.cyc_order = m = len(self.alpha)
        if oris is None:
            if m == 3:
                oris = [t for t in itertools.product((0, 1, 2), repeat=n) if sum(t) % 3 == 0]   # utils.py:206-209
            else:
                oris = list(itertools.product(range(m), repeat=n))
        self.oris = np.array(oris, dtype=np.int64).reshape(-1, n)
        self.order = math.factorial(n) * len(self.oris)
        self.mode = orientation_mode(self.oris, n) if m == 3 else None
        self.seg = np.repeat(np.arange(m), self.alpha)               # segment of position p
        self._c = None
        if self.mode is not None and (n if self.mode == 'full' else n - 1) <= 8:
            # the standard C_3 layouts: the whole transform lives behind the C ABI (tables built in
            # C++ / on the device); this class is a thin caller
            a = np.array(self.alpha, dtype=np.int32)
            pz = np.zeros((3, n), dtype=np.int32)
            for k, part in enumerate(self.parts):
                pz[k, :len(part)] = part
            handle = C.c_void_p()
            N.check(N.lib.snfft_wreath_create(n, N.ptr(a), N.ptr(pz), int(self.mode == 'full'), device,
                                              C.byref(handle)))
            self._c = handle
            dims = np.zeros(6, dtype=np.int64)
            N.check(N.lib.snfft_wreath_dims(self._c, N.ptr(dims)))
            assert (int(dims[0]), int(dims[1]), int(dims[2])) == (self.D, self.N, self.block)
            self.nH = int(dims[3])
            self.dft_m = n if self.mode == 'full' else n - 1
            return
        tabs = _alpha_tables(self.alpha, device)
        T = tabs['T']
        if self.mode is not None:
            self.dft_m = n if self.mode == 'full' else n - 1
            self.freq = torch.from_numpy(dft_frequencies(self.alpha, self.reps, self.mode)).to(dev)
        else:
            # characters chi_i(o) (wreath.py:34-58): o o t_i summed over the alpha-segments, the
            # k-th segment weighted by k (px_wreath.py:17-30 for two segments)
            og = self.oris[:, T - 1]                                      # [n_ori, N, n]
            expo = (og * self.seg).sum(-1) % m
            omega = np.exp(2j * np.pi * np.arange(m) / float(m)) if m != 2 else np.array([1.0 + 0j, -1.0 + 0j])
            X = omega[expo]                                               # [n_ori, N]
            self.Xre = torch.from_numpy(np.ascontiguousarray(X.real)).to(dev)
            self.Xim = torch.from_numpy(np.ascontiguousarray(X.imag)).to(dev)
            self.XreT = self.Xre.t().contiguous()
            self.XimT = self.Xim.t().contiguous()
        # R[h] = kron_k rho_{parts_k}(h_k), h in itertools.product order (wreath.py:191-224)
        R = None
        for ak, pk in zip(self.alpha, self.parts):
            if ak == 0:
                continue
            plan = get_plan(ak, device)
            perms = torch.tensor(list(itertools.permutations(range(1, ak + 1))), dtype=torch.int32, device=dev)
            mats = plan.yor(plan.irrep_index(pk), perms)              # device YOR of the factor
            if R is None:
                R = mats
            else:
                R = torch.einsum('xab,ycd->xyacbd', R, mats).reshape(
                    R.shape[0] * mats.shape[0], R.shape[1] * mats.shape[1], R.shape[2] * mats.shape[2])
        self.nH = R.shape[0]
        assert self.nH == tabs['nH']
        self.R = R.reshape(self.nH, self.block * self.block).contiguous()
        self.RT = self.R.t().contiguous()
        # sigma = t_i h t_j^-1 for every (i, j, h): one gather index into F[sigma][i]
        self.idx, self.inv_idx = tabs['idx'], tabs['inv_idx']

    def close(self):
        if getattr(self, '_c', None):
            N.lib.snfft_wreath_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _tables(self):
        """R (Young-subgroup matrices) and the coset tables for the host-side helpers of a plan that
        runs behind the C ABI (``inverse_at``)."""
        import torch
        if getattr(self, 'R', None) is None:
            dev = f'cuda:{self.device}'
            R = None
            for ak, pk in zip(self.alpha, self.parts):
                if ak == 0:
                    continue
                plan = get_plan(ak, self.device)
                perms = torch.tensor(list(itertools.permutations(range(1, ak + 1))), dtype=torch.int32, device=dev)
                mats = plan.yor(plan.irrep_index(pk), perms)
                R = mats if R is None else torch.einsum('xab,ycd->xyacbd', R, mats).reshape(
                    R.shape[0] * mats.shape[0], R.shape[1] * mats.shape[1], R.shape[2] * mats.shape[2])
            self.R = R.reshape(R.shape[0], self.block * self.block).contiguous()
        return self.R

    # -- thin wrappers over the C ABI
    def _gemm(self, a, b):
        import torch
        m, k = a.shape
        n = b.shape[1]
        out = torch.empty((m, n), dtype=torch.float64, device=a.device)
        N.check(N.lib.snfft_dgemm(N.ptr(a), N.ptr(b), N.ptr(out), m, n, k,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def _gather(self, src, idx):
        import torch
        out = torch.empty(idx.numel(), dtype=torch.float64, device=src.device)
        N.check(N.lib.snfft_gather(N.ptr(src), N.ptr(idx), N.ptr(out), idx.numel(), 1,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def _dft(self, re, im, inverse):
        """``snfft_c3_dft``: forward real [n!, 3^m] -> (re, im) [n!, N]; inverse the adjoint."""
        import torch
        count = re.shape[0]
        width = 3 ** self.dft_m if inverse else self.N
        out_re = torch.empty((count, width), dtype=torch.float64, device=re.device)
        out_im = torch.empty((count, width), dtype=torch.float64, device=re.device)
        N.check(N.lib.snfft_c3_dft(self.dft_m, N.ptr(re), N.ptr(im) if im is not None else None, N.ptr(self.freq),
                                   self.N, count, N.ptr(out_re), N.ptr(out_im), int(inverse),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out_re, out_im

    def character_sums(self, f):
        """Stage 1: (re, im) of F_i(sigma) = sum_o f(sigma, o) chi_i(o), each [n!, N]."""
        if self.mode is not None:
            if getattr(self, 'freq', None) is None:
                import torch
                self.freq = torch.from_numpy(dft_frequencies(self.alpha, self.reps, self.mode)).to(f.device)
            return self._dft(f, None, False)
        return self._gemm(f, self.Xre), self._gemm(f, self.Xim)

    def forward(self, f):
        """complex128 [D, D] = sum_g f(g) rho(g)."""
        import torch
        if f.dtype != torch.float64 or not f.is_cuda or tuple(f.shape) != (math.factorial(self.n), len(self.oris)):
            raise TypeError('f must be a float64 CUDA tensor [n!, n_orientations]')
        f = f.contiguous()
        if self._c is not None:
            re = torch.empty((self.D, self.D), dtype=torch.float64, device=f.device)
            im = torch.empty((self.D, self.D), dtype=torch.float64, device=f.device)
            with torch.cuda.device(self.device):
                N.check(N.lib.snfft_wreath_forward(self._c, N.ptr(f), N.ptr(re), N.ptr(im),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return torch.complex(re, im)
        with torch.cuda.device(self.device):
            planes = []
            for F in self.character_sums(f):                          # [n!, N]
                G = self._gather(F.reshape(-1), self.idx)             # [N*N*nH] regrouped by cosets
                planes.append(self._gemm(G.reshape(self.N * self.N, self.nH), self.R))
        out = torch.complex(planes[0], planes[1]).reshape(self.N, self.N, self.block, self.block)
        return out.permute(0, 2, 1, 3).reshape(self.D, self.D)

    def inverse(self, fhat):
        """complex128 [n!, n_ori]: D tr(rho(g)^H f_hat) / |G| at every group element
        (cube_ift.py:56-67 divided by the group order, test_inv_fft.py:140)."""
        import torch
        if self._c is not None:
            fre, fim = fhat.real.contiguous(), fhat.imag.contiguous()
            # interleaved complex128 straight from the kernel (out_im = NULL): no plane merge pass
            out = torch.empty((math.factorial(self.n), len(self.oris)), dtype=torch.complex128, device=fhat.device)
            with torch.cuda.device(self.device):
                N.check(N.lib.snfft_wreath_inverse(self._c, N.ptr(fre), N.ptr(fim), N.ptr(out), None,
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return out
        B = fhat.reshape(self.N, self.block, self.N, self.block).permute(0, 2, 1, 3).reshape(
            self.N * self.N, self.block * self.block)
        with torch.cuda.device(self.device):
            E = []
            for plane in (B.real.contiguous(), B.imag.contiguous()):
                e = self._gemm(plane, self.RT)                        # [N*N, nH]
                E.append(self._gather(e.reshape(-1), self.inv_idx).reshape(math.factorial(self.n), self.N))
            if self.mode is not None:
                re, im = self._dft(E[0], E[1], True)
            else:
                re = self._gemm(E[0], self.XreT) + self._gemm(E[1], self.XimT)
                im = self._gemm(E[1], self.XreT) - self._gemm(E[0], self.XimT)
        return torch.complex(re, im) * (self.D / self.order)

    def inverse_at(self, fhat, elements):
        """D tr(rho(g)^H f_hat) / |G| at chosen group elements [(orientation tuple, permutation
        tuple), ...] (cube_ift.py:56-67 per state, the consumer-facing form of
        compute_policy.py:66-88).  Only the rows sigma = t_i h t_j^-1 that g touches are formed:
        for every j the coset i of sigma t_j is known, so the trace runs over N blocks, not N^2."""
        import torch
        dev = fhat.device
        if self._c is not None:
            # one launch for all states (snfft_wreath_inverse_at): sigma as lexicographic ranks, the
            # orientation as its index on the plan's axis (base-3 number of the first dft_m letters)
            sig = _lex_ranks(np.array([s for _, s in elements], dtype=np.int64).reshape(-1, self.n))
            o = np.array([t for t, _ in elements], dtype=np.int64).reshape(-1, self.n)
            if self.mode == 'zero_sum' and (o.sum(axis=1) % 3).any():
                raise ValueError('orientation tuples must sum to 0 mod 3 on this plan')
            oi = (o[:, :self.dft_m] @ (3 ** np.arange(self.dft_m - 1, -1, -1))).astype(np.int32)
            fre, fim = fhat.real.contiguous(), fhat.imag.contiguous()
            sd, od = torch.from_numpy(sig).to(dev), torch.from_numpy(oi).to(dev)
            re = torch.empty(len(elements), dtype=torch.float64, device=dev)
            im = torch.empty_like(re)
            with torch.cuda.device(self.device):
                N.check(N.lib.snfft_wreath_inverse_at(self._c, N.ptr(fre), N.ptr(fim), N.ptr(sd), N.ptr(od), len(elements),
                                                      N.ptr(re), N.ptr(im),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return torch.complex(re, im)
        return self.inverse_at_host(fhat, elements)

    def inverse_at_host(self, fhat, elements):
        """The same values with the coset search driven from the host, one state at a time: the
        path of plans without a C handle (general C_m, arbitrary orientation lists) and an
        independent cross-check of the kernel."""
        import torch
        dev = fhat.device
        b, N_ = self.block, self.N
        T = np.array(self.reps, dtype=np.int64)
        index = {tuple(t): i for i, t in enumerate(self.reps)}
        m = self.cyc_order
        bounds = np.cumsum((0,) + self.alpha)
        blocks = fhat.reshape(N_, b, N_, b).permute(0, 2, 1, 3)       # [i, j, b, b]
        omega = np.exp(2j * np.pi * np.arange(m) / float(m)) if m != 2 else np.array([1.0 + 0j, -1.0 + 0j])
        Hlist = [p.tup_rep for p in young_subgroup_perm(self.alpha)]
        hindex = {h: q for q, h in enumerate(Hlist)}
        out = []
        Rm = self._tables().reshape(len(Hlist), b, b)
        for ori, sigma in elements:
            sig = np.array(sigma, dtype=np.int64)
            ii, hh = [], []
            for j in range(N_):
                st = sig[T[j] - 1]                                    # sigma t_j
                rep = np.concatenate([np.sort(st[bounds[k]:bounds[k + 1]]) for k in range(m)])
                i = index[tuple(int(v) for v in rep)]
                h = tuple(int(v) for v in (np.argsort(T[i]) + 1)[st - 1])   # t_i^-1 sigma t_j
                ii.append(i)
                hh.append(hindex[h])
            o = np.array(ori, dtype=np.int64)
            expo = (o[T[ii] - 1] * self.seg).sum(-1) % m
            chi = torch.from_numpy(omega[expo]).to(dev)               # D(o)_ii
            it = torch.tensor(ii, device=dev)
            jt = torch.arange(N_, device=dev)
            blk = blocks[it, jt]                                      # f_hat block (i(j), j): [N, b, b]
            rho = Rm[torch.tensor(hh, device=dev)]                    # rho_parts(h_j): [N, b, b] real
            val = (chi.conj()[:, None, None] * rho * blk).sum()
            out.append(val * (self.D / self.order))
        return torch.stack(out)


def wreath_forward(f, alpha, parts, oris=None, device: int = 0):
    return WreathPlan(alpha, parts, oris, device).forward(f)
