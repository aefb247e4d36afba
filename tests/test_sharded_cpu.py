"""The sharded (multi-GPU) transform on CPU: shard tables from the C library, compute replaced by
numpy (oracle for the local S_m blocks, numpy interpretation of the shard's level tables), exchange
over a world_size-2 Gloo group.  Checks that the N > 1 path reproduces the oracle's transform."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import snfft_oracle as O
from conftest import rel_fro
from snfft_b200 import Plan
from snfft_b200.sharded import ShardedPlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CpuShardedPlan(ShardedPlan):
    """ShardedPlan with the device compute swapped for numpy (host tables only, device = -1)."""

    def __init__(self, n, rank, world, split):
        self.n, self.rank, self.world, self.device, self.group = n, rank, world, -1, None
        self.m = split
        self.plan = Plan(n, -1)
        self.local_plan = None
        import ctypes as C
        from snfft_b200 import _native as N
        handle = C.c_void_p()
        N.check(N.lib.snfft_shard_create(self.plan._h, rank, world, split, C.byref(handle)))
        self._h = handle
        self.mfact = math.factorial(split)
        self.nblocks = math.factorial(n) // self.mfact
        self.block_range = []
        for r in range(world):
            b, e = C.c_int64(), C.c_int64()
            N.check(N.lib.snfft_shard_blocks(self._h, r, C.byref(b), C.byref(e)))
            self.block_range.append((b.value, e.value))
        self.sizes = [[self._size(r, k) for k in range(n + 1)] for r in range(world)]
        self._idx = {}

    def _local_forward(self, blocks):
        # coset-major S_m blocks -> packed spectra, through the oracle (lexicographic input)
        c2l = O.coset_major_to_lex(self.m)
        lex = np.empty_like(blocks.numpy())
        lex[:, c2l] = blocks.numpy()
        return torch.from_numpy(O.pack(O.fft_all(lex), self.m))

    def _local_inverse(self, spectra):
        c2l = O.coset_major_to_lex(self.m)
        out = np.stack([O.inverse_direct(O.unpack(s, self.m), self.m) for s in spectra.numpy()])
        return torch.from_numpy(np.ascontiguousarray(out[:, c2l]))

    def _level(self, k, x, inverse):
        x = x.numpy()
        groups = math.factorial(self.n) // math.factorial(k)
        nprog = len(self.plan.layout(k - 1)[0])
        progs = [self.export_panel(k, p) for p in range(nprog)]
        ins, outs = progs[0]['in_stride'], progs[0]['out_stride']
        res = np.zeros(groups * (ins if inverse else outs))
        for g in range(groups):
            src = x[g * (outs if inverse else ins):(g + 1) * (outs if inverse else ins)]
            dst = res[g * (ins if inverse else outs):(g + 1) * (ins if inverse else outs)]
            for prog in progs:
                w, d = prog['width'], prog['d']
                if w == 0:
                    continue
                rd = prog['dest'] if inverse else prog['rowbase']
                wr = prog['rowbase'] if inverse else prog['dest']
                buf = np.stack([src[rd[s]:rd[s] + w] for s in range(k * d)])
                order = range(len(prog['m']))
                for b in (reversed(order) if inverse else order):
                    mm = prog['m'][b]
                    sl = prog['slots'][b, :mm]
                    cm = prog['coef_inv' if inverse else 'coef_fwd'][b, :mm * mm].reshape(mm, mm)
                    buf[sl] = cm @ buf[sl]
                for s in range(k * d):
                    dst[wr[s]:wr[s] + w] = buf[s]
        return torch.from_numpy(res)


def _worker(rank, world, n, split, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sp = CpuShardedPlan(n, rank, world, split)
        f = np.random.default_rng(77).standard_normal(math.factorial(n))       # lexicographic
        coset = f[O.coset_major_to_lex(n)]
        b, e = sp.local_elements()
        compact = sp.forward(torch.from_numpy(coset[b:e].copy()))
        full = sp.gather_spectrum(compact).numpy()
        back = sp.inverse(compact).numpy()
        ref = O.pack(O.fft_all(f), n)
        ret[rank] = (float(np.linalg.norm(full - ref) / np.linalg.norm(ref)),
                     float(np.abs(back - coset[b:e]).max()),
                     [int(sp.sizes[r][n]) for r in range(world)])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n,split', [(5, 3), (6, 4), (6, 5), (5, 4)])
def test_two_rank_sharded_transform_matches_oracle(n, split):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() * 7 + n * 13 + split) % 2000
    mp.spawn(_worker, args=(world, n, split, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err_fwd, err_inv, sizes = ret[rank]
        assert err_fwd < 1e-12 and err_inv < 1e-10, (rank, err_fwd, err_inv)
        assert sum(sizes) == math.factorial(n)


def test_shard_bookkeeping_single_process():
    n, m = 7, 4
    for world in (1, 2, 3, 8):
        plans = [CpuShardedPlan(n, r, world, m) for r in range(world)]
        sp = plans[0]
        # blocks tile [0, n!/m!) and column ownership tiles every irrep
        assert sp.block_range[0][0] == 0 and sp.block_range[-1][1] == math.factorial(n) // math.factorial(m)
        assert all(sp.block_range[r][1] == sp.block_range[r + 1][0] for r in range(world - 1))
        parts, dims, _ = sp.plan.layout()
        for i, d in enumerate(dims):
            cols = np.concatenate([sp.columns(i, r) for r in range(world)])
            assert sorted(cols.tolist()) == list(range(d))
        idx = np.concatenate([sp.gather_index(r) for r in range(world)])
        assert sorted(idx.tolist()) == list(range(math.factorial(m)))
        for k in range(m, n + 1):
            assert sum(sp.sizes[r][k] for r in range(world)) == math.factorial(k)
    with pytest.raises(Exception):
        CpuShardedPlan(5, 0, 200, 4)        # fewer blocks than ranks
