"""Sharded transform on the GPU.  The kernels of the column shards are exercised on ONE device by
running all ranks of a virtual world inside one process (the exchange is a Python hand-over);
with two or more GPUs the same path runs over NCCL with one process per GPU."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import snfft_oracle as O
from conftest import rel_fro
from snfft_b200 import get_plan
from snfft_b200.sharded import ShardedPlan

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('n,world,split', [(6, 2, 4), (8, 2, 7), (8, 3, 5), (9, 4, 7), (10, 8, 7), (9, 4, 8), (8, 3, 7)])
def test_virtual_world_on_one_gpu(n, world, split):
    nf = math.factorial(n)
    f = np.random.default_rng(n * 100 + world).standard_normal(nf)
    coset = torch.from_numpy(np.ascontiguousarray(f[get_plan(n).coset_to_lex()])).cuda()
    ranks = [ShardedPlan(n, r, world, device=0, split=split) for r in range(world)]
    sends = [sp.forward_local(coset[slice(*sp.local_elements())]) for sp in ranks]
    compact = [sp.forward_finish([sends[r][q] for r in range(world)]) for q, sp in enumerate(ranks)]
    # every rank holds its columns of the single-GPU spectrum
    full = get_plan(n).forward(coset, 'coset')
    parts, dims, offs = get_plan(n).layout()
    for q, sp in enumerate(ranks):
        assert compact[q].numel() == sp.sizes[q][n]
        at = 0
        for i, (d, o) in enumerate(zip(dims, offs)):
            cols = torch.from_numpy(sp.columns(i).astype(np.int64)).cuda()
            want = full[o:o + d * d].reshape(d, d)[:, cols]
            got = compact[q][at:at + d * len(cols)].reshape(d, len(cols))
            at += d * len(cols)
            if len(cols):
                assert ((got - want).norm() / want.norm().clamp_min(1e-300)).item() < 1e-10
    assert sum(c.numel() for c in compact) == nf
    if n <= 9:
        ref = O.pack(O.fft_all(f), n)
        assert rel_fro(full.cpu().numpy(), ref) < 1e-10
    # inverse: back to the coset-major slices
    sends = [sp.inverse_local(compact[q]) for q, sp in enumerate(ranks)]
    for r, sp in enumerate(ranks):
        back = sp.inverse_finish([sends[q][r] for q in range(world)])
        b, e = sp.local_elements()
        assert ((back - coset[b:e]).norm() / coset[b:e].norm()).item() < 1e-10


@pytest.mark.parametrize('n,world,split', [(6, 2, 4), (8, 3, 5), (9, 4, 7), (10, 8, 7), (10, 5, 8), (9, 4, 8), (8, 3, 7)])
def test_peer_memory_exchange_on_one_gpu(n, world, split):
    # the exchange kernel (snfft_shard_exchange: forward push, inverse pull) with all ranks of a
    # virtual world in one process: the peer pointers are ordinary device pointers here
    nf = math.factorial(n)
    coset = torch.randn(nf, dtype=torch.float64, device='cuda',
                        generator=torch.Generator(device='cuda').manual_seed(n + world))
    ranks = [ShardedPlan(n, r, world, device=0, split=split) for r in range(world)]
    bufs = [torch.full((sp.column_buffer_elements(),), float('nan'), dtype=torch.float64, device='cuda')
            for sp in ranks]
    ptrs = [t.data_ptr() for t in bufs]
    for sp in ranks:
        sp.forward_push(sp.forward_local_spectra(coset[slice(*sp.local_elements())]), ptrs)
    compact = [sp.forward_finish_columns(bufs[q]) for q, sp in enumerate(ranks)]
    # identical to the pack / hand-over / unpack path
    sends = [sp.forward_local(coset[slice(*sp.local_elements())]) for sp in ranks]
    for q, sp in enumerate(ranks):
        want = sp.forward_finish([sends[r][q] for r in range(world)])
        assert torch.equal(compact[q], want)
    for q, sp in enumerate(ranks):
        sp.inverse_local_columns(compact[q], out=bufs[q])
    for r, sp in enumerate(ranks):
        back = sp.inverse_finish_pull(ptrs)
        b, e = sp.local_elements()
        assert ((back - coset[b:e]).norm() / coset[b:e].norm()).item() < 1e-10


@pytest.mark.parametrize('n,world,split', [(6, 2, 4), (8, 3, 5), (9, 4, 7), (10, 6, 7), (10, 5, 8), (9, 2, 8)])
def test_whole_schedule_in_c_virtual_world(n, world, split):
    """snfft_forward_sharded / snfft_inverse_sharded (the whole schedule of a rank behind one C
    call, device-side barriers between the ranks) with all ranks of a virtual world in one
    process: one stream per rank, "peer" pointers are ordinary device pointers.  (At most 6 ranks:
    the ranks' streams must map to distinct hardware queues - CUDA_DEVICE_MAX_CONNECTIONS is 8 -
    or a spinning barrier kernel blocks the kernels of the rank it waits for.  One process per GPU
    has no such limit.)"""
    nf = math.factorial(n)
    coset = torch.randn(nf, dtype=torch.float64, device='cuda',
                        generator=torch.Generator(device='cuda').manual_seed(3 * n + world))
    ranks = [ShardedPlan(n, r, world, device=0, split=split) for r in range(world)]
    cols_elems, _, _, pad_words = ranks[0].workspace_sizes()
    assert all(sp.workspace_sizes()[0] == cols_elems for sp in ranks)    # symmetric allocation
    cols = [torch.full((cols_elems,), float('nan'), dtype=torch.float64, device='cuda') for _ in ranks]
    pads = [torch.zeros(pad_words, dtype=torch.int64, device='cuda') for _ in ranks]
    wsk = [sp.make_workspace([t.data_ptr() for t in cols], [t.data_ptr() for t in pads]) for sp in ranks]
    streams = [torch.cuda.Stream() for _ in ranks]
    torch.cuda.synchronize()
    for rep in range(2):                                # twice: the barrier epochs keep counting
        compact, back = [None] * world, [None] * world
        for q, sp in enumerate(ranks):
            with torch.cuda.stream(streams[q]):
                compact[q] = sp.forward_c(coset[slice(*sp.local_elements())], wsk[q])
        torch.cuda.synchronize()
        for q, sp in enumerate(ranks):
            with torch.cuda.stream(streams[q]):
                back[q] = sp.inverse_c(compact[q], wsk[q])
        torch.cuda.synchronize()
        if any(int(k['status'].item()) != 0 for k in wsk):
            # two virtual ranks landed on one hardware queue: a barrier gave up after 2 s instead of
            # hanging.  Cannot happen with one process per GPU (test_nccl_two_gpus covers that).
            pytest.skip('the streams of the virtual ranks were serialised by the hardware queues')
        # identical to the pack / hand-over / unpack path of the Python driver
        sends = [sp.forward_local(coset[slice(*sp.local_elements())]) for sp in ranks]
        for q, sp in enumerate(ranks):
            want = sp.forward_finish([sends[r][q] for r in range(world)])
            assert torch.equal(compact[q], want)
            b, e = sp.local_elements()
            assert ((back[q] - coset[b:e]).norm() / coset[b:e].norm()).item() < 1e-10


def _nccl_worker(rank, world, n, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        sp = ShardedPlan(n, rank, world, device=rank)
        gen = torch.Generator(device=f'cuda:{rank}').manual_seed(5)
        coset = torch.randn(math.factorial(n), dtype=torch.float64, device=f'cuda:{rank}', generator=gen)
        b, e = sp.local_elements()
        compact = sp.forward(coset[b:e].contiguous())
        full = sp.gather_spectrum(compact)
        want = get_plan(n, rank).forward(coset, 'coset')
        back = sp.inverse(compact)
        # the same through the peer-memory exchange (symmetric memory), when this box can map it
        p2p = sp.enable_p2p()
        same = True
        if p2p:
            compact2 = sp.forward(coset[b:e].contiguous())
            back2 = sp.inverse(compact2)
            same = bool(torch.equal(compact2, compact)) and bool(torch.equal(back2, back))
        ret[rank] = (((full - want).norm() / want.norm()).item(),
                     ((back - coset[b:e]).norm() / coset[b:e].norm()).item(), p2p, same,
                     getattr(sp, '_p2p_error', ''))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_nccl_two_gpus():
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(2, 9, 29731, ret), nprocs=2, join=True)
    for rank in range(2):
        assert ret[rank][0] < 1e-10 and ret[rank][1] < 1e-10
        assert ret[rank][3], 'peer-memory exchange differs from the NCCL exchange'
    print('peer-memory exchange used:', ret[0][2], ret[0][4])
