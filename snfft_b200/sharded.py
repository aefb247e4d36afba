"""Multi-GPU S_n transform of ONE function, sharded by the coset recursion.

One process per GPU (``torch.distributed``).  Rank r owns a contiguous range of the n!/m! blocks
of S_m that make up the function in coset-major order, transforms them locally (levels 2..m, the
ordinary single-GPU kernels), takes part in ONE exchange in which every rank receives all blocks
but only its own columns, and finishes levels m+1..n on its column shard without further
communication.  The result on rank r is ``f_hat(lambda)[:, owned columns]`` for every lambda.

The reference's only distributed pattern is "every rank computes a full-size partial f_hat, gather
to rank 0, sum" (cube_ft.py:102-106); because columns of f_hat are independent the exchange here
moves world-times fewer bytes and needs no reduction.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native as N
from .plan import Plan, get_plan


class ShardedPlan:
    def __init__(self, n: int, rank: int, world: int, device: int = 0, split: int | None = None,
                 group=None):
        self.n, self.rank, self.world, self.device, self.group = n, rank, world, device, group
        if split is None:
            # the largest m < n whose block count n!/m! is a multiple of `world` (even load).  Uneven
            # block counts are supported (pass `split`): for S_12 on 8 GPUs m = 10 (17 / 16 blocks)
            # is 6 % faster per rank in kernel time than m = 9, but its first 8-GPU run showed a
            # host-side stall that is not understood yet, so it is not the default.
            split = max(1, n - 1)
            for m in range(n - 1, 0, -1):
                if (math.factorial(n) // math.factorial(m)) % world == 0:
                    split = m
                    break
        self.m = split
        self.plan = get_plan(n, device)
        self.local_plan = get_plan(self.m, device)
        handle = C.c_void_p()
        N.check(N.lib.snfft_shard_create(self.plan._h, rank, world, self.m, C.byref(handle)))
        self._h = handle
        self.mfact = math.factorial(self.m)
        self.nblocks = math.factorial(n) // self.mfact
        self.block_range = []
        for r in range(world):
            b, e = C.c_int64(), C.c_int64()
            N.check(N.lib.snfft_shard_blocks(self._h, r, C.byref(b), C.byref(e)))
            self.block_range.append((b.value, e.value))
        self.sizes = [[self._size(r, k) for k in range(n + 1)] for r in range(world)]
        self._idx = {}
        self._p2p = None        # set by enable_p2p(): the peer-memory workspace of the C schedule

    def close(self):
        if getattr(self, '_h', None):
            N.lib.snfft_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host-side description
    def _size(self, r, k):
        if k < self.m:
            return 0
        out = C.c_int64()
        N.check(N.lib.snfft_shard_size(self._h, r, k, C.byref(out)))
        return out.value

    def gather_index(self, r: int) -> np.ndarray:
        """Positions inside a packed S_m spectrum that rank r needs (its columns), compact order."""
        if r not in self._idx:
            idx = np.empty(self.sizes[r][self.m], dtype=np.int64)
            N.check(N.lib.snfft_shard_gather_index(self._h, r, N.ptr(idx)))
            self._idx[r] = idx
        return self._idx[r]

    def columns(self, irrep: int, r: int | None = None, k: int | None = None) -> np.ndarray:
        """Columns of f_hat(irrep) owned by rank r (default: this rank), in compact order."""
        r = self.rank if r is None else r
        k = self.n if k is None else k
        cnt = C.c_int32()
        N.check(N.lib.snfft_shard_columns(self._h, r, k, irrep, None, C.byref(cnt)))
        cols = np.empty(cnt.value, dtype=np.int32)
        N.check(N.lib.snfft_shard_columns(self._h, r, k, irrep, N.ptr(cols), C.byref(cnt)))
        return cols

    def export_panel(self, k: int, panel: int):
        prog = self.plan.export_panel(k, panel)
        rows = k * prog['d']
        width, ins, outs = C.c_int32(), C.c_int64(), C.c_int64()
        dest = np.zeros(rows, dtype=np.int64)
        rowbase = np.zeros(rows, dtype=np.int64)
        N.check(N.lib.snfft_shard_export_panel(self._h, k, panel, C.byref(width), N.ptr(dest),
                                               N.ptr(rowbase), C.byref(ins), C.byref(outs)))
        prog.update(width=width.value, dest=dest, rowbase=rowbase, in_stride=ins.value,
                    out_stride=outs.value)
        return prog

    def local_elements(self, r: int | None = None):
        """[begin, end) of the coset-major function owned by rank r."""
        b, e = self.block_range[self.rank if r is None else r]
        return b * self.mfact, e * self.mfact

    def unpack(self, compact, r: int | None = None):
        """{partition: [d, w]} views of a compact level-n spectrum."""
        r = self.rank if r is None else r
        parts, dims, _ = self.plan.layout()
        out, at = {}, 0
        for i, (p, d) in enumerate(zip(parts, dims)):
            w = len(self.columns(i, r))
            out[p] = compact[at:at + d * w].reshape(d, w)
            at += d * w
        return out

    # ------------------------------------------------------------------ compute hooks (device)
    def _local_forward(self, blocks):
        return self.local_plan.forward(blocks, 'coset')

    def _local_inverse(self, blocks):
        return self.local_plan.inverse(blocks, 'coset')

    def _level(self, k, x, inverse, out=None):
        import torch
        groups = math.factorial(self.n) // math.factorial(k)
        size_out = self.sizes[self.rank][k - 1] * k if inverse else self.sizes[self.rank][k]
        if out is None:
            out = torch.empty(groups * size_out, dtype=torch.float64, device=x.device)
        else:
            out = out.reshape(-1)[:groups * size_out]
        # x is always an intermediate of this class, so the forward direction may overwrite it
        xs, ss = (out, x) if inverse else (x, out)
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_shard_level_scratch(self._h, k, N.ptr(xs), N.ptr(ss), groups, int(inverse),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def _index(self, r, like):
        import torch
        key = ('t', r, str(like.device))
        if key not in self._idx:
            self._idx[key] = torch.from_numpy(self.gather_index(r)).to(like.device)
        return self._idx[key]

    # ------------------------------------------------------------------ the one exchange
    def _exchange(self, send, recv):
        """send[q] -> rank q, recv[q] <- rank q.  NCCL: one all-to-all (uneven splits); Gloo has no
        all-to-all for this, so there it is batched point-to-point."""
        import torch.distributed as dist
        if self.world == 1:
            recv[0].copy_(send[0])
            return
        if dist.get_backend(self.group) == 'nccl':
            dist.all_to_all(recv, send, group=self.group)
            return
        recv[self.rank].copy_(send[self.rank])
        ops = []
        for q in range(self.world):
            if q == self.rank:
                continue
            ops.append(dist.P2POp(dist.isend, send[q], self._peer(q), group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv[q], self._peer(q), group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def _peer(self, q):
        import torch.distributed as dist
        return q if self.group is None else dist.get_global_rank(self.group, q)

    # ------------------------------------------------------------------ transforms
    # Each direction is split into the part before and after the exchange so that tests can run
    # several ranks inside one process.
    def forward_local(self, f_local):
        """Levels 2..m on this rank's blocks; returns send[q] = my blocks restricted to q's columns."""
        b, e = self.block_range[self.rank]
        local = self._local_forward(f_local.reshape(e - b, self.mfact))
        return [local[:, self._index(q, local)].contiguous() for q in range(self.world)]

    def forward_finish(self, recv):
        """recv[r] = blocks of rank r restricted to my columns; levels m+1..n on my column shard."""
        import torch
        x = torch.cat(recv).reshape(-1)
        for k in range(self.m + 1, self.n + 1):
            x = self._level(k, x, False)
        return x

    def forward(self, f_local):
        """f_local: this rank's slice of the coset-major function (``local_elements()``).
        Returns the compact spectrum of this rank (use ``unpack`` / ``columns``)."""
        import torch
        if getattr(self, '_p2p', None):
            return self.forward_p2p(f_local)
        send = self.forward_local(f_local)
        mine = self.sizes[self.rank][self.m]
        recv = [torch.empty((eb - bb, mine), dtype=send[0].dtype, device=send[0].device)
                for bb, eb in self.block_range]
        self._exchange(send, recv)
        return self.forward_finish(recv)

    def inverse_local(self, compact):
        """Levels n..m+1 backwards on my column shard; returns send[r] = blocks of rank r."""
        x = compact
        for k in range(self.n, self.m, -1):
            x = self._level(k, x, True)
        x = x.reshape(self.nblocks, self.sizes[self.rank][self.m])
        return [x[bb:eb].contiguous() for bb, eb in self.block_range]

    def inverse_finish(self, recv):
        """recv[q] = my blocks restricted to q's columns; levels m..2 backwards."""
        import torch
        b, e = self.block_range[self.rank]
        local = torch.empty((e - b, self.mfact), dtype=recv[0].dtype, device=recv[0].device)
        for q in range(self.world):
            local[:, self._index(q, local)] = recv[q]
        return self._local_inverse(local).reshape(-1)

    def inverse(self, compact):
        """Inverse of ``forward``: returns this rank's slice of the coset-major function."""
        import torch
        if getattr(self, '_p2p', None):
            return self.inverse_p2p(compact)
        send = self.inverse_local(compact)
        b, e = self.block_range[self.rank]
        recv = [torch.empty((e - b, self.sizes[q][self.m]), dtype=send[0].dtype, device=send[0].device)
                for q in range(self.world)]
        self._exchange(send, recv)
        return self.inverse_finish(recv)

    # ------------------------------------------------------------------ peer-memory exchange
    # The same exchange as ONE kernel per direction over NVLink peer memory (snfft_shard_exchange):
    # forward = push (every rank stores rank r's columns of its blocks straight into rank r's
    # column-shard buffer), inverse = pull (every rank loads its blocks out of every rank's
    # column-shard buffer).  No packing / unpacking passes and no NCCL call; the remote side of
    # both is contiguous.  The *_push / *_pull methods take device pointers that this process can
    # dereference; `enable_p2p` gets them from torch's symmetric memory (one process per GPU) and
    # makes forward / inverse use them.
    def column_buffer_elements(self):
        """Elements of a column-shard buffer that fits every rank (symmetric allocation)."""
        return self.nblocks * max(self.sizes[r][self.m] for r in range(self.world))

    def forward_local_spectra(self, f_local):
        """Levels 2..m on this rank's blocks: packed S_m spectra [blocks, m!]."""
        b, e = self.block_range[self.rank]
        return self.local_plan.forward(f_local.reshape(e - b, self.mfact), 'coset')

    def _exchange_kernel(self, peer_ptrs, local, inverse):
        import torch
        arr = (C.c_void_p * self.world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_shard_exchange(self._h, arr, N.ptr(local), int(inverse),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def forward_push(self, spectra, peer_ptrs):
        """peer_ptrs[r] = device pointer of rank r's column-shard buffer: stores rank r's columns of
        my blocks there."""
        self._exchange_kernel(peer_ptrs, spectra, False)

    def forward_finish_columns(self, columns):
        """columns = my column-shard buffer once every rank has pushed; levels m+1..n."""
        x = columns.reshape(-1)[:self.nblocks * self.sizes[self.rank][self.m]]
        for k in range(self.m + 1, self.n + 1):
            x = self._level(k, x, False)
        return x

    def inverse_local_columns(self, compact, out=None):
        """Levels n..m+1 backwards on my column shard: [all blocks, my columns], written into `out`
        (my column-shard buffer) if given."""
        x = compact
        for k in range(self.n, self.m, -1):
            x = self._level(k, x, True, out=out if k == self.m + 1 else None)
        return x

    def inverse_finish_pull(self, peer_ptrs):
        """peer_ptrs[r] = device pointer of rank r's column-shard buffer after its
        `inverse_local_columns`; loads my blocks from every rank's columns, then levels m..2."""
        import torch
        b, e = self.block_range[self.rank]
        local = torch.empty((e - b, self.mfact), dtype=torch.float64, device=torch.device('cuda', self.device))
        self._exchange_kernel(peer_ptrs, local, True)
        return self._local_inverse(local).reshape(-1)

    def enable_p2p(self):
        """Allocate the column-shard exchange buffer in symmetric memory and rendezvous (collective
        over the group).  Returns False - and leaves the NCCL exchange in place - if this torch
        build or this machine cannot map peer memory."""
        import torch
        import torch.distributed as dist
        if getattr(self, '_p2p', None) is not None:
            return True
        if self.world == 1 or self.world > 16 or dist.get_backend(self.group) != 'nccl':
            return False
        group = self.group if self.group is not None else dist.group.WORLD
        dev = torch.device('cuda', self.device)
        state = None
        try:
            import torch.distributed._symmetric_memory as symm
            cols, _local, _work, pad_words = self.workspace_sizes()
            pad = (pad_words + 31) // 32 * 32      # keep the column buffer 256-byte aligned
            buf = symm.empty(pad + cols, dtype=torch.float64, device=dev)
            hc = symm.rendezvous(buf, group)
            buf[:pad].zero_()                       # signal pad: epochs only grow from zero
            torch.cuda.synchronize(dev)
            ptrs = [int(p) for p in hc.buffer_ptrs]
            state = dict(buffer=buf, hc=hc, columns=buf[pad:], ptrs=[p + 8 * pad for p in ptrs],
                         wsk=self.make_workspace([p + 8 * pad for p in ptrs], ptrs))
        except Exception as exc:  # no symmetric memory here: keep the NCCL path
            self._p2p_error = repr(exc)
        # every rank must take the same exchange: agree on the outcome (a rank whose allocation
        # failed would otherwise wait in all_to_all while the others wait in the peer barrier)
        flag = torch.tensor([1 if state is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self._p2p = None
            return False
        dist.barrier(group=group)   # every pad is zeroed before anybody raises a flag
        self._p2p = state
        return True

    # ------------------------------------------------------------------ whole schedule in C
    def workspace_sizes(self):
        """(column buffer, local, shard work, signal pad words): see ``snfft_shard_ws_sizes``."""
        sizes = (C.c_int64 * 4)()
        N.check(N.lib.snfft_shard_ws_sizes(self._h, sizes))
        return [int(v) for v in sizes]

    def make_workspace(self, col_ptrs, pad_ptrs):
        """The persistent workspace of ``snfft_forward_sharded`` / ``snfft_inverse_sharded``:
        col_ptrs[r] / pad_ptrs[r] = device pointers of rank r's column-shard buffer and signal pad
        that THIS process can dereference (symmetric memory / CUDA IPC; plain allocations when all
        ranks live in one process).  Local and shard scratch buffers are allocated here, once."""
        import torch
        dev = torch.device('cuda', self.device)
        cols, local, work, _pad = self.workspace_sizes()
        ws = N.ShardWorkspace()
        for r in range(self.world):
            ws.peer_cols[r] = int(col_ptrs[r])
            ws.peer_pads[r] = int(pad_ptrs[r])
        keep = dict(local_a=torch.empty(max(2, local), dtype=torch.float64, device=dev),
                    local_b=torch.empty(max(2, local), dtype=torch.float64, device=dev),
                    shard_a=torch.empty(max(2, work), dtype=torch.float64, device=dev),
                    shard_b=torch.empty(max(2, work), dtype=torch.float64, device=dev),
                    status=torch.zeros(1, dtype=torch.int32, device=dev))
        for name, t in keep.items():
            setattr(ws, name, t.data_ptr())
        keep['ws'] = ws
        return keep

    def forward_c(self, f_local, wsk, out=None):
        """``snfft_forward_sharded`` on the current stream: the whole forward schedule of this rank."""
        import torch
        if out is None:
            out = torch.empty(self.sizes[self.rank][self.n], dtype=torch.float64, device=f_local.device)
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_forward_sharded(self._h, N.ptr(f_local), N.ptr(out), C.byref(wsk['ws']),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def inverse_c(self, compact, wsk, out=None):
        """``snfft_inverse_sharded`` on the current stream."""
        import torch
        if out is None:
            b, e = self.local_elements()
            out = torch.empty(e - b, dtype=torch.float64, device=compact.device)
        with torch.cuda.device(self.device):
            N.check(N.lib.snfft_inverse_sharded(self._h, N.ptr(compact), N.ptr(out), C.byref(wsk['ws']),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def barrier_timed_out(self):
        """True if a device-side barrier of this rank gave up waiting for a peer (synchronises)."""
        p = getattr(self, '_p2p', None)
        return bool(p and int(p['wsk']['status'].item()) != 0)

    def forward_p2p(self, f_local):
        return self.forward_c(f_local, self._p2p['wsk'])

    def inverse_p2p(self, compact):
        return self.inverse_c(compact, self._p2p['wsk'])

    def gather_spectrum(self, compact):
        """Full packed spectrum assembled on every rank (for checks; moves the whole spectrum)."""
        import torch
        import torch.distributed as dist
        parts, dims, offs = self.plan.layout()
        # ranks own different numbers of columns: gather equal-sized (padded) pieces
        longest = max(self.sizes[r][self.n] for r in range(self.world))
        padded = torch.zeros(longest, dtype=compact.dtype, device=compact.device)
        padded[:compact.numel()] = compact
        pieces = [torch.empty(longest, dtype=compact.dtype, device=compact.device) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(pieces, padded, group=self.group)
        else:
            pieces[0] = padded
        full = torch.empty(math.factorial(self.n), dtype=compact.dtype, device=compact.device)
        for r, piece in enumerate(pieces):
            at = 0
            for i, (d, o) in enumerate(zip(dims, offs)):
                cols = torch.from_numpy(self.columns(i, r).astype(np.int64)).to(compact.device)
                w = len(cols)
                if w:
                    full[o:o + d * d].reshape(d, d)[:, cols] = piece[at:at + d * w].reshape(d, w)
                at += d * w
        return full
