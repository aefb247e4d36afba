"""CPU baseline legs for bench.py -- TEST/BENCH INFRASTRUCTURE ONLY (never on the product path).

Times the oracle port of the reference's algorithms on the host cores of the box:
  forward : the Clausen--Baum recursion of fft.py:72-113 (``snfft_oracle.fft_all``), one batch of
            functions per worker process (the pattern of s8fft.py:66-72: a Pool over work items);
  inverse : the reference has no fast inverse, only the direct per-element evaluation of
            tile/par_inv_ft.py:27-45 (``d * sum(fhat o rho(g)) / n!`` per irrep), timed on a sample
            of group elements per worker.
The unmodified reference cannot travel to the GPU box (/root/reference does not exist there) and
needs 43 min for one S_8 function, so the port -- same arithmetic, shared
sub-transforms, numpy-vectorised over the batch -- is what is timed; kind = "port".
"""
from __future__ import annotations

import math
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def _forward_worker(args):
    n, batch, seed = args
    sys.path.insert(0, HERE)
    import numpy as np
    import snfft_oracle as O
    f = np.random.default_rng(seed).standard_normal((batch, math.factorial(n)))
    O.fft_all(f[:1, :math.factorial(min(n, 4))]) if n >= 4 else None   # warm the generator caches
    t0 = time.perf_counter()
    O.fft_all(f)
    return time.perf_counter() - t0


def _inverse_worker(args):
    n, elements, seed = args
    sys.path.insert(0, HERE)
    import numpy as np
    import snfft_oracle as O
    rng = np.random.default_rng(seed)
    fhat = {p: rng.standard_normal((len(O.tableaux(p)),) * 2) for p in O.partitions(n)}
    perms = [tuple(int(v) + 1 for v in rng.permutation(n)) for _ in range(elements)]
    for p in O.partitions(n):                     # warm the generator tables
        O.yor(p, perms[0])
    t0 = time.perf_counter()
    nf = math.factorial(n)
    for g in perms:
        acc = 0.0
        for lam, mat in fhat.items():
            acc += mat.shape[0] * (mat * O.yor(lam, g)).sum()
        acc /= nf
    return time.perf_counter() - t0


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class Baseline:
    """A persistent pool so that repeated steps do not pay process start-up."""

    def __init__(self, n: int, cores: int | None = None, fwd_batch: int = 8, inv_elements: int = 600):
        self.n = n
        self.cores = cores or host_cores()
        self.fwd_batch = fwd_batch
        self.inv_elements = inv_elements
        self.pool = mp.get_context('spawn').Pool(self.cores)

    def close(self):
        self.pool.close()
        self.pool.join()

    def step(self, seed: int = 0):
        """One bounded sample of the forward+inverse workload on all cores.
        Returns (forward elements/s, inverse elements/s, combined element-transforms/s, seconds)."""
        nf = math.factorial(self.n)
        t0 = time.perf_counter()
        self.pool.map(_forward_worker, [(self.n, self.fwd_batch, seed * 1000 + i) for i in range(self.cores)])
        t1 = time.perf_counter()
        self.pool.map(_inverse_worker, [(self.n, self.inv_elements, seed * 1000 + i) for i in range(self.cores)])
        t2 = time.perf_counter()
        r_fwd = self.cores * self.fwd_batch * nf / (t1 - t0)
        r_inv = self.cores * self.inv_elements / (t2 - t1)
        combined = 2.0 / (1.0 / r_fwd + 1.0 / r_inv)   # rate at which a forward+inverse pass would run
        return r_fwd, r_inv, combined, t2 - t0

    def sample_description(self):
        return (f"S_{self.n}: forward = oracle fft_all (Clausen-Baum port) on {self.cores} x {self.fwd_batch} "
                f"functions; inverse = direct per-element evaluation (tile/par_inv_ft.py) on "
                f"{self.cores} x {self.inv_elements} elements; value = 2/(1/r_fwd + 1/r_inv) "
                f"element-transforms/s, the rate of a full forward+inverse pass")
