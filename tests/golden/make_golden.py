"""Generate golden vectors by importing the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py [--with-s8]
Writes tests/golden/*.npz next to this file.  Two shims are applied *outside* the reference
tree: ``np.math = math`` (perm2.py:216 uses np.math.factorial, removed in numpy 2) and nothing
else.  Inputs are drawn with numpy default_rng(seed) in lexicographic sn(n) order.
"""
import math
import os
import sys
import time
import warnings

import numpy as np

np.math = math
warnings.filterwarnings('ignore')
REF = '/root/reference'
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import fft as ref_fft            # noqa: E402
import perm2 as ref_perm2        # noqa: E402
import utils as ref_utils        # noqa: E402
import yor as ref_yor            # noqa: E402
from perm import Perm as RefPerm  # noqa: E402
from young_tableau import FerrersDiagram as RefFerrers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def packed(mats, n):
    return np.concatenate([mats[p].reshape(mats[p].shape[:-2] + (-1,)) for p in ref_utils.partitions(n)], axis=-1)


def table_function(values, n):
    """Callable for the reference: values indexed by lexicographic rank; accepts Perm2 (full
    tuples) and Perm (tuples that may be shorter than n)."""
    lut = {p.tup_rep: values[..., i] for i, p in enumerate(ref_perm2.sn(n))}

    def f(p):
        t = tuple(p.tup_rep)
        return lut[t + tuple(range(len(t) + 1, n + 1))]
    return f


def main():
    out = {}
    # --- combinatorics: partitions order, tableaux order, branch_down ---------------------------
    for n in range(1, 8):
        parts = ref_utils.partitions(n)
        out[f'partitions_{n}'] = np.array([p + (0,) * (n - len(p)) for p in parts], dtype=np.int32)
        tabs = []
        for p in parts:
            fd = RefFerrers(p)
            for t in fd.tableaux:
                tabs.append(t.vals)
        out[f'tableaux_{n}'] = np.array(tabs, dtype=np.int32)   # concatenated over partitions
    # --- yor_trans for every shape with n <= 7 ---------------------------------------------------
    for n in range(2, 8):
        for p in ref_utils.partitions(n):
            fd = RefFerrers(p)
            for k in range(1, n):
                out['yortrans_{}_{}'.format('_'.join(map(str, p)), k)] = ref_yor.yor_trans(fd, (k, k + 1))
    # --- yor of random permutations, n = 5..8 ----------------------------------------------------
    rng = np.random.default_rng(2024)
    for n in (5, 6, 7, 8):
        parts = ref_utils.partitions(n)
        for trial in range(4):
            p = parts[int(rng.integers(len(parts)))]
            tup = tuple(int(x) + 1 for x in rng.permutation(n))
            fd = RefFerrers(p)
            mat = ref_yor.yor(fd, ref_perm2.Perm2.from_tup(tup), use_cache=False)
            out['yor_{}_{}'.format('_'.join(map(str, p)), '_'.join(map(str, tup)))] = mat
    # --- the reference's own test function (test.py:76-85), n = 6 ---------------------------------
    f_test = lambda p: 1 if p[1] == 2 else 1.5   # noqa: E731
    out['testpy_fft_6'] = packed({p: ref_fft.fft(f_test, RefFerrers(p)) for p in ref_utils.partitions(6)}, 6)
    out['testpy_ft2_6'] = packed({p: ref_fft.fourier_transform2(f_test, RefFerrers(p)) for p in ref_utils.partitions(6)}, 6)
    # --- random functions, seeds = n (config 1: S_5 seed 5) ---------------------------------------
    for n in range(2, 8):
        values = np.random.default_rng(n).standard_normal(math.factorial(n))
        f = table_function(values, n)
        t0 = time.time()
        out[f'direct_{n}'] = packed({p: ref_fft.fourier_transform2(f, RefFerrers(p)) for p in ref_utils.partitions(n)}, n)
        if n <= 6:
            out[f'clausen_{n}'] = packed({p: ref_fft.fft(f, RefFerrers(p)) for p in ref_utils.partitions(n)}, n)
        print(f'n={n} done in {time.time() - t0:.1f}s', flush=True)
    # --- direct inverse of one irrep contribution (tile/par_inv_ft.py:42), n = 5 ------------------
    n = 5
    values = np.random.default_rng(n).standard_normal(math.factorial(n))
    f = table_function(values, n)
    recon = np.zeros(math.factorial(n))
    for p in ref_utils.partitions(n):
        fd = RefFerrers(p)
        fhat = ref_fft.fourier_transform2(f, fd)
        for i, g in enumerate(ref_perm2.sn(n)):
            rep = ref_yor.yor(fd, g, use_cache=False)
            recon[i] += fhat.shape[0] * (fhat * rep).sum() / math.factorial(n)
    out['inverse_direct_5'] = recon
    np.savez_compressed(os.path.join(HERE, 'reference_small.npz'), **out)
    print('wrote reference_small.npz', len(out), 'arrays')

    if '--with-s8' in sys.argv:
        # config 2: S_8, batch 1024, seed 8; functions 0, 511, 1023 through fourier_transform2
        n = 8
        batch = np.random.default_rng(8).standard_normal((1024, math.factorial(n)))
        pick = batch[[0, 511, 1023]]                      # [3, 8!]
        vals = pick.T.reshape(math.factorial(n), 3, 1, 1)  # f(perm) -> [3,1,1], broadcasts over d x d
        lut = {p.tup_rep: vals[i] for i, p in enumerate(ref_perm2.sn(n))}
        f = lambda p: lut[p.tup_rep]                      # noqa: E731
        t0 = time.time()
        res = {}
        for p in ref_utils.partitions(n):
            res[p] = ref_fft.fourier_transform2(f, RefFerrers(p))
            print(p, f'{time.time() - t0:.0f}s', flush=True)
        np.savez_compressed(os.path.join(HERE, 'reference_s8.npz'), rows=np.array([0, 511, 1023]),
                            direct_8=packed(res, n))
        print('wrote reference_s8.npz')


if __name__ == '__main__':
    main()
