"""Golden vectors of the wreath-product path from the UNMODIFIED reference (build container only):
    python tests/golden/make_golden_wreath.py
Writes tests/golden/reference_wreath.npz.  The reference computes the cyclic scalars in complex64
(wreath.py:53, :331), so these vectors are accurate to ~1e-7 only."""
import itertools
import math
import os
import sys
import warnings

import numpy as np

np.math = math
warnings.filterwarnings('ignore')
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

import coset_utils            # noqa: E402
import multi                  # noqa: E402
import perm2                  # noqa: E402
import wreath                 # noqa: E402
from young_tableau import wreath_dim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [((0, 3, 2), ((), (2, 1), (1, 1))),       # test_wreath.py:16-35
         ((1, 2, 2), ((1,), (1, 1), (2,))),       # SURVEY appendix probe
         ((2, 2, 1), ((2,), (1, 1), (1,)))]


def main():
    out = {}
    rng = np.random.default_rng(38)
    for ci, (alpha, parts) in enumerate(CASES):
        n = sum(alpha)
        sn = perm2.sn(n)
        ydict = wreath.wreath_yor(alpha, parts)
        reps = coset_utils.coset_reps(sn, coset_utils.young_subgroup_perm(alpha))
        cyc = wreath.cyclic_irreps(alpha)
        out[f'case{ci}_alpha'] = np.array(alpha)
        out[f'case{ci}_reps'] = np.array([p.tup_rep for p in reps])
        out[f'case{ci}_young'] = np.array([p.tup_rep for p in coset_utils.young_subgroup_perm(alpha)])
        # representation matrices at a few group elements
        gs, os_, mats = [], [], []
        for _ in range(6):
            g = tuple(int(x) + 1 for x in rng.permutation(n))
            o = tuple(int(x) for x in rng.integers(0, 3, n))
            gs.append(g)
            os_.append(o)
            mats.append(wreath.wreath_rep(o, g, ydict, reps, cyc))
        out[f'case{ci}_g'] = np.array(gs)
        out[f'case{ci}_o'] = np.array(os_)
        out[f'case{ci}_rho'] = np.array(mats)
        # direct Fourier transform of a random function on {sum(o) = 0 mod 3} x S_n
        oris = [t for t in itertools.product((0, 1, 2), repeat=n) if sum(t) % 3 == 0]
        f = np.random.default_rng(100 + ci).standard_normal((len(sn), len(oris)))
        save = {}
        for r, p in enumerate(sn):
            for q, o in enumerate(oris):
                scal = wreath.block_cyclic_irreps(o, reps, cyc)
                multi.mult_yor_block(ydict[p.tup_rep], f[r, q], scal, save)
        block = wreath_dim(parts)
        out[f'case{ci}_fhat'] = multi.convert_yor_matrix(save, block, len(reps))
        print(alpha, parts, 'D =', out[f'case{ci}_fhat'].shape[0])
    np.savez_compressed(os.path.join(HERE, 'reference_wreath.npz'), **out)


if __name__ == '__main__':
    main()
