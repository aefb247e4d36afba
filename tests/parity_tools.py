"""Checks of the device transform against the oracle that scale to n = 12 (the oracle as the
CHECKER only; shared by the GPU tests and by bench.py's parity block).

Sparse-support input: f = sum_j c_j delta_{sigma_j}  =>  f_hat(lambda) = sum_j c_j rho_lambda(sigma_j)
(fft.py:124-130).  Single columns of rho_lambda(sigma_j) are built on the host from the oracle's
restatement of yor.py (generator word applied to a unit vector), so the largest irreps of S_12
(d = 7700) can be compared without their 474 MB matrices, and every column is addressed by its
GLOBAL index: a column-ownership mistake in the sharded transform cannot cancel out.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(ROOT, 'oracle') not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import snfft_oracle as O  # noqa: E402


def sparse_support(n, count=4, seed=0):
    """`count` random permutations (1-based tuples) with N(0,1) coefficients."""
    rng = np.random.default_rng(seed)
    perms = [tuple(int(v) + 1 for v in rng.permutation(n)) for _ in range(count)]
    coefs = rng.standard_normal(count)
    return perms, coefs


def expected_columns(partition, perms, coefs, cols):
    """Columns `cols` of f_hat(partition) for the sparse-support input."""
    out = np.zeros((len(O.tableaux(partition)), len(cols)))
    for p, c in zip(perms, coefs):
        out += c * O.yor_columns(partition, p, cols)
    return out


def pick_irreps(parts, dims, big=3, small=3):
    """The `big` largest irreps plus `small` spread over the rest (indices into `parts`)."""
    order = sorted(range(len(dims)), key=lambda i: -dims[i])
    rest = order[big:]
    step = max(1, len(rest) // max(1, small))
    return order[:big] + rest[::step][:small]


def check_columns(get_columns, parts, dims, perms, coefs, irreps, owned=None, per_irrep=3):
    """max over the chosen irreps / columns of ||got - expected|| / ||expected||.

    get_columns(i, cols) -> [d_i, len(cols)] numpy array of the device result for GLOBAL column
    indices `cols` of irrep i; owned(i) -> the global columns available (default: all)."""
    worst, checked = 0.0, []
    for i in irreps:
        avail = np.arange(dims[i]) if owned is None else np.asarray(owned(i))
        if len(avail) == 0:
            continue
        pick = sorted({int(avail[0]), int(avail[len(avail) // 2]), int(avail[-1])})[:per_irrep]
        ref = expected_columns(parts[i], perms, coefs, pick)
        got = get_columns(i, pick)
        err = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
        worst = max(worst, err)
        checked.append((parts[i], dims[i], pick, err))
    return worst, checked
