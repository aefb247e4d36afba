"""Ferrers diagrams and standard Young tableaux with the reference's interface
(reference ``young_tableau.py``).

The basis order of every irrep matrix is the reference's: standard tableaux sorted by
last-letter order (young_tableau.py:193-206).  That list is the concatenation, over
``branch_down()`` order, of the tableaux of each sub-shape with the largest letter placed in the
removed corner, so it is generated here by recursion on the Young lattice instead of filtering
all (n-1)! fillings (young_tableau.py:98-125) -- which is what makes n = 10..12 reachable.
"""
from __future__ import annotations

from functools import total_ordering

FERRERS_CACHE = {}


def get_minus_partition(partition, idx):
    """The partition with one box removed from row ``idx`` (young_tableau.py:22-39)."""
    rows = list(partition)
    rows[idx] -= 1
    return tuple(r for r in rows if r > 0)


def _corner_rows(partition):
    """Rows whose last box can be removed, top to bottom."""
    rows = []
    for r, length in enumerate(partition):
        below = partition[r + 1] if r + 1 < len(partition) else 0
        if length - 1 >= below:
            rows.append(r)
    return rows


class FerrersDiagram:
    TABLEAUX_CACHE = {}

    def __init__(self, partition):
        self.partition = tuple(partition)
        self.size = sum(self.partition)
        cached = FerrersDiagram.TABLEAUX_CACHE.get(self.partition)
        self.tableaux = cached if cached is not None else self.gen_tableaux()
        FERRERS_CACHE[self.partition] = self

    @staticmethod
    def from_partition(partition):
        partition = tuple(partition)
        if partition not in FERRERS_CACHE:
            FERRERS_CACHE[partition] = FerrersDiagram(partition)
        return FERRERS_CACHE[partition]

    def branch_down(self):
        """Sub-diagrams obtained by removing a corner box, rows top to bottom
        (young_tableau.py:62-81); empty for a single box."""
        if self.size == 1:
            return []
        return [FerrersDiagram.from_partition(get_minus_partition(self.partition, r))
                for r in _corner_rows(self.partition)]

    def gen_tableaux(self, perms=None):
        """Sorted list of the standard tableaux of this shape, ``idx`` set on each."""
        if self.size == 0:
            return []
        if perms is not None:
            # explicit candidate fillings (the reference's optional argument): filter and sort
            tabs = sorted(make_young_tableau(self.partition, tuple(p)) for p in perms
                          if YoungTableau.valid_static(self.partition, tuple(p)))
        else:
            tabs = [_tableau_from_rows(self.partition, rows) for rows in _standard_fillings(self.partition)]
        for idx, tab in enumerate(tabs):
            tab.set_idx(idx)
        FerrersDiagram.TABLEAUX_CACHE[self.partition] = tabs
        return tabs

    def n_tabs(self):
        return len(self.tableaux)

    def pp_tableaux(self):
        for tab in self.tableaux:
            print(tab)
            print('======')

    def __repr__(self):
        return '\n'.join('[ ]' * length for length in self.partition)


_FILLINGS = {}


def _standard_fillings(partition):
    """Row contents of every standard tableau, in last-letter order."""
    partition = tuple(partition)
    hit = _FILLINGS.get(partition)
    if hit is not None:
        return hit
    n = sum(partition)
    if n == 1:
        out = [((1,),)]
    else:
        out = []
        for r in _corner_rows(partition):
            for rows in _standard_fillings(get_minus_partition(partition, r)):
                grown = list(rows)
                if r < len(grown):
                    grown[r] = grown[r] + (n,)
                else:
                    grown.append((n,))
                out.append(tuple(grown))
    _FILLINGS[partition] = out
    return out


def _tableau_from_rows(partition, rows):
    vals = tuple(x for row in rows for x in row)
    return YoungTableau(list(rows), vals)


def n_tabs(partition):
    return FerrersDiagram.from_partition(partition).n_tabs()


def wreath_dim(parts):
    """Dimension of the S_alpha irrep indexed by a tuple of partitions (young_tableau.py:134-145)."""
    out = 1
    for p in parts:
        if len(p):
            out *= n_tabs(p)
    return out


def make_young_tableau(shape, vals):
    """Fill ``shape`` row by row with ``vals`` (young_tableau.py:147-166)."""
    contents, at = [], 0
    for length in shape:
        contents.append(tuple(vals[at:at + length]))
        at += length
    return YoungTableau(contents, tuple(vals))


def swap(x, i, j):
    return j if x == i else (i if x == j else x)


@total_ordering
class YoungTableau:
    CACHE = {}

    def __init__(self, contents, vals=None):
        self.contents = contents
        self.partition = tuple(len(row) for row in contents)
        self.size = sum(self.partition)
        self.vals = vals
        self.idx = None
        self._row, self._col = {}, {}
        for r, row in enumerate(contents):
            for c, x in enumerate(row):
                self._row[x] = r + 1
                self._col[x] = c + 1
        YoungTableau.CACHE.setdefault(self.partition, {})[vals] = self

    def __lt__(self, other):
        """Last-letter order: the largest letter on which the two differ sits in a higher row
        (young_tableau.py:193-206)."""
        for letter in range(self.size, 0, -1):
            mine, theirs = self.get_row(letter), other.get_row(letter)
            if mine != theirs:
                return mine < theirs
        return False

    def __eq__(self, other):
        return isinstance(other, YoungTableau) and self.contents == other.contents

    def __hash__(self):
        return hash(tuple(tuple(r) for r in self.contents))

    def set_idx(self, idx):
        self.idx = idx

    @staticmethod
    def valid_static(shape, contents):
        """Rows and columns increase when ``contents`` fills ``shape`` row by row."""
        starts, at = [], 0
        for length in shape:
            starts.append(at)
            at += length
        for r, length in enumerate(shape):
            for c in range(length):
                here = contents[starts[r] + c]
                if c > 0 and here < contents[starts[r] + c - 1]:
                    return False
                if r > 0 and here < contents[starts[r - 1] + c]:
                    return False
        return True

    def valid(self):
        for r, row in enumerate(self.contents):
            for c, x in enumerate(row):
                if c > 0 and row[c - 1] > x:
                    return False
                if r > 0 and self.contents[r - 1][c] > x:
                    return False
        return True

    def get_row(self, val):
        return self._row[val]

    def get_col(self, val):
        return self._col[val]

    def get_row2(self, val):
        return self._row.get(val)

    def get_col2(self, val):
        return self._col.get(val)

    def transpose(self, transposition):
        """The tableau with letters i, j exchanged if that is standard, else None
        (young_tableau.py:305-312)."""
        i, j = transposition
        swapped = tuple(swap(x, i, j) for x in self.vals)
        return YoungTableau.CACHE[self.partition].get(swapped, None)

    def content(self, x):
        return self.get_col(x) - self.get_row(x)

    def dist(self, i, j):
        """Axial distance used by YOR: content(j) - content(i) (young_tableau.py:326-327)."""
        return self.content(j) - self.content(i)

    def ax_dist(self, i, j):
        return (self.get_col(j) - self.get_col(i)) + (self.get_row(i) - self.get_row(j))

    def __repr__(self):
        return '\n'.join(''.join('[{}]'.format(x) for x in row) for row in self.contents)
