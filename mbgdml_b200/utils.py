"""Fragment enumeration helpers (reference mbgdml/utils.py:449-524)."""
from __future__ import annotations

import itertools

import numpy as np

from . import _native


class CombGenerator:
    """Lazy result of :func:`gen_combs`.

    Iterating it yields the reference's tuples in the reference's order
    (``itertools.product`` order, strictly ascending tuples only).  The predictors do
    not iterate it: they hand ``sets`` to the GPU, which enumerates the same sequence
    on device (``frag_enum``), so nothing is materialised on the host.
    """

    def __init__(self, sets):
        self.sets = [np.asarray(s, dtype=np.int64).ravel() for s in sets]
        self._iter = None

    def to_array(self):
        """The full (n, order) int array, enumerated on the GPU (mbg_enumerate)."""
        return _native.get_context().enumerate(self.sets).astype(np.int64)

    def __iter__(self):
        return self

    def __next__(self):
        if self._iter is None:
            self._iter = iter(tuple(row) for row in self.to_array())
        return next(self._iter)


def gen_combs(sets, replacement=False):
    """utils.py:449-489: combinations with one element per set, without repeats and
    (``replacement=False``) only in ascending order."""
    if replacement:
        raise NotImplementedError("gen_combs(replacement=True) is not used by the prediction path")
    return CombGenerator(sets)


def chunk_iterable(iterable, n):
    """utils.py:492-509."""
    iterator = iter(iterable)
    for first in iterator:
        yield tuple(itertools.chain([first], itertools.islice(iterator, n - 1)))


def chunk_array(array, n):
    """utils.py:512-524."""
    for i in range(0, len(array), n):
        yield array[i : i + n]
