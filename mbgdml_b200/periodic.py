"""Periodic cell for minimum-image n-body predictions (reference mbgdml/periodic.py:32-140)."""
from __future__ import annotations

import numpy as np

from . import _native


class Cell:
    """Holds the cell vectors and the pair cutoff; the minimum-image arithmetic itself
    runs on the GPU (fused into the culling and contraction kernels; ``r_mic``/``d_mic``
    call the standalone ``mbg_r_mic`` operator)."""

    def __init__(self, cell_v, cutoff=None, pbc=True):
        self.cell_v = cell_v  # the setter also resets cutoff to half the shortest vector
        if cutoff is not None:
            self.cutoff = cutoff
        self.pbc = pbc

    @property
    def cell_v(self):
        return getattr(self, "_cell_v", None)

    @cell_v.setter
    def cell_v(self, var):
        var = np.array(var)
        self._cell_v = var
        self.cutoff = np.min(np.linalg.norm(var, axis=1)) / 2.0  # periodic.py:121-126

    @property
    def volume(self):
        vec = self.cell_v
        return np.dot(vec[2], np.cross(vec[0], vec[1]))

    def native_pbc(self):
        """``pbc`` as the three flags ``find_mic(d, cell_v, pbc=self.pbc)`` sees (periodic.py:78; ASE's ``pbc2pbc``
        broadcasts a scalar), int32[3]; ``None`` when all three directions are periodic."""
        flags = np.ascontiguousarray(np.broadcast_to(np.asarray(self.pbc, dtype=bool), (3,)), dtype=np.int32)
        return None if flags.all() else flags

    def native_cell(self):
        return np.ascontiguousarray(self.cell_v, dtype=np.float64).reshape(9), float(self.cutoff)

    def r_mic(self, r):
        """periodic.py:86-107: coordinates relative to the first atom in the minimum image,
        or ``None`` if any pair distance is >= cutoff."""
        assert np.ndim(r) == 2
        cell, cutoff = self.native_cell()
        out, ok = _native.get_context().r_mic(cell, cutoff, r, pbc=self.native_pbc())
        return out if ok else None

    def d_mic(self, d, check_cutoff=True):
        """periodic.py:61-84: minimum image of the vectors ``d``; with ``check_cutoff`` ``None`` when any pair distance
        among the imaged vectors is >= cutoff."""
        d = np.asarray(d, dtype=np.float64)
        assert d.ndim == 2
        cell, cutoff = self.native_cell()
        out, ok = _native.get_context().d_mic(cell, cutoff, d, check_cutoff=check_cutoff, pbc=self.native_pbc())
        if check_cutoff and not ok:
            return None
        return out
