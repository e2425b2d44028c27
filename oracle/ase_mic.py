"""CPU restatement of `ase.geometry.find_mic` (third-party, absent from /root/reference).

TEST INFRASTRUCTURE ONLY -- nothing under mbgdml_b200/ may import this file.

ASE is listed unpinned in the reference's setup.py:12 and is not installed in
this image; the single call site on the hot path is reference
mbgdml/periodic.py:78 (`find_mic(d, self.cell_v, pbc=self.pbc)`).  This file
restates the published algorithm of ASE >= 3.21
(`ase/geometry/geometry.py`: find_mic, naive_find_mic, general_find_mic,
wrap_positions; `ase/geometry/minkowski_reduction.py`: minkowski_reduce,
after Nguyen & Stehle 2009):

* all three directions periodic: try the naive form
  f = solve(cell^T, v^T)^T ; f -= floor(f + 0.5) ; v' = f @ cell
  and keep it iff every |v'| < 0.5 * min(cell lengths);
* otherwise Minkowski-reduce the cell, wrap v into the reduced cell
  (fractional %= 1.0) and take the shortest of the images
  v + (h,k,l) @ rcell, (0,0,0) tried first and then
  itertools.product([-1,0,1], repeat=3) order (first minimum wins).

PARITY UNPINNED: there is no real ASE here and the reference has no test of
its periodic path, so this restatement is checked only
against a brute-force image search (tests/test_oracle_mic.py).
Partial periodicity (a False entry in pbc): find_mic then always takes the general
branch; minkowski_reduce Gauss-reduces the two periodic vectors when two directions
are periodic (permute periodic first, reduce, undo the permutation, keep the
handedness) and leaves the cell alone when one is; wrap_positions wraps the periodic
fractional coordinates only and the image search runs over the periodic directions
only.  Checked against a brute-force search over the periodic images.
"""
import itertools

import numpy as np

MAX_IT = 5000


class _CycleChecker:
    def __init__(self, d):
        n = {2: 6, 3: 12}[d]
        max_cycle_length = int(np.prod([n - i for i in range(d)]) * d)
        self.visited = np.zeros((max_cycle_length, 3 * d), dtype=int)

    def add_site(self, H):
        H = np.asarray(H).ravel()
        found = (self.visited == H).all(axis=1).any()
        self.visited = np.roll(self.visited, 1, axis=0)
        self.visited[0] = H
        return bool(found)


def _reduction_gauss(B, hu, hv):
    cyc = _CycleChecker(d=2)
    u = hu @ B
    v = hv @ B
    for _ in range(MAX_IT):
        x = int(round(np.dot(u, v) / np.dot(u, u)))
        hu, hv = hv - x * hu, hu
        u = hu @ B
        v = hv @ B
        site = np.array([hu, hv])
        if np.dot(u, u) >= np.dot(v, v) or cyc.add_site(site):
            return hv, hu
    raise RuntimeError("Gaussian basis not found")


def _relevant_vectors_2d(u, v):
    cs = np.array(list(itertools.product([-1, 0, 1], repeat=2)))
    vs = cs @ np.array([u, v])
    indices = np.argsort(np.linalg.norm(vs, axis=1))[:7]
    return vs[indices], cs[indices]


def _closest_vector(t0, u, v):
    t = t0
    a = np.zeros(2, dtype=int)
    rs, cs = _relevant_vectors_2d(u, v)
    dprev = float("inf")
    for _ in range(MAX_IT):
        ds = np.linalg.norm(rs + t, axis=1)
        index = int(np.argmin(ds))
        if index == 0 or ds[index] >= dprev:
            return a
        dprev = ds[index]
        r = rs[index]
        kopt = int(round(-np.dot(t, r) / np.dot(r, r)))
        a = a + kopt * cs[index]
        t = t0 + a[0] * u + a[1] * v
    raise RuntimeError("Closest vector not found")


def _reduction_full(B):
    cyc = _CycleChecker(d=3)
    H = np.eye(3, dtype=int)
    norms = np.linalg.norm(B, axis=1)
    for _ in range(MAX_IT):
        H = H[np.argsort(norms, kind="merge")]
        hw = H[2]
        hu, hv = _reduction_gauss(B, H[0], H[1])
        H = np.array([hu, hv, hw])
        R = H @ B
        u, v, _w = R
        X = u / np.linalg.norm(u)
        Y = v - X * np.dot(v, X)
        Y /= np.linalg.norm(Y)
        pu, pv, pw = R @ np.array([X, Y]).T
        nb = _closest_vector(pw, pu, pv)
        H[2] = np.array([nb[0], nb[1], 1]) @ H
        R = H @ B
        norms = np.diag(R @ R.T)
        if norms[2] >= norms[1] or cyc.add_site(H):
            return R, H
    raise RuntimeError("Reduced basis not found")


def minkowski_reduce(cell, pbc=True):
    """Return (reduced_cell, op) with reduced_cell = op @ cell; only the periodic directions are reduced."""
    cell = np.asarray(cell, dtype=float)
    pbc = _pbc3(pbc)
    dim = int(pbc.sum())
    op = np.eye(3, dtype=int)
    if dim == 2:
        perm = np.argsort(pbc, kind="merge")[::-1]  # periodic directions first
        pcell = cell[perm][:, perm]
        norms = np.linalg.norm(pcell, axis=1)
        norms[2] = float("inf")
        op = op[np.argsort(norms)]
        hu, hv = _reduction_gauss(pcell, op[0], op[1])
        op[0], op[1] = hu, hv
        invperm = np.argsort(perm)
        op = op[invperm][:, invperm]
        index = int(np.argmin(pbc))  # the non-periodic direction: handedness judged with the plane normal in its place
        normal = np.cross(cell[index - 2], cell[index - 1])
        normal /= np.linalg.norm(normal)
        c0, c1 = cell.copy(), op @ cell
        c0[index], c1[index] = normal, normal
        if np.sign(np.linalg.det(c0)) != np.sign(np.linalg.det(c1)):
            op[index - 1] *= -1
    elif dim == 3:
        _, op = _reduction_full(cell)
        if np.sign(np.linalg.det(cell)) != np.sign(np.linalg.det(op @ cell)):
            op = -op
    norms1 = np.sort(np.linalg.norm(cell, axis=1))
    norms2 = np.sort(np.linalg.norm(op @ cell, axis=1))
    if (norms2 > norms1 + 1e-12).any():
        raise RuntimeError("Minkowski reduction failed")
    return op @ cell, op


def _pbc3(pbc):
    out = np.empty(3, dtype=bool)
    out[:] = pbc
    return out


def wrap_positions_eps0(positions, cell, pbc=True):
    """ase.geometry.wrap_positions(positions, cell, pbc, eps=0): periodic fractional coordinates into [0,1)."""
    pbc = _pbc3(pbc)
    fractional = np.linalg.solve(cell.T, np.asarray(positions).T).T
    for i in range(3):
        if pbc[i]:
            fractional[:, i] %= 1.0
    return np.dot(fractional, cell)


def naive_find_mic(v, cell):
    f = np.linalg.solve(cell.T, np.transpose(v)).T
    f -= np.floor(f + 0.5)
    vmin = f @ cell
    vlen = np.linalg.norm(vmin, axis=1)
    return vmin, vlen


def general_find_mic(v, cell, pbc=True):
    pbc = _pbc3(pbc)
    rcell, _ = minkowski_reduce(cell, pbc)
    positions = wrap_positions_eps0(v, rcell, pbc)
    ranges = [np.arange(-1 * p, p + 1) for p in pbc.astype(int)]
    hkls = np.array([(0, 0, 0)] + list(itertools.product(*ranges)))
    vrvecs = hkls @ rcell
    x = positions + vrvecs[:, None]
    lengths = np.linalg.norm(x, axis=2)
    indices = np.argmin(lengths, axis=0)
    vmin = x[indices, np.arange(len(positions)), :]
    vlen = lengths[indices, np.arange(len(positions))]
    return vmin, vlen


def find_mic(v, cell, pbc=True):
    cell = np.asarray(cell, dtype=float)
    pbc = cell.any(1) & _pbc3(pbc)
    dim = int(np.sum(pbc))
    v = np.asarray(v, dtype=float)
    single = v.ndim == 1
    v = np.atleast_2d(v)

    if dim > 0:
        naive_ok = False
        if dim == 3:
            vmin, vlen = naive_find_mic(v, cell)
            if (vlen < 0.5 * min(np.linalg.norm(cell, axis=1))).all():
                naive_ok = True
        if not naive_ok:
            vmin, vlen = general_find_mic(v, cell, pbc=pbc)
    else:
        vmin = v.copy()
        vlen = np.linalg.norm(vmin, axis=1)

    if single:
        return vmin[0], vlen[0]
    return vmin, vlen
