"""CPU: the restated ase.geometry.find_mic (oracle/ase_mic.py, parity unpinned -- no real ASE here)
against a brute-force image search."""
import numpy as np
import pytest

from oracle.mbgdml_oracle import ase_mic

CELLS = {
    "cubic": np.eye(3) * 12.4,
    "ortho": np.diag([9.0, 11.5, 14.2]),
    "triclinic": np.array([[12.6, 0.0, 0.0], [2.8, 12.1, 0.0], [-1.9, 3.1, 11.8]]),
    "skewed": np.array([[10.0, 0.0, 0.0], [7.0, 8.0, 0.0], [3.0, -5.0, 9.0]]),
}


def brute(v, cell, n=4):
    best = np.full(len(v), np.inf)
    for h in range(-n, n + 1):
        for k in range(-n, n + 1):
            for l in range(-n, n + 1):
                best = np.minimum(best, np.linalg.norm(v + np.array([h, k, l]) @ cell, axis=1))
    return best


@pytest.mark.parametrize("name", list(CELLS))
def test_find_mic_is_the_minimum_image(name):
    cell = CELLS[name]
    rng = np.random.default_rng(3)
    v = rng.uniform(-1.5, 1.5, size=(300, 3)) @ cell
    vmin, vlen = ase_mic.find_mic(v, cell)
    assert np.allclose(np.linalg.norm(vmin, axis=1), vlen)
    assert np.allclose(vlen, brute(v, cell), atol=1e-10)
    # the result differs from the input by a lattice vector
    frac = np.linalg.solve(cell.T, (vmin - v).T).T
    assert np.allclose(frac, np.round(frac), atol=1e-9)


def test_minkowski_reduced_cell_spans_same_lattice():
    for cell in CELLS.values():
        rcell, op = ase_mic.minkowski_reduce(cell)
        assert abs(abs(np.linalg.det(op)) - 1) < 1e-12
        assert np.allclose(op @ cell, rcell)
        assert np.all(np.sort(np.linalg.norm(rcell, axis=1)) <= np.sort(np.linalg.norm(cell, axis=1)) + 1e-12)


@pytest.mark.parametrize("name", ["cubic", "ortho", "triclinic"])
@pytest.mark.parametrize("pbc", [(True, True, False), (True, False, True), (False, True, False), (False, False, False)])
def test_find_mic_partial_periodicity(name, pbc):
    """A False entry in pbc: the minimum over the images of the periodic directions only."""
    cell = CELLS[name]
    rng = np.random.default_rng(5)
    v = rng.uniform(-1.5, 1.5, size=(200, 3)) @ cell
    vmin, vlen = ase_mic.find_mic(v, cell, pbc=pbc)
    best = np.full(len(v), np.inf)
    rng_ = [range(-4, 5) if p else range(1) for p in pbc]
    for h in rng_[0]:
        for k in rng_[1]:
            for l in rng_[2]:
                best = np.minimum(best, np.linalg.norm(v + np.array([h, k, l]) @ cell, axis=1))
    assert np.allclose(vlen, best, atol=1e-10) and np.allclose(np.linalg.norm(vmin, axis=1), vlen)
    frac = np.linalg.solve(cell.T, (vmin - v).T).T
    assert np.allclose(frac, np.round(frac), atol=1e-9)
    assert not np.any(np.round(frac)[:, ~np.array(pbc)])  # no translation along a non-periodic direction
