"""Matern-5/2 covariances (analysis.models.gdml_mat52) and the force-field kernel matrix
(GDMLTrain._assemble_kernel_mat).  tests/golden/train_kernel.npz holds outputs of the live reference
(oracle/make_golden.py:case_train_kernel).

CPU tests pin the oracle to those outputs; GPU tests compare the CUDA path (through the C ABI) with the golden
outputs, with the oracle on other seeded inputs, and -- at sizes the oracle cannot reach -- through the property
that ties the kernel matrix to the prediction kernels: (K alpha)_i = F_raw(R_i) when R_d_desc_alpha is built
from alpha (train.py:881-884, predict.py), plus symmetry.  Tolerance 1e-9 relative (max|delta| / max|ref|).
"""
import numpy as np
import pytest

from helpers import load, relerr, unpack_models
from mbgdml_b200 import synthetic
from oracle import mbgdml_oracle as orc

TOL = 1e-9
CASES = ["h2o", "h2o2", "h2o3", "meoh"]


def _train_inputs(comp, T, seed):
    rng = np.random.default_rng(seed)
    Rt = synthetic.fragment_geometries(rng, comp, T)
    n = Rt.shape[1]
    D = n * (n - 1) // 2
    X, G = np.zeros((T, D)), np.zeros((T, D, 3))
    for t in range(T):
        X[t], G[t] = orc.from_r(Rt[t].reshape(-1))
    perms = synthetic.fragment_perms(comp)
    return Rt, X, G, perms, orc.tril_perms_lin_from_perms(perms)


# ---------------------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("name", CASES)
def test_oracle_kernel_mat_matches_reference(name):
    g = load("train_kernel.npz")
    for use_E in (0, 1):
        K = orc.assemble_kernel_mat(g[f"{name}_R_desc"], g[f"{name}_R_d_desc"], g[f"{name}_tril_perms_lin"],
                                    float(g[f"{name}_sig"]), use_E_cstr=bool(use_E))
        ref = g[f"{name}_K_E{use_E}"]
        assert K.shape == ref.shape
        n3 = 3 * g[f"{name}_R"].shape[1] * g[f"{name}_R"].shape[0]
        assert relerr(K[:n3, :n3], ref[:n3, :n3]) <= 1e-11  # force-force blocks on their own scale
        assert relerr(K, ref) <= 1e-11


def test_oracle_kernel_mat_reference_fixture():
    """The reference's own known answer: tests/data/train/1h2o/K.npy (sigma 42, first 20 training points)."""
    g = load("train_kernel.npz")
    K = orc.assemble_kernel_mat(g["fix1h2o_R_desc"], g["fix1h2o_R_d_desc"], g["fix1h2o_tril_perms_lin"], 42.0)
    assert K.shape == g["fix1h2o_K"].shape == (180, 180)
    assert relerr(K, g["fix1h2o_K"]) <= 1e-13
    x, gj = orc.from_r(g["fix1h2o_R"][3].reshape(-1))  # descriptors of the fixture come from its geometries
    assert np.array_equal(x, g["fix1h2o_R_desc"][3]) and np.array_equal(gj, g["fix1h2o_R_d_desc"][3])


def test_oracle_mat52_matches_reference():
    g = load("train_kernel.npz")
    for k in range(3):
        md = {kk[len(f"mat52_{k}_"):]: v for kk, v in g.items() if kk.startswith(f"mat52_{k}_") and "_R" not in kk[7:]
              and "_out" not in kk[7:]}
        md = {**md, "R_desc": g[f"mat52_{k}_R_desc"], "R_d_desc_alpha": g[f"mat52_{k}_R_d_desc_alpha"]}
        m = orc.Model(md)
        for tag in ("1", "5"):
            got = orc.gdml_mat52(m, md["z"], g[f"mat52_{k}_R{tag}"])
            assert got.shape == g[f"mat52_{k}_out{tag}"].shape
            assert relerr(got, g[f"mat52_{k}_out{tag}"]) <= 1e-13


def test_kernel_mat_times_alpha_is_the_predicted_force_oracle():
    """The property the full-size GPU test relies on, checked on the oracle first."""
    comp, T, sig = ["h2o", "h2o"], 6, 20.0
    _, X, G, perms, tpl = _train_inputs(comp, T, 5)
    n = 6
    K = orc.assemble_kernel_mat(X, G, tpl, sig)
    assert relerr(K, K.T) <= 1e-13
    rng = np.random.default_rng(6)
    alpha = rng.normal(size=(T, n, 3))
    i, j = np.tril_indices(n, -1)
    RdA = np.einsum("tkc,tkc->tk", G, alpha[:, j, :] - alpha[:, i, :])  # Desc.d_desc_dot_vec (desc.py:341-358)
    md = {"z": np.zeros(n, int), "r_unit": np.array("Angstrom"), "comp_ids": np.array(comp), "sig": sig,
          "R_desc": X.T.copy(), "R_d_desc_alpha": RdA, "c": 0.0, "std": 1.0, "perms": perms, "tril_perms_lin": tpl}
    m = orc.Model(md)
    Ka = (K @ alpha.reshape(-1)).reshape(T, -1)
    for t in range(T):
        out = orc.predict_gdml_wkr(X[t], G[t], n, sig, len(perms), m.R_desc_perms, m.R_d_desc_alpha_perms)
        assert relerr(out[1:], Ka[t]) <= 1e-12


# ---------------------------------------------------------------------------------------- GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_kernel_mat_golden(name):
    from mbgdml_b200._gdml.train import GDMLTrain, assemble_kernel_mat

    g = load("train_kernel.npz")
    args = (g[f"{name}_R_desc"], g[f"{name}_R_d_desc"], g[f"{name}_tril_perms_lin"], float(g[f"{name}_sig"]))
    n3 = 3 * g[f"{name}_R"].shape[1] * g[f"{name}_R"].shape[0]
    for use_E in (0, 1):
        K = assemble_kernel_mat(*args, use_E_cstr=bool(use_E))
        ref = g[f"{name}_K_E{use_E}"]
        assert K.shape == ref.shape
        assert relerr(K[:n3, :n3], ref[:n3, :n3]) <= TOL
        if use_E:  # energy rows / columns / energy-energy block, each on its own scale
            assert relerr(K[n3:, :n3], ref[n3:, :n3]) <= TOL
            assert relerr(K[:n3, n3:], ref[:n3, n3:]) <= TOL
            assert relerr(K[n3:, n3:], ref[n3:, n3:]) <= TOL
    K2 = GDMLTrain()._assemble_kernel_mat(*args, None, use_E_cstr=False, alloc_extra_rows=3)
    assert K2.shape == (n3 + 3, n3) and np.array_equal(K2[:n3], assemble_kernel_mat(*args))


@pytest.mark.gpu
def test_gpu_kernel_mat_column_subsets_golden():
    """`col_idxs` (train.py:1105-1115, 1187-1245) against the live reference's outputs (tests/golden/train_kernel_cols.npz,
    oracle/make_golden.py:case_train_kernel_cols): a leading slice of training points, sorted index arrays without and
    with energy-constraint columns, `alloc_extra_rows`; and, for every form incl. the one the reference itself fails
    on (a slice of training points that does not start at 0), the defining property K_sub == K_full[:, cols]."""
    from mbgdml_b200._gdml.train import assemble_kernel_mat

    g = load("train_kernel_cols.npz")
    args = (g["R_desc"], g["R_d_desc"], g["tril_perms_lin"], float(g["sig"]))
    T, D = g["R_desc"].shape
    dim_i = 3 * int((1 + np.sqrt(8 * D + 1)) / 2)
    for name in ("lead", "idx", "idxE", "leadE"):
        cols = np.s_[int(g[f"{name}_slice"][0]):int(g[f"{name}_slice"][1])] if f"{name}_slice" in g else g[f"{name}_cols"]
        use_E, extra = bool(g[f"{name}_use_E"]), int(g[f"{name}_extra"])
        K = assemble_kernel_mat(*args, use_E_cstr=use_E, col_idxs=cols, alloc_extra_rows=extra)
        ref = g[f"{name}_K"]
        assert K.shape == (ref.shape[0] + extra, ref.shape[1]), name
        assert relerr(K[: ref.shape[0]], ref) <= TOL, name
        if extra:
            assert not K[ref.shape[0]:].any()
    K_full = assemble_kernel_mat(*args)
    for cols in (np.s_[2 * dim_i: 5 * dim_i], np.s_[5:40:3], np.array([0, 17, 18, 35, 36, T * dim_i - 1]), np.s_[: dim_i]):
        assert relerr(assemble_kernel_mat(*args, col_idxs=cols), K_full[:, cols]) <= 1e-12
    with pytest.raises(ValueError):
        assemble_kernel_mat(*args, col_idxs=np.arange(T * dim_i + 1))
    with pytest.raises(AssertionError):
        assemble_kernel_mat(*args, col_idxs=np.array([5, 3]))


@pytest.mark.gpu
def test_gpu_torch_assemble_front_end():
    """``GDMLTorchAssemble`` (torchtools.py:72-398) as ``GDMLTrain._assemble_kernel_mat`` drives it (train.py:1190-1288):
    sequential column blocks incl. energy-constraint columns in two ``forward`` calls, and the tuple form of a
    "fancy" column subset, each against the columns of the full matrix."""
    import torch

    from mbgdml_b200._gdml.torchtools import GDMLTorchAssemble
    from mbgdml_b200._gdml.train import assemble_kernel_mat

    g = load("train_kernel_cols.npz")
    R_desc, R_d_desc, tpl, sig = g["R_desc"], g["R_d_desc"], g["tril_perms_lin"], float(g["sig"])
    T, D = R_desc.shape
    dim_i = 3 * int((1 + np.sqrt(8 * D + 1)) / 2)
    for use_E in (False, True):
        K_full = assemble_kernel_mat(R_desc, R_d_desc, tpl, sig, use_E_cstr=use_E)
        n_rows = K_full.shape[0]
        J = list(range(T + (T if use_E else 0)))
        out, done = np.full((n_rows, n_rows), np.nan), []
        asm = GDMLTorchAssemble(J, tpl, sig, use_E, torch.from_numpy(R_desc), torch.from_numpy(R_d_desc), out=out,
                                callback=done.append)
        asm.forward(torch.arange(0, len(J), 2))
        assert np.isnan(out).any()
        asm.forward(torch.arange(1, len(J), 2))
        assert relerr(out, K_full) <= 1e-12 and sum(done) == n_rows
        # fancy indexing: partials 1, 4 of point 0, all of point 2, partial 7 of point 5 (+ energy column 3)
        cols = np.concatenate([[1, 4], 2 * dim_i + np.arange(dim_i), [5 * dim_i + 7], [T * dim_i + 3] if use_E else []]).astype(int)
        J = [(0, 0, [1, 4]), (2, 2, list(range(dim_i))), (2 + dim_i, 5, [7])] + ([(3 + dim_i, T * dim_i + 3, [0])] if use_E else [])
        out = torch.full((n_rows, len(cols)), float("nan"), dtype=torch.float64, device="cuda")
        GDMLTorchAssemble(J, tpl, sig, use_E, torch.from_numpy(R_desc).cuda(), torch.from_numpy(R_d_desc).cuda(), out=out).forward(range(len(J)))
        assert relerr(out.cpu().numpy(), K_full[:, cols]) <= 1e-12


@pytest.mark.gpu
def test_gpu_kernel_mat_reference_fixture():
    """CUDA path vs. the reference's own tests/data/train/1h2o/K.npy."""
    from mbgdml_b200._gdml.train import assemble_kernel_mat

    g = load("train_kernel.npz")
    K = assemble_kernel_mat(g["fix1h2o_R_desc"], g["fix1h2o_R_d_desc"], g["fix1h2o_tril_perms_lin"], 42.0)
    assert relerr(K, g["fix1h2o_K"]) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("comp,T,sig", [(["meoh", "h2o"], 5, 40.0), (["meoh", "meoh"], 4, 60.0),
                                        (["meoh", "meoh", "meoh"], 3, 150.0), (["h2o"], 33, 3.0)])
def test_gpu_kernel_mat_vs_oracle(comp, T, sig):
    """Other fragment shapes (9, 12 and 18 atoms: 18, 36 and 100 permutations -> several permutation chunks)."""
    from mbgdml_b200._gdml.train import assemble_kernel_mat

    _, X, G, _, tpl = _train_inputs(comp, T, 11)
    n3 = T * 3 * ((1 + int(np.sqrt(8 * X.shape[1] + 1))) // 2)
    for use_E in (False, True):
        K = assemble_kernel_mat(X, G, tpl, sig, use_E_cstr=use_E)
        ref = orc.assemble_kernel_mat(X, G, tpl, sig, use_E_cstr=use_E)
        assert relerr(K[:n3, :n3], ref[:n3, :n3]) <= TOL
        if use_E:
            assert relerr(K[n3:, :n3], ref[n3:, :n3]) <= TOL and relerr(K[:n3, n3:], ref[:n3, n3:]) <= TOL
            assert relerr(K[n3:, n3:], ref[n3:, n3:]) <= TOL


@pytest.mark.gpu
def test_gpu_kernel_mat_full_size_property():
    """T = 256 water trimers (K is 6912 x 6912, 65 536 blocks): symmetry, and K alpha equals the forces the
    prediction kernels give at the training geometries for the model built from alpha."""
    from mbgdml_b200._gdml.train import assemble_kernel_mat
    from mbgdml_b200.gdml_predict import GDMLPredict

    comp, T, sig = ["h2o", "h2o", "h2o"], 256, 100.0
    Rt, X, G, perms, tpl = _train_inputs(comp, T, 21)
    n = 9
    K = assemble_kernel_mat(X, G, tpl, sig)
    assert K.shape == (T * 27, T * 27)
    assert relerr(K, K.T) <= TOL
    rng = np.random.default_rng(22)
    alpha = rng.normal(size=(T, n, 3))
    i, j = np.tril_indices(n, -1)
    RdA = np.einsum("tkc,tkc->tk", G, alpha[:, j, :] - alpha[:, i, :])
    md = {"z": np.concatenate([synthetic.MONOMERS[c][0] for c in comp]), "r_unit": np.array("Angstrom"), "sig": sig,
          "R_desc": X.T.copy(), "R_d_desc_alpha": RdA, "c": 0.0, "std": 1.0, "perms": perms, "tril_perms_lin": tpl}
    _, F = GDMLPredict(md).predict(Rt.reshape(T, -1))
    assert relerr(F, (K @ alpha.reshape(-1)).reshape(T, -1)) <= TOL


@pytest.mark.gpu
def test_gpu_mat52_golden_and_oracle():
    import mbgdml_b200 as mb
    from mbgdml_b200.analysis import gdml_mat52, gdml_mat52_wrk

    g = load("train_kernel.npz")
    for k in range(3):
        md = {kk[len(f"mat52_{k}_"):]: v for kk, v in g.items() if kk.startswith(f"mat52_{k}_")}
        model = mb.gdmlModel({kk: v for kk, v in md.items() if not kk.startswith(("R1", "R5", "out"))})
        for tag in ("1", "5"):
            got = gdml_mat52(model, md["z"], md[f"R{tag}"])
            assert got.shape == md[f"out{tag}"].shape
            assert relerr(got, md[f"out{tag}"]) <= TOL
        # the worker form on the expanded rows
        om = orc.Model({kk: v for kk, v in md.items() if not kk.startswith(("R1", "R5", "out"))})
        x, _ = orc.from_r(md["R1"].reshape(-1))
        assert relerr(gdml_mat52_wrk(x, md["sig"], om.n_perms, om.R_desc_perms), md["out1"]) <= TOL
    # wide fragment, ragged training-set size, many structures (oracle on the same inputs)
    md = synthetic.make_model(["meoh", "meoh", "h2o"], n_train=77, sig=250.0, seed=3)
    R = synthetic.fragment_geometries(np.random.default_rng(4), ["meoh", "meoh", "h2o"], 19)
    got = gdml_mat52(mb.gdmlModel(dict(md)), md["z"], R)
    assert relerr(got, orc.gdml_mat52(orc.Model(dict(md)), md["z"], R)) <= TOL


@pytest.mark.gpu
def test_gpu_kernel_mat_permutation_set_not_closed_under_inversion():
    """A permutation list that is not a group (identity + one 3-cycle of the molecules, without its inverse): K is
    no longer symmetric and the kernel must take the all-ordered-pairs path instead of mirroring blocks."""
    from mbgdml_b200._gdml.train import assemble_kernel_mat

    comp, T, sig = ["h2o", "h2o", "h2o"], 5, 60.0
    _, X, G, perms, _ = _train_inputs(comp, T, 31)
    ident = np.arange(9)
    cyc = next(p for p in perms if not np.array_equal(p[p], ident) and not np.array_equal(p, ident))
    tpl = orc.tril_perms_lin_from_perms(np.array([ident, cyc]))
    for use_E in (False, True):
        K = assemble_kernel_mat(X, G, tpl, sig, use_E_cstr=use_E)
        ref = orc.assemble_kernel_mat(X, G, tpl, sig, use_E_cstr=use_E)
        n3 = T * 27
        assert relerr(ref[:n3, :n3], ref[:n3, :n3].T) > 1e-6  # the case really is asymmetric
        assert relerr(K[:n3, :n3], ref[:n3, :n3]) <= TOL
        if use_E:
            assert relerr(K[n3:, :n3], ref[n3:, :n3]) <= TOL and relerr(K[:n3, n3:], ref[:n3, n3:]) <= TOL
            assert relerr(K[n3:, n3:], ref[n3:, n3:]) <= TOL
