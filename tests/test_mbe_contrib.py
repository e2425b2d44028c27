"""MBE recombination (reference mbgdml/mbe.py:73-508: mbe_worker, mbe_contrib, decomp_to_total), pinned by the
reference's tests/test_mbe.py:38-100 known answer and by outputs of the live reference (tests/golden/
mbe_contrib_4h2o.npz, written by oracle/make_golden.py).  Bit-exact: the device sums in the reference's order."""
import numpy as np
import pytest

from helpers import load
from oracle import mbgdml_oracle as orc

E_KNOWN = np.array([-305.30075391, -305.29900387, -305.29489548])  # reference tests/test_mbe.py:99


def lower_args(g, k, keep=None):
    sel = slice(None) if keep is None else keep
    return (g[f"low{k}_E"][sel], g[f"low{k}_G"][sel], g[f"low{k}_entity_ids"],
            {int(g[f"low{k}_r_prov_id"][0]): "x" if keep is not None else "345767a1c57d8d2a4e0314d924b547ad"},
            g[f"low{k}_r_prov_specs"][sel])


def run_cases(mod, g):
    E, G = np.zeros(g["E_true"].shape), np.zeros(g["G_true"].shape)
    for k in range(3):
        E, G = mod.mbe_contrib(E, G, g["entity_ids"], None, None, *lower_args(g, k), operation="add")
        assert np.array_equal(E, g[f"add{k}_E"]) and np.array_equal(G, g[f"add{k}_G"])
    assert np.allclose(E, E_KNOWN)
    E, G = mod.mbe_contrib(g["E_true"].copy(), g["G_true"].copy(), g["entity_ids"], {0: "x"}, g["rm_specs"],
                           *lower_args(g, 1, g["rm_keep"]), operation="remove")
    assert np.array_equal(E, g["rm_E"]) and np.array_equal(G, g["rm_G"])
    E, F = mod.decomp_to_total(g["d2t_E_nbody"].copy(), g["d2t_F_nbody"].copy(), g["d2t_entity_ids"], g["d2t_combs"])
    assert np.array_equal(E, g["d2t_E"]) and np.array_equal(F, g["d2t_F"])
    with pytest.raises(ValueError):
        mod.mbe_contrib(E, F, g["d2t_entity_ids"], None, None, *lower_args(g, 0), operation="scale")


def test_oracle_reproduces_reference():
    run_cases(orc, load("mbe_contrib_4h2o.npz"))


def test_host_join_matches_linear_scan():
    """_mbe_pairs (sort join) finds exactly the rows the reference's np.all(...) scan finds, in loop order,
    including duplicated lower rows (first occurrence wins) and missing fragments."""
    from mbgdml_b200.mbe import _mbe_pairs

    rng = np.random.default_rng(0)
    n_ent, n_low = 6, 3
    specs = np.array([[rng.integers(0, 2), i] + sorted(rng.choice(40, n_ent, replace=False).tolist()) for i in range(7)])
    import itertools
    rows = []
    for s in specs:
        for comb in itertools.combinations(s[2:], n_low):
            if rng.random() < 0.7:
                rows.append([s[0], s[1], *comb])
    rows = np.array(rows + rows[:5])  # duplicates: the first occurrence must win
    rows = rows[rng.permutation(len(rows))]
    eids_low = np.repeat(np.arange(n_low), 2)
    r_idxs = [5, 0, 3, 6]
    off, low_idx, slot_ent = _mbe_pairs(np.repeat(np.arange(n_ent), 2), specs, r_idxs, eids_low, rows)
    want_idx, want_slot, want_off = [], [], [0]
    for i_r in r_idxs:
        s = specs[i_r]
        for comb in itertools.combinations(s[2:], n_low):
            hits = np.where(np.all(rows == np.array([s[0], s[1], *comb]), axis=1))[0]
            if hits.size:
                want_idx.append(hits[0])
                want_slot.append([np.where(s[2:] == c)[0][0] for c in comb])
        want_off.append(len(want_idx))
    assert np.array_equal(off, want_off) and np.array_equal(low_idx, want_idx) and np.array_equal(slot_ent, want_slot)


@pytest.mark.gpu
def test_device_mbe_contrib_bit_exact():
    import mbgdml_b200.mbe as mbe

    run_cases(mbe, load("mbe_contrib_4h2o.npz"))
    with pytest.raises(NotImplementedError):
        g = load("mbe_contrib_4h2o.npz")
        mbe.mbe_contrib(np.zeros(3), np.zeros((3, 12, 3)), g["entity_ids"], None, None, *lower_args(g, 0), use_ray=True)


@pytest.mark.gpu
def test_decomp_plus_recombination_equals_total():
    """Reference tests/test_predictsets.py:40-77: predict_decomp + decomp_to_total == predict (6 H2O, 8 frames,
    criteria on), here on random-init models at the north_star tolerance."""
    import mbgdml_b200 as mb
    from mbgdml_b200 import synthetic
    from mbgdml_b200.mbe import decomp_to_total
    from helpers import b200_models, relerr

    g = load("ref_6h2o.npz")
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    fam = [["h2o"], ["h2o", "h2o"], ["h2o", "h2o", "h2o"]]
    dicts = [synthetic.make_model(f, n_train=120, sig=s, seed=k, c=0.3 * k, std=1.0 + k)
             for k, (f, s) in enumerate(zip(fam, (10.0, 100.0, 400.0)))]
    models = b200_models(dicts, [None, 6.0, 10.0])
    E, F = mb.mbePredict(models, mb.predict_gdml).predict(Z, R, eids, cids)
    # like data/predictset.py:77-94, the decomposed path is driven with the decomp predictor
    E_data, F_data, comb_data, orders = mb.mbePredict(models, mb.predict_gdml_decomp).predict_decomp(Z, R, eids, cids)
    assert orders == [1, 2, 3]
    E_sum, F_sum = np.zeros_like(E), np.zeros_like(F)
    for E_n, F_n, combs in zip(E_data, F_data, comb_data):
        E_t, F_t = decomp_to_total(E_n, F_n, eids, combs)
        E_sum += E_t
        F_sum += F_t
    assert relerr(E_sum, E) <= 1e-9 and relerr(F_sum, F) <= 1e-9
