"""Shared test helpers: golden fixtures -> oracle models / package models."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name), allow_pickle=True))


def unpack_models(g, prefix="m"):
    """Inverse of oracle/make_golden.py:pack_models."""
    n = int(g[f"{prefix}n"])
    out = []
    for i in range(n):
        pre = f"{prefix}{i}_"
        out.append({k[len(pre):]: g[k] for k in g if k.startswith(pre)})
    return out


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def oracle_models(dicts, cutoffs, desc="com", bounds=None):
    from oracle import mbgdml_oracle as orc

    out = []
    for k, (d, cut) in enumerate(zip(dicts, cutoffs)):
        bound = "upper" if bounds is None else bounds[k]
        if desc == "com":
            crit = orc.Criteria(orc.com_distance_sum, {"entity_ids": np.asarray(d["entity_ids"])}, cut, bound=bound)
        else:
            crit = orc.Criteria(orc.max_atom_pair_dist, {}, cut, bound=bound)
        out.append(orc.Model(dict(d), criteria=crit))
    return out


def b200_models(dicts, cutoffs, desc="com", bounds=None):
    import mbgdml_b200 as mb

    out = []
    for k, (d, cut) in enumerate(zip(dicts, cutoffs)):
        bound = "upper" if bounds is None else bounds[k]
        if desc == "com":
            crit = mb.Criteria(mb.com_distance_sum, {"entity_ids": np.asarray(d["entity_ids"])}, cut, bound=bound)
        else:
            crit = mb.Criteria(mb.max_atom_pair_dist, {}, cut, bound=bound)
        out.append(mb.gdmlModel(dict(d), criteria=crit))
    return out


def cutoffs_from(arr):
    return [None if np.isnan(c) else float(c) for c in np.asarray(arr, dtype=float)]
