"""Single-frame latency of the prediction path through the public API (host buffers in, host buffers out):
16-H2O cluster (BASELINE config 1, 696 fragments) and one frame of the 140-H2O periodic box (config 4).
   python tools/latency_check.py > gpurun_out/latency.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mbgdml_b200 as mb  # noqa: E402
from mbgdml_b200 import _native, synthetic  # noqa: E402


def run(name, Z, R, eids, cids, models, cell, scalers, reps=30):
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=scalers, periodic_cell=cell)
    ctx = _native.get_context()
    for _ in range(5):
        pred.predict(Z, R, eids, cids)
    wall, dev, launches = [], [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        E, F = pred.predict(Z, R, eids, cids)
        wall.append((time.perf_counter() - t0) * 1e3)
        t = ctx.last_timing()
        dev.append(t["ms_contract"] + t["ms_other"])
        launches = [(r["n_frag_atoms"], r["n_fragments"], round(r["ms"], 4)) for r in ctx.last_launches()]
    return {"case": name, "wall_ms_median": float(np.median(wall)), "wall_ms_min": float(np.min(wall)),
            "device_ms_median": float(np.median(dev)), "contraction_launches": launches, "E": float(np.ravel(E)[0])}


def main():
    out = []
    dicts = [synthetic.make_model(["h2o"] * k, n_train=1000, sig=[10.0, 100.0, 400.0][k - 1], seed=k) for k in (1, 2, 3)]
    Z, R, eids, cids = synthetic.make_cluster(16, seed=0)
    models = [mb.gdmlModel(dict(d)) for d in dicts]
    out.append(run("16-H2O cluster, 1+2+3-body, criteria None (16 + 120 + 560 fragments)", Z, R, eids, cids, models, None, []))
    Zb, Rb, eb, cb, cellv = synthetic.make_periodic_box(140, 16.12, 1, seed=0)
    cuts = [None, 6.0, 10.0]
    models = [mb.gdmlModel(dict(d), criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": d["entity_ids"]}, c))
              for d, c in zip(dicts, cuts)]
    cell = mb.Cell(cellv, pbc=True)
    sc = [mb.mbeAlchemyScale(0, 3, mb.linear_switching, {"factor": 0.5}), mb.mbeAlchemyScale(0, 2, mb.linear_switching, {"factor": 0.5})]
    out.append(run("140-H2O periodic box, one frame", Zb, Rb[0], eb, cb, models, cell, sc, reps=10))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
