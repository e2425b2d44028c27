"""Load balance of the fragment sharding, measured on ONE GPU: every shard of an 8-rank evaluation of the weak-scaling
bench shape (box140, 64 frames) and of the strong-scaling shape (one 64-H2O structure, 41 664 trimers) is evaluated
in turn with tiny models; the accepted fragments per shard are what each rank would contract.
   python tools/shard_balance.py > gpurun_out/shard_balance.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mbgdml_b200 as mb  # noqa: E402
from mbgdml_b200 import _engine, _native, synthetic  # noqa: E402


def shards(ctx, Z, R, eids, cids, model, cell, world):
    pred = mb.mbePredict([model], mb.predict_gdml, periodic_cell=cell)
    acc = []
    for r in range(world):
        src = [mb.gen_combs(pred.get_avail_entities(cids, model.comp_ids))]
        _, _, _, st = _engine.predict_total(ctx, Z, R, eids, [model], src, [[]], cell, shard=(r, world))
        acc.append(int(st[0][2]))
    return acc


def main():
    ctx = _native.get_context()
    md = synthetic.make_model(["h2o"] * 3, n_train=4, sig=400.0, seed=3)
    out = []
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(140, 16.12, 64, seed=0)
    model = mb.gdmlModel(dict(md), criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": md["entity_ids"]}, 10.0))
    for kind in ("pairlist", "full"):
        ctx.set_culling(kind)
        a = shards(ctx, Z, R, eids, cids, model, mb.Cell(cell_v, pbc=True), 8)
        out.append({"case": f"box140 x 64 frames, 8 shards, culling {kind}", "accepted_per_shard": a,
                    "max_over_mean": max(a) / np.mean(a)})
    ctx.set_culling("auto")
    Zc, Rc, ec, cc = synthetic.make_cluster(64, seed=0)
    a = shards(ctx, Zc, Rc[None], ec, cc, mb.gdmlModel(dict(md)), None, 8)
    out.append({"case": "64-H2O cluster, 41 664 trimers, 8 shards", "accepted_per_shard": a, "max_over_mean": max(a) / np.mean(a)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
