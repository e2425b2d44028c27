"""The contraction launch ONE rank of an 8-GPU strong-scaling run sees (BASELINE config 3: one 64-H2O structure, rank r of
G culls and contracts its granules of the 41 664 trimers), on a single GPU -- for `ncu` and for timing the shard shape.
   python tools/shard_shape.py [ranks=8] [rank=0] > gpurun_out/shard_shape.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mbgdml_b200 as mb  # noqa: E402
from mbgdml_b200 import _engine, _native  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ctx = _native.get_context(0)
    ws = bench.make_workload("trimers64", 1, 1000)
    models = [mb.gdmlModel(dict(d)) for d in ws["model_dicts"]]
    pred = mb.mbePredict(models, mb.predict_gdml)
    R = torch.from_numpy(ws["R"]).cuda()
    out = torch.zeros(1 + 3 * len(ws["Z"]), dtype=torch.float64, device="cuda")
    recs = []
    for _ in range(6):
        _engine.predict_total(ctx, ws["Z"], R, ws["eids"], models,
                              [mb.gen_combs(pred.get_avail_entities(ws["cids"], m.comp_ids)) for m in models], [[]], None,
                              shard=(rank, world), device_out=(out[:1], out[1:], None))
        torch.cuda.synchronize()
        recs.append([(r["n_frag_atoms"], r["n_fragments"], round(r["ms"], 4)) for r in ctx.last_launches()])
    print(json.dumps({"ranks": world, "rank": rank, "launches_per_call": recs}))


if __name__ == "__main__":
    main()
