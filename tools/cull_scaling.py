"""Candidate generation against system size: the full enumeration (k_cull over all C(n, 3) tuples) and the pair list
(csrc/pairlist.cuh) on water boxes of growing size at the density of the benchmark box, one frame, physical cut-off
8 A, 2-body < 6.0 and 3-body < 10.0 (com_distance_sum), T = 1000.  Device time of everything that is not the
contraction (`ms_other`: table upload, pair list, cull, scan, compact) and of the contraction, per call.
   python tools/cull_scaling.py [n_mol ...] > gpurun_out/cull_scaling.json"""
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mbgdml_b200 as mb  # noqa: E402
from mbgdml_b200 import _engine, _native, synthetic  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [140, 512, 1000, 2000]
    ctx = _native.get_context()
    dicts = [synthetic.make_model(["h2o"] * k, n_train=1000, sig=[100.0, 400.0][k - 2], seed=k) for k in (2, 3)]
    models = [mb.gdmlModel(dict(d), criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": d["entity_ids"]}, c))
              for d, c in zip(dicts, [6.0, 10.0])]
    out = []
    for n in sizes:
        box = 16.12 * (n / 140) ** (1 / 3)
        Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=n, box=box, n_frames=1, seed=0, frame_sigma=0.1)
        cell = mb.Cell(cell_v, cutoff=min(8.0, box / 2), pbc=True)
        pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
        rec = {"n_mol": n, "box_A": round(box, 3), "cutoff_A": float(cell.cutoff), "candidates": [math.comb(n, 2), math.comb(n, 3)]}
        ref = None
        for kind in ("full", "pairlist"):
            ctx.set_culling(kind)
            other, contract = [], []
            for _ in range(4):
                srcs = [mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)) for m in models]
                E, F, _, stats = _engine.predict_total(ctx, Z, R, eids, models, srcs, [[], []], cell)
                t = ctx.last_timing()
                other.append(t["ms_other"]), contract.append(t["ms_contract"])
            rec["accepted"] = [s[2] for s in stats]
            rec[f"ms_other_{kind}"] = round(float(np.min(other[1:])), 4)
            rec[f"ms_contract_{kind}"] = round(float(np.min(contract[1:])), 4)
            if ref is None:
                ref = (E, F, [s[2] for s in stats])
            else:
                assert ref[2] == [s[2] for s in stats]
                rec["max_rel_dF_pairlist_vs_full"] = float(np.max(np.abs(F - ref[1])) / np.max(np.abs(ref[1])))
        ctx.set_culling("auto")
        out.append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
