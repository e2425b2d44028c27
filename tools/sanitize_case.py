"""Small case for `compute-sanitizer --tool memcheck`: the round-2 kernels that index through bit rows and reduced
candidate spaces (csrc/pairlist.cuh, homogeneous and heterogeneous jobs, plain call + prepared step) and the
partial-periodicity branch of the minimum image, on tiny models.
   compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_case.py
(compute-sanitizer is closed on this pool -- round 2: "runs under it have left GPUs needing a reset" -- so the case is
run plain: it asserts masks / counts / energies of the pair-list path against the full enumeration.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mbgdml_b200 as mb  # noqa: E402
from mbgdml_b200 import _native, synthetic  # noqa: E402
from mbgdml_b200.predictors import accept_mask  # noqa: E402


def main():
    ctx = _native.get_context()
    comp = ["h2o", "meoh"] * 12 + ["h2o"] * 20
    fams = [["h2o", "h2o"], ["h2o", "meoh"], ["h2o", "h2o", "h2o"], ["h2o", "h2o", "meoh"], ["h2o", "meoh", "meoh"]]
    dicts = [synthetic.make_model(f, n_train=6, sig=100.0, seed=k) for k, f in enumerate(fams)]
    models = [mb.gdmlModel(dict(d), criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": d["entity_ids"]}, c))
              for d, c in zip(dicts, [6.0, 6.5, 10.0, 10.5, 11.0])]
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=len(comp), box=14.0, n_frames=2, seed=0, comp=comp)
    for pbc in (True, (True, True, False)):
        cell = mb.Cell(cell_v, cutoff=6.5, pbc=pbc)
        out = {}
        for kind in ("full", "pairlist"):
            ctx.set_culling(kind)
            pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
            masks = [accept_mask(Z, R, eids, mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)), m, cell) for m in models]
            E1, F1 = pred.predict(Z, R, eids, cids)  # plain call
            E2, F2 = pred.predict(Z, R, eids, cids)  # prepared step
            out[kind] = (masks, E1, F1, E2, F2, list(pred.last_stats))
        ctx.set_culling("auto")
        a, b = out["full"], out["pairlist"]
        assert all(np.array_equal(x, y) for x, y in zip(a[0], b[0]))
        assert a[5] == b[5]
        for k in (1, 2, 3, 4):
            assert np.allclose(a[k], b[k], rtol=1e-11, atol=1e-11)
        print("pbc", pbc, "accepted", [s[2] for s in a[5]], "E", float(a[1][0]))
    print("ok")


if __name__ == "__main__":
    main()
