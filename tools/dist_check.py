"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank evaluates its block-cyclic share of the fragments; after the single NCCL all-reduce the result
must equal the reference outputs committed in tests/golden (periodic box: E, F and stress)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mbgdml_b200 as mb  # noqa: E402
from helpers import b200_models, cutoffs_from, load, relerr, unpack_models  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    models = b200_models(mds, cutoffs_from(g["crit_cutoffs"]))
    al = [mb.mbeAlchemyScale(int(e), int(o), mb.linear_switching, {"factor": float(f)}) for e, o, f in g["alch"]]
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=al, periodic_cell=mb.Cell(g["cub_cell"], pbc=True),
                         compute_stress=True)
    pred.shard_group = dist.group.WORLD  # sharding is opt-in
    E, F, S = pred.predict(g["cub_Z"], g["cub_R"], g["cub_eids"], g["cub_cids"])
    errs = (relerr(E, g["cub_E"]), relerr(F, g["cub_F"]), relerr(S, g["cub_stress"]))
    acc = torch.tensor([s[2] for s in pred.last_stats], dtype=torch.float64, device="cuda")
    mine = acc.clone()
    dist.all_reduce(acc)
    print(f"rank {rank}/{world}: accepted here {mine.tolist()}, all ranks {acc.tolist()}, rel.err E/F/stress {errs}",
          flush=True)
    assert max(errs) <= 1e-9, errs
    dist.barrier()
    if rank == 0:
        print("DIST_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
