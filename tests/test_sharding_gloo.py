"""CPU, world_size 2 over gloo: the fragment partition (block-cyclic granules, the rule the CUDA
kernels implement) covers every fragment exactly once, and summing the per-rank partial E/F with one
all-reduce reproduces the single-process result.  Per-rank arithmetic is the oracle here (no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load, oracle_models, relerr, unpack_models
from mbgdml_b200 import sharding


def test_partition_is_exact_cover():
    for n in (0, 1, 255, 256, 257, 5000, 70001):
        for world in (1, 2, 3, 8):
            masks = [sharding.shard_mask(n, r, world) for r in range(world)]
            assert np.array_equal(np.sum(masks, axis=0), np.ones(n, dtype=int))
            if n >= 256 * world:
                sizes = [m.sum() for m in masks]
                assert max(sizes) - min(sizes) <= sharding.GRANULE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    from oracle import mbgdml_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the product's sharding switch: local by default although a process group exists (the reference's predict is a
    # local call); an explicit group must be NCCL (the all-reduce runs on a CUDA buffer)
    from mbgdml_b200 import mbe as _mbe

    assert _mbe._shard_group(None) is None
    try:
        _mbe._shard_group(dist.group.WORLD)
        raise AssertionError("a gloo group must be refused")
    except RuntimeError:
        pass
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    models = oracle_models(mds, [None, 4.2, 7.5])
    n_frames, n_atoms = R.shape[0], R.shape[1]
    buf = torch.zeros(n_frames + n_frames * n_atoms * 3, dtype=torch.float64)
    E, F = buf[:n_frames].numpy(), buf[n_frames:].numpy().reshape(n_frames, n_atoms, 3)
    for m in models:
        sets = orc.get_avail_entities(cids, m.comp_ids)
        sizes = [len(s) for s in sets]
        per_frame = int(np.prod(sizes))
        # global candidate index = frame * per_frame + product index (frame-major), as on device;
        # a tiny granule (4) is used so that both ranks get work on this small system
        for frame in range(n_frames):
            for cand, comb in enumerate(np.ndindex(*sizes)):
                gidx = frame * per_frame + cand
                if (gidx // 4) % world != rank:
                    continue
                tup = tuple(int(sets[k][i]) for k, i in enumerate(comb))
                if any(a >= b for a, b in zip(tup, tup[1:])):
                    continue
                e, f = orc.predict_gdml(Z, R[frame], eids, [tup], m)
                E[frame] += e
                F[frame] += f
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        ret["E"] = E.copy()
        ret["F"] = F.copy()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        E, F = ret["E"], ret["F"]
    g = load("cluster_6h2o.npz")
    assert relerr(E, g["crit_E"]) <= 1e-11 and relerr(F, g["crit_F"]) <= 1e-11
