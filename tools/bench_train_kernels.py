"""Device timing of the SURVEY 8(f)4 kernels (kernel matrix, covariances) at training-set size 1000.
   python tools/bench_train_kernels.py > gpurun_out/train_kernels.json
CUDA-event times come from the library's own launch records (events on its launch stream)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mbgdml_b200 as mb  # noqa: E402
from mbgdml_b200 import _native, synthetic  # noqa: E402
from mbgdml_b200.analysis import gdml_mat52  # noqa: E402


def desc_of(R):
    n = R.shape[1]
    i, j = np.tril_indices(n, -1)
    d = R[:, i, :] - R[:, j, :]
    pd = np.sqrt((d * d).sum(-1))
    return 1.0 / pd, d / (pd ** 3)[..., None]


def main():
    ctx = _native.get_context(0)
    peak = ctx.measure_fp64_peak()[0]
    out = {"fp64_fma_peak_tflops": peak, "kernel_mat": [], "mat52": []}
    for comp, T in ((["h2o"], 1000), (["h2o", "h2o"], 1000), (["h2o", "h2o", "h2o"], 1000)):
        rng = np.random.default_rng(0)
        Rt = synthetic.fragment_geometries(rng, comp, T)
        X, G = desc_of(Rt)
        perms = synthetic.fragment_perms(comp)
        n, P = Rt.shape[1], len(perms)
        D = n * (n - 1) // 2
        tpl = (np.array([synthetic._desc_perm(p) for p in perms]) + np.arange(P)[:, None] * D).flatten("F")
        rows = T * 3 * n
        K = torch.empty((rows, rows), dtype=torch.float64, device="cuda")
        ms = []
        for _ in range(4):
            ctx.kernel_mat(X, G, tpl, 100.0, False, out=K.data_ptr())
            ctx.synchronize()
            ms.append(ctx.last_launches()[0]["ms"])
        t0 = time.perf_counter()
        sym = float((K - K.T).abs().max() / K.abs().max())
        # algorithmic work per (i, j) block: P (D + 2 (3N)(N-1) 2 + (3N)^2 2 + 6 D) + (3N)^2 (N-1) 2 flop
        N3 = 3 * n
        flop = P * (3 * D + 4 * N3 * (n - 1) + 2 * N3 * N3 + 6 * D) + 2 * N3 * N3 * (n - 1)
        best = min(ms[1:])
        out["kernel_mat"].append({"n_atoms": n, "n_train": T, "n_perms": P, "rows": rows, "K_bytes": rows * rows * 8,
                                  "ms": best, "blocks_per_s": T * T / best * 1e3, "tflops_algorithmic": flop * T * T / best / 1e9,
                                  "write_gbs": rows * rows * 8 / best / 1e6, "max_asym_rel": sym})
        del K
        torch.cuda.empty_cache()
    for comp, M in ((["h2o"], 20000), (["h2o", "h2o"], 4000), (["h2o", "h2o", "h2o"], 1000)):
        md = synthetic.make_model(comp, n_train=1000, sig=100.0, seed=1)
        model = mb.gdmlModel(dict(md))
        R = synthetic.fragment_geometries(np.random.default_rng(2), comp, M)
        ms = []
        for _ in range(3):
            got = gdml_mat52(model, md["z"], R)
            ms.append(ctx.last_launches()[0]["ms"])
        n = len(md["z"])
        D, P = n * (n - 1) // 2, md["perms"].shape[0]
        best = min(ms[1:])
        out["mat52"].append({"n_atoms": n, "n_structures": M, "rows": 1000 * P, "ms": best,
                             "covariances_per_s": M * 1000 * P / best * 1e3, "write_gbs": got.nbytes / best / 1e6,
                             "tflops_algorithmic": M * 1000 * P * (3 * D + 8) / best / 1e9})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
