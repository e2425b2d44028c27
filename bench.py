#!/usr/bin/env python
"""Benchmark of the many-body GDML prediction path (BASELINE.json metric: n-body fragment E+F
evaluations per second; FP64 roofline %).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one pass of the hot path (enumerate -> cull -> contract -> scatter [-> all-reduce]) over
one batch of synthetic frames.  Workloads (BASELINE.json configs):
  box140   (default) 140-H2O periodic box, minimum-image 1+2+3-body MBE with criteria and alchemical
           scaling, n_train = 1000 models, --frames frames per GPU per step (config 4, weak scaling)
  trimers64  64-H2O cluster, 3-body model only, 41,664 trimers per frame (config 3)
  monomers   1-body water model on 100,000 monomer structures (config 2)
  mixed140   100 H2O + 40 MeOH periodic box, 9 models with fragments of 3..18 atoms (config 5)

`value` is measured with the inputs resident in HBM (device pointers through the C ABI, CUDA events);
`e2e` goes through the public mbePredict.predict call with pinned HOST buffers (H2D + D2H inside the
timed region).  Multi-GPU: one process per GPU (torchrun), fragments partitioned block-cyclically,
one NCCL all-reduce of [E | F]; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "nbody_fragment_EF_evals_per_s"
UNIT = "fragments/s"


# ----------------------------------------------------------------------------------------------
# workloads (synthetic structures + random-init models of the named shapes)
# ----------------------------------------------------------------------------------------------
def make_workload(name, n_frames, n_train=1000):
    from mbgdml_b200 import synthetic as syn

    w = {"name": name, "n_train": n_train}
    if name == "box140":
        Z, R, eids, cids, cell = syn.make_periodic_box(n_mol=140, box=16.12, n_frames=n_frames, seed=0, frame_sigma=0.1)
        w.update(Z=Z, R=R, eids=eids, cids=cids, cell=cell)
        w["families"] = [["h2o"], ["h2o", "h2o"], ["h2o", "h2o", "h2o"]]
        w["sigs"] = [10.0, 100.0, 400.0]
        w["cutoffs"] = [None, 6.0, 10.0]
        w["alchemy"] = [(0, 3, 0.5), (0, 2, 0.5)]
        w["label"] = (f"140-H2O periodic box L=16.12 A, MIC cutoff L/2, 1+2+3-body MBE, criteria "
                      f"com_distance_sum None/6.0/10.0, alchemical factor 0.5, n_train={n_train}")
    elif name == "mixed140":
        comp = ["h2o"] * 100 + ["meoh"] * 40
        Z, R, eids, cids, cell = syn.make_periodic_box(n_mol=140, box=17.8, n_frames=n_frames, seed=0, comp=comp,
                                                       frame_sigma=0.1)
        w.update(Z=Z, R=R, eids=eids, cids=cids, cell=cell)
        w["families"] = [["h2o"], ["meoh"], ["h2o", "h2o"], ["h2o", "meoh"], ["meoh", "meoh"], ["h2o", "h2o", "h2o"],
                         ["h2o", "h2o", "meoh"], ["h2o", "meoh", "meoh"], ["meoh", "meoh", "meoh"]]
        w["sigs"] = [10.0, 20.0, 100.0, 120.0, 150.0, 400.0, 420.0, 450.0, 500.0]
        w["cutoffs"] = [None, None, 6.0, 6.5, 7.0, 10.0, 10.5, 11.0, 11.5]
        w["alchemy"] = [(100, 3, 0.5)]
        w["label"] = (f"100 H2O + 40 MeOH periodic box L=17.8 A, 9 models (fragments of 3..18 atoms), MIC cutoff L/2, "
                      f"criteria com_distance_sum, alchemical 3-body factor 0.5 on one methanol, n_train={n_train}")
    elif name == "trimers64":
        Z, R0, eids, cids = syn.make_cluster(64, seed=0)
        w.update(Z=Z, R=syn.jitter_frames(R0, n_frames, sigma=0.02, seed=0), eids=eids, cids=cids, cell=None)
        w["families"] = [["h2o", "h2o", "h2o"]]
        w["sigs"] = [100.0]
        w["cutoffs"] = [None]
        w["alchemy"] = []
        w["label"] = f"64-H2O cluster, 3-body model only (41,664 trimers per frame), n_train={n_train}"
    elif name == "monomers":
        rng = np.random.default_rng(0)
        n = 100000
        r = syn.WATER_R[None] + rng.normal(scale=0.05, size=(n, 3, 3))
        w.update(Z=np.tile(syn.WATER_Z, n), R=np.repeat(r.reshape(1, n * 3, 3), n_frames, axis=0),
                 eids=np.repeat(np.arange(n), 3), cids=np.array(["h2o"] * n), cell=None)
        w["families"] = [["h2o"]]
        w["sigs"] = [10.0]
        w["cutoffs"] = [None]
        w["alchemy"] = []
        w["label"] = f"1-body water model on 100,000 monomer structures per frame, n_train={n_train}"
    else:
        raise SystemExit(f"unknown workload {name}")
    w["model_dicts"] = [syn.make_model(f, n_train=n_train, sig=s, seed=k)
                        for k, (f, s) in enumerate(zip(w["families"], w["sigs"]))]
    return w


def flops_per_fragment(md):
    """ALGORITHMIC work W_frag = T*P*(9*D + 10) FP64 flop, transcendentals not counted."""
    n = len(md["z"])
    D = n * (n - 1) // 2
    return md["R_desc"].shape[1] * md["perms"].shape[0] * (9 * D + 10)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), smax.append(float(r[1])), power.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] if power else sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port (NumPy restatement of the reference) on the host cores
# ----------------------------------------------------------------------------------------------
_CPU_W = None  # workload shared with forked workers (set before the pool is created)


def _cpu_chunk(args):
    from threadpoolctl import threadpool_limits

    from oracle import mbgdml_oracle as orc

    model_idx, combs = args
    w = _CPU_W
    d = w["model_dicts"][model_idx]
    crit = orc.Criteria(orc.com_distance_sum, {"entity_ids": d["entity_ids"]}, w["cutoffs"][model_idx])
    model = _cpu_chunk.cache.get(model_idx)
    if model is None:
        model = _cpu_chunk.cache[model_idx] = orc.Model(dict(d), criteria=crit)
    cell = orc.Cell(w["cell"]) if w["cell"] is not None else None
    with threadpool_limits(limits=1):
        orc.predict_gdml(w["Z"], w["R"][0], w["eids"], combs, model, cell)
    return orc.predict_gdml.last_n_eval


_cpu_chunk.cache = {}


def cpu_reference_pass(w, n_candidates, pool, cores):
    """One bounded sample of the workload on the host: the first `n_candidates` candidate tuples of the
    highest-order model on frame 0 (culling included), fanned out over the cores in chunks as the
    reference's ray branch does (mbe.py:823-868).  Returns (accepted fragments, seconds)."""
    from oracle import mbgdml_oracle as orc

    k = len(w["model_dicts"]) - 1
    sets = orc.get_avail_entities(w["cids"], w["model_dicts"][k]["comp_ids"])
    combs = []
    for c in orc.gen_combs(sets):
        combs.append(c)
        if len(combs) >= n_candidates:
            break
    chunk = max(1, len(combs) // (cores * 8))
    tasks = [(k, combs[i:i + chunk]) for i in range(0, len(combs), chunk)]
    t0 = time.perf_counter()
    n_eval = sum(pool.map(_cpu_chunk, tasks))
    return n_eval, time.perf_counter() - t0, len(combs)


def cpu_sample_size(w):
    """Candidate tuples per CPU sample: ~10-20 s of work on 16 host cores (the GPU does all of them)."""
    return {"box140": 200000, "trimers64": 6000, "monomers": 100000, "mixed140": 60000}[w["name"]]


def ncu_traffic(workload, n_frag_atoms, frames, use_dmma):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum from the
    committed `ncu --set full` capture of the same command, profiles/ncu_traffic.json), else None."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    for row in table:
        if (row["workload"] == workload and row["n_frag_atoms"] == n_frag_atoms and row["frames"] == frames
                and row["contraction"] == ("dmma" if use_dmma else "fma")):
            return row["dram_bytes_per_launch"]
    return None


def run_reference_arm(args):
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    global _CPU_W
    w = make_workload(args.workload, 1, args.n_train)
    w_small = _CPU_W = w
    cores = os.cpu_count() or 1
    n_cand = cpu_sample_size(w)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_reference_pass(w_small, max(cores * 8, n_cand // 8), pool, cores)
        n_tot, t_tot = 0, 0.0
        for _ in range(args.steps):
            n, t, n_scanned = cpu_reference_pass(w_small, n_cand, pool, cores)
            n_tot += n
            t_tot += t
    value = n_tot / t_tot
    sample = (f"per step: first {n_scanned} candidate tuples of the {len(w['families'][-1])}-body model on frame 0 "
              f"({n_tot // args.steps} accepted), culling included")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "parallelism": f"{cores} host processes"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import mbgdml_b200 as mb
    from mbgdml_b200 import _engine, _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # weak scaling (default, the driver's contract): fixed frames per GPU; strong: --frames is the whole job and the
    # fragments of every frame are sharded over the ranks (BASELINE config 3: one 64-H2O structure on 1/2/4/8 GPUs)
    n_frames = args.frames * world if args.scaling == "weak" else args.frames
    w = make_workload(args.workload, n_frames, args.n_train)
    ctx = _native.get_context(local)
    # a dedicated (non-default) torch stream: the library launches on it, the CUDA events that time the
    # steps are recorded on it, and NCCL orders the all-reduce after it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    models = [mb.gdmlModel(dict(d), criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": d["entity_ids"]}, c))
              for d, c in zip(w["model_dicts"], w["cutoffs"])]
    scalers = [mb.mbeAlchemyScale(e, o, mb.linear_switching, {"factor": f}) for e, o, f in w["alchemy"]]
    cell = mb.Cell(w["cell"], pbc=True) if w["cell"] is not None else None
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=scalers, periodic_cell=cell)
    Z, eids, cids = w["Z"], w["eids"], w["cids"]
    n_atoms = len(Z)

    # pinned host input for the end-to-end leg, device-resident copy for the kernel leg
    R_pinned = torch.empty((n_frames, n_atoms, 3), dtype=torch.float64).pin_memory()
    R_pinned.copy_(torch.from_numpy(w["R"]))
    R_host = R_pinned.numpy()
    R_dev = R_pinned.to(dev)
    buf = torch.zeros(n_frames * (1 + 3 * n_atoms), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def sources():
        return [mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)) for m in models]

    def step_device():
        E_t, F_t = buf[:n_frames], buf[n_frames:]
        _, _, _, st = _engine.predict_total(ctx, Z, R_dev, eids, models, sources(), [pred._scalers_for(m) for m in models],
                                            cell, shard=(rank, world), device_out=(E_t, F_t, None))
        if world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        return st

    def step_e2e():
        E, F = pred.predict(Z, R_host, eids, cids)
        return float(E.sum() + F[0, 0, 0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed (untimed) in between."""
        total_ms, last = 0.0, None
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            last = fn()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            wall = (time.perf_counter() - t0) * 1e3
            total_ms += max(e0.elapsed_time(e1), wall if fn is step_e2e else 0.0)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), last

    ctx.set_contraction(args.contraction)
    use_dmma = args.contraction != "fma"
    peak_fma, clk_est = ctx.measure_fp64_peak() if rank == 0 else (0.0, 0.0)
    peak_dmma = ctx.measure_fp64_tensor_peak() if rank == 0 else 0.0
    # the roofline of a kernel is the peak of the pipe it runs on
    peak_tflops = peak_dmma if use_dmma else peak_fma

    for _ in range(max(args.warmup, 3)):
        st = step_device()
    barrier()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, st = timed(step_device, args.steps)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    recs = ctx.last_launches()  # contraction launches of the last step on this rank

    # accepted fragments over all ranks
    acc = torch.tensor([s[2] for s in st], dtype=torch.float64, device=dev)
    enum = torch.tensor([s[1] for s in st], dtype=torch.float64, device=dev)
    nl = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(acc), dist.all_reduce(enum), dist.all_reduce(nl)
    acc, enum = acc.cpu().numpy(), enum.cpu().numpy()
    frag_per_step = float(acc.sum())
    value = frag_per_step * args.steps / (ms_dev * 1e-3)

    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_value = frag_per_step * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (the contraction launch with the most work in the last step)
    def rec_flops(r):  # W_frag = T * P * (9 D + 10) of the model this launch contracted against
        n = r["n_frag_atoms"]
        return args.n_train * r["n_perms"] * (9 * (n * (n - 1) // 2) + 10)

    wfrag = {r["n_frag_atoms"]: rec_flops(r) for r in recs}
    best = max(recs, key=lambda r: r["n_fragments"] * rec_flops(r)) if recs else None
    roofline = None
    if best and best["ms"] > 0:
        dom = [r for r in recs if (r["n_frag_atoms"], r["n_perms"]) == (best["n_frag_atoms"], best["n_perms"])]
        flop = sum(r["n_fragments"] * rec_flops(r) for r in dom)
        ms = sum(r["ms"] for r in dom)
        achieved = flop / (ms * 1e-3) / 1e12
        kname = f"k_matern_mma<NA={best['n_frag_atoms']}>" if use_dmma else f"k_matern<NA={best['n_frag_atoms']}>"
        roofline = {
            "bound": "tensor" if use_dmma else "fp64_fma", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / peak_tflops if peak_tflops else None,
            "traffic": ncu_traffic(args.workload, best["n_frag_atoms"], args.frames, use_dmma),
            "kernel": kname, "launches_in_step": len(dom),
            "avg_launch_ms": ms / len(dom), "alg_flop_per_fragment": rec_flops(best),
            "fragments_per_launch": sum(r["n_fragments"] for r in dom) / len(dom),
            "pipe": ("FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA.8x8x4)" if use_dmma else "FP64 FMA pipe (DFMA)"),
            "peak_source": "measured live on this GPU, CUDA events: back-to-back DMMA m8n8k4 on 8 independent accumulators "
                           "(mbg_measure_fp64_tensor_peak) for the tensor pipe, 8 independent DFMA chains "
                           "(mbg_measure_fp64_peak) for the FMA pipe; MEASURED_PEAKS.json has no FP64 entry; "
                           "nominal 148 SM x 64 FMA x 2 x 1.965 GHz = 37.2 TFLOP/s for either",
            "peak_fp64_fma_tflops": peak_fma, "peak_fp64_tensor_tflops": peak_dmma,
            "frac_of_fma_pipe_peak": achieved / peak_fma if peak_fma else None,
            "contract_share_of_step": sum(r["ms"] for r in recs) / (ms_dev / args.steps),
        }
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            roofline["hbm_peak_gbs_measured"] = mp.get("hbm_gbs")
        except Exception:
            pass

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp_

        global _CPU_W
        _CPU_W = w
        cores = os.cpu_count() or 1
        with mp_.get_context("fork").Pool(cores) as pool:
            n, t, n_scanned = cpu_reference_pass(w, cpu_sample_size(w), pool, cores)
        cpu_baseline = {
            "value": n / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n_scanned} candidate tuples of the {len(w['families'][-1])}-body model on frame 0 "
                      f"({n} accepted, {t:.1f} s), culling included; NumPy oracle port, one process per core",
        }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": w["label"], "frames_per_step": n_frames,
            "frames_per_gpu": args.frames if args.scaling == "weak" else n_frames / world,
            "fragments_enumerated_per_step": [int(x) for x in enum], "fragments_accepted_per_step": [int(x) for x in acc],
            "structures_per_s": n_frames * args.steps / (ms_dev * 1e-3),
            "parallelism": f"fragment-sharded x{world}, one all-reduce" if world > 1 else "single GPU",
            "l2": "flushed between steps (256 MiB fill, untimed); model tables are L2-resident by design within a step",
        },
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(R_host.nbytes), "d2h_bytes_per_step": int(8 * n_frames * (1 + 3 * n_atoms))},
        "gpu_launches": int(nl.item()), "clocks": clocks,
        "fp64_peak_tflops_measured": {"fma_pipe": peak_fma, "tensor_pipe": peak_dmma},
        "contraction_launches": [
            {"n_frag_atoms": r["n_frag_atoms"], "n_perms": r["n_perms"], "n_fragments": r["n_fragments"], "ms": round(r["ms"], 4),
             "tflops": round(r["n_fragments"] * rec_flops(r) / (r["ms"] * 1e-3) / 1e12, 2) if r["ms"] > 0 else None}
            for r in recs],
        "sm_clock_mhz_implied_by_fma_peak": clk_est, "contraction": "dmma" if use_dmma else "fma",
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="box140", choices=["box140", "trimers64", "monomers", "mixed140"])
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --frames per GPU (default); strong: --frames in total, fragments sharded over the GPUs")
    ap.add_argument("--n-train", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contraction", default="auto", choices=["auto", "fma", "dmma"],
                    help="contraction kernel: FP64 tensor pipe (dmma, default) or FMA pipe (fma)")
    args = ap.parse_args()
    if args.frames is None:
        args.frames = {"box140": 8, "trimers64": 4, "monomers": 1, "mixed140": 4}[args.workload]
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
