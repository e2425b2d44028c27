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
# CPU arm: the reference's own predict_gdml (unmodified package from oracle/_ref, see oracle/fetch_ref.py) on the
# host cores; the NumPy oracle port only when that install is absent.  Also the checker for `parity`.
# ----------------------------------------------------------------------------------------------
_CPU_W = None  # workload shared with forked workers (set before the pool is created)


def cpu_kind():
    from oracle import reference

    return "reference" if reference.available() else "port"


def _cpu_objects(model_idx):
    """(predict function, model, cell, scalers) of the CPU implementation in this worker (cached)."""
    hit = _cpu_objects.cache.get(model_idx)
    if hit is not None:
        return hit
    w = _CPU_W
    d = w["model_dicts"][model_idx]
    order = len(d["comp_ids"])
    if cpu_kind() == "reference":
        from oracle import reference

        ref = reference.load()
        crit = ref.Criteria(ref.com_distance_sum, {"entity_ids": d["entity_ids"]}, w["cutoffs"][model_idx])
        model = ref.gdmlModel(dict(d), criteria=crit)
        cell = ref.Cell(w["cell"], pbc=True) if w["cell"] is not None else None
        scalers = [ref.mbeAlchemyScale(e, o, ref.linear_switching, {"factor": f}) for e, o, f in w["alchemy"] if o == order]
        fn = lambda Z, R, eids, combs: ref.predict_gdml(Z, R, eids, combs, model, cell, alchemy_scalers=scalers)[:2]
    else:
        from oracle import mbgdml_oracle as orc

        crit = orc.Criteria(orc.com_distance_sum, {"entity_ids": d["entity_ids"]}, w["cutoffs"][model_idx])
        model = orc.Model(dict(d), criteria=crit)
        cell = orc.Cell(w["cell"]) if w["cell"] is not None else None
        scalers = [orc.AlchemyScale(e, o, f) for e, o, f in w["alchemy"] if o == order]
        fn = lambda Z, R, eids, combs: orc.predict_gdml(Z, R, eids, combs, model, cell, alchemy_scalers=scalers)[:2]
    _cpu_objects.cache[model_idx] = fn
    return fn


_cpu_objects.cache = {}


def _cpu_chunk(args):
    from threadpoolctl import threadpool_limits

    model_idx, combs = args
    w = _CPU_W
    fn = _cpu_objects(model_idx)
    with threadpool_limits(limits=1):
        E, F = fn(w["Z"], w["R"][0], w["eids"], [tuple(int(e) for e in c) for c in combs])
    return float(E), F


def cpu_sample(w):
    """The bounded sample of the workload the CPU arm evaluates: the first `cpu_sample_size` candidate tuples of the
    highest-order model on frame 0, in gen_combs order (culling is part of the pass).  Returns (model index, combs)."""
    from oracle import mbgdml_oracle as orc

    k = len(w["model_dicts"]) - 1
    sets = orc.get_avail_entities(w["cids"], w["model_dicts"][k]["comp_ids"])
    n_candidates = cpu_sample_size(w)
    combs = []
    for c in orc.gen_combs(sets):
        combs.append(c)
        if len(combs) >= n_candidates:
            break
    return k, np.array(combs, dtype=np.int64).reshape(len(combs), -1)


def cpu_sample_accepted(w, k, combs):
    """Accept mask of the sample by the vectorised oracle culling (untimed; gives the fragment count of the metric)."""
    from oracle import cull_vec
    from oracle import mbgdml_oracle as orc

    d = w["model_dicts"][k]
    om = orc.Model(dict(d), criteria=orc.Criteria(orc.com_distance_sum, {"entity_ids": d["entity_ids"]}, w["cutoffs"][k]))
    return cull_vec.accept_mask(w["Z"], w["R"][0], w["eids"], combs, om, orc.Cell(w["cell"]) if w["cell"] is not None else None)


def cpu_reference_pass(k, combs, pool, cores):
    """One pass over the sample, fanned out over the cores in chunks as the reference's ray branch does
    (mbe.py:823-868).  Returns (E, F, seconds)."""
    chunk = max(1, len(combs) // (cores * 8))
    tasks = [(k, combs[i:i + chunk]) for i in range(0, len(combs), chunk)]
    t0 = time.perf_counter()
    parts = pool.map(_cpu_chunk, tasks)
    dt = time.perf_counter() - t0
    return sum(p[0] for p in parts), sum(p[1] for p in parts), dt


def cpu_sample_size(w):
    """Candidate tuples per CPU sample: ~10-20 s of work on 16 host cores (the GPU does all of them)."""
    return {"box140": 200000, "trimers64": 6000, "monomers": 100000, "mixed140": 60000}[w["name"]]


def ncu_traffic(workload, n_frag_atoms, frames, contraction):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the
    committed `ncu --set full` capture of the same command (profiles/ncu_traffic.json): a STATIC figure that names
    the capture it came from -- (bytes, source) or (None, None)."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None, None
    for row in table:
        if (row["workload"] == workload and row["n_frag_atoms"] == n_frag_atoms and row["frames"] == frames
                and row["contraction"] == contraction):
            return row["dram_bytes_per_launch"], "static: " + row.get("source", "ncu --set full capture, profiles/")
    return None, None


def sample_text(w, n_scanned, n_acc):
    return (f"first {n_scanned} candidate tuples of the {len(w['families'][-1])}-body model on frame 0 in gen_combs order "
            f"({n_acc} accepted), culling included")


def run_reference_arm(args):
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    global _CPU_W
    w = _CPU_W = make_workload(args.workload, 1, args.n_train)
    cores = os.cpu_count() or 1
    k, combs = cpu_sample(w)
    n_acc = int(cpu_sample_accepted(w, k, combs).sum())
    kind = cpu_kind()
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_reference_pass(k, combs[:max(cores * 8, len(combs) // 8)], pool, cores)
        t_tot = 0.0
        for _ in range(args.steps):
            t_tot += cpu_reference_pass(k, combs, pool, cores)[2]
    value = n_acc * args.steps / t_tot
    impl_text = ("unmodified reference package (oracle/_ref): mbgdml.predictors.predict_gdml per chunk, one process per "
                 "core, the multiprocessing mirror of the ray fan-out of mbe.py:823-868" if kind == "reference"
                 else "NumPy oracle port (oracle/_ref absent), one process per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "parallelism": f"{cores} host processes", "implementation": impl_text},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "per step: " + sample_text(w, len(combs), n_acc)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def candidate_generation_record(ctx, n_mol=1000):
    """How the enumerate -> cull -> compact stage scales beyond the benchmark box (outside the timed region): one frame
    of n_mol waters at the box's density, cut-off 8 A, 2-body < 6.0 / 3-body < 10.0, through the full enumeration
    (every C(n, 3) tuple, as the reference does) and through the pair list (csrc/pairlist.cuh).  `ms_*` = device time of
    everything but the contraction; the accept masks of all candidates are compared byte for byte."""
    import math

    import mbgdml_b200 as mb
    from mbgdml_b200 import _engine, synthetic
    from mbgdml_b200.predictors import accept_mask

    dicts = [synthetic.make_model(["h2o"] * k, n_train=8, sig=[100.0, 400.0][k - 2], seed=k) for k in (2, 3)]
    models = [mb.gdmlModel(dict(d), criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": d["entity_ids"]}, c))
              for d, c in zip(dicts, [6.0, 10.0])]
    box = 16.12 * (n_mol / 140) ** (1 / 3)
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=n_mol, box=box, n_frames=1, seed=0, frame_sigma=0.1)
    cell = mb.Cell(cell_v, cutoff=8.0, pbc=True)
    pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
    rec = {"workload": f"{n_mol} H2O, one frame, L = {box:.2f} A, cut-off 8 A, com_distance_sum 6.0 / 10.0",
           "candidates": [math.comb(n_mol, 2), math.comb(n_mol, 3)]}
    masks = {}
    try:
        for kind in ("full", "pairlist"):
            ctx.set_culling(kind)
            best = None
            for _ in range(3):
                srcs = [mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)) for m in models]
                _, _, _, stats = _engine.predict_total(ctx, Z, R, eids, models, srcs, [[], []], cell)
                t = ctx.last_timing()["ms_other"]
                best = t if best is None else min(best, t)
            rec[f"ms_{kind}"] = round(best, 4)
            rec["accepted"] = [int(s_[2]) for s_ in stats]
            masks[kind] = accept_mask(Z, R, eids, mb.gen_combs(pred.get_avail_entities(cids, models[1].comp_ids)), models[1], cell)
    finally:
        ctx.set_culling("auto")
    rec["accept_masks_equal"] = bool(np.array_equal(masks["full"], masks["pairlist"]))
    rec["accepted_in_mask"] = int((masks["pairlist"] == 1).sum())
    return rec


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
    if world > 1:
        pred.shard_group = dist.group.WORLD  # opt-in fragment sharding + one all-reduce inside predict()
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
    exp_sqrt_rate = ctx.measure_exp_sqrt_rate() if rank == 0 else 0.0
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

    # the same step replayed from a captured CUDA graph (mbg_step_create / mbg_step_run_device): device-resident in/out
    graph = None
    if args.contraction in ("auto", "dmma"):
        gstep = _engine.make_step(ctx, Z, eids, models, sources(), [pred._scalers_for(m) for m in models], cell, n_frames,
                                  shard=(rank, world))

        def step_graph():
            gstep.run_device(R_dev.data_ptr(), buf[:n_frames].data_ptr(), buf[n_frames:].data_ptr())
            if world > 1:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)

        for _ in range(3):
            step_graph()
        ms_graph, _ = timed(step_graph, args.steps)
        graph = {"ms_per_step": ms_graph / args.steps, "value": frag_per_step * args.steps / (ms_graph * 1e-3), "unit": UNIT,
                 "what": "the step of `value` replayed as one CUDA graph launch (tables resident, no host round trip)"}
        gstep.close()

    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_value = frag_per_step * args.steps / (ms_e2e * 1e-3)

    # N > 1: the all-reduced result of frame 0 equals the un-sharded single-rank evaluation of the same frame
    mg_check = None
    if world > 1:
        step_device()
        torch.cuda.synchronize(dev)
        E_all, F_all = buf[:1].cpu().numpy().copy(), buf[n_frames:n_frames + 3 * n_atoms].cpu().numpy().copy()
        one = torch.zeros(1 + 3 * n_atoms, dtype=torch.float64, device=dev)
        _engine.predict_total(ctx, Z, R_dev[:1].contiguous(), eids, models, sources(), [pred._scalers_for(m) for m in models],
                              cell, shard=(0, 1), device_out=(one[:1], one[1:], None))
        torch.cuda.synchronize(dev)
        one = one.cpu().numpy()
        err = torch.tensor([abs(E_all[0] - one[0]) / max(abs(one[0]), 1e-300),
                            np.abs(F_all - one[1:]).max() / max(np.abs(one[1:]).max(), 1e-300)], dtype=torch.float64, device=dev)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        mg_check = {"what": "all-reduced [E|F] of frame 0 vs the same frame evaluated un-sharded on every rank (max over ranks)",
                    "max_rel_err_E": float(err[0].item()), "max_rel_err_F": float(err[1].item()),
                    "ok": bool(err.max().item() <= 1e-11)}

    # strong-scaling sub-record (BASELINE config 3): ONE 64-H2O structure, its 41,664 trimers sharded over the ranks
    strong = None
    if not args.no_strong:
        ws = make_workload("trimers64", 1, args.n_train)
        ms_ = [mb.gdmlModel(dict(d)) for d in ws["model_dicts"]]
        ps = mb.mbePredict(ms_, mb.predict_gdml)
        Rs = torch.from_numpy(ws["R"]).to(dev)
        na_s = len(ws["Z"])
        bs = torch.zeros(1 + 3 * na_s, dtype=torch.float64, device=dev)

        def step_strong():
            _engine.predict_total(ctx, ws["Z"], Rs, ws["eids"], ms_, [mb.gen_combs(ps.get_avail_entities(ws["cids"], m.comp_ids)) for m in ms_],
                                  [[]], None, shard=(rank, world), device_out=(bs[:1], bs[1:], None))
            if world > 1:
                dist.all_reduce(bs, op=dist.ReduceOp.SUM)

        for _ in range(3):
            step_strong()
        n_s = 10
        ms_plain, _ = timed(step_strong, n_s)
        kern_ms = sum(r["ms"] for r in ctx.last_launches())
        # the same evaluation as a prepared step (what mbePredict.predict replays from its second call on): this
        # rank's shard of the structure in one CUDA-graph launch, then the all-reduce
        sstep = _engine.make_step(ctx, ws["Z"], ws["eids"], ms_, [mb.gen_combs(ps.get_avail_entities(ws["cids"], m.comp_ids)) for m in ms_],
                                  [[]], None, 1, shard=(rank, world))

        def step_strong_graph():
            sstep.run_device(Rs.data_ptr(), bs[:1].data_ptr(), bs[1:].data_ptr())
            if world > 1:
                dist.all_reduce(bs, op=dist.ReduceOp.SUM)

        for _ in range(3):
            step_strong_graph()
        ms_s, _ = timed(step_strong_graph, n_s)
        sstep.close()
        strong = {"workload": ws["label"], "frames": 1, "fragments": 41664, "n_gpus": world, "steps": n_s,
                  "ms_per_structure": ms_s / n_s, "structures_per_s": n_s / (ms_s * 1e-3),
                  "ms_per_structure_plain_call": ms_plain / n_s,
                  "contraction_ms_rank0": kern_ms, "scaling": "strong",
                  "efficiency_inputs": "efficiency(N) = ms_per_structure(1) / (N * ms_per_structure(N)) across the --gpus lines"}

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
        kname = (f"k_matern_pw<NA={best['n_frag_atoms']}>" if args.contraction in ("auto", "dmma") else
                 f"k_matern_mma<NA={best['n_frag_atoms']}>" if use_dmma else f"k_matern<NA={best['n_frag_atoms']}>")
        traffic, traffic_source = ncu_traffic(args.workload, best["n_frag_atoms"], args.frames,
                                              {"auto": "dmma"}.get(args.contraction, args.contraction))
        roofline = {
            "bound": "tensor" if use_dmma else "fp64_fma", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / peak_tflops if peak_tflops else None,
            "traffic": traffic, "traffic_source": traffic_source,
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
            "contract_share_of_graph_step": (sum(r["ms"] for r in recs) / graph["ms_per_step"]) if graph else None,
            # transcendental side of the same kernel (one exp and one sqrt per (query, training row) pair)
            "exp_sqrt_per_s": sum(r["n_fragments"] for r in dom) * args.n_train * best["n_perms"] / (ms * 1e-3),
            "exp_sqrt_rate_measured": exp_sqrt_rate,
            "exp_sqrt_frac": (sum(r["n_fragments"] for r in dom) * args.n_train * best["n_perms"] / (ms * 1e-3) / exp_sqrt_rate
                              if exp_sqrt_rate else None),
            "exp_sqrt_rate_source": "mbg_measure_exp_sqrt_rate: exp(-c sqrt(x)) with the CUDA math library, 4 chains per thread, "
                                    "CUDA events, measured live",
        }
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            roofline["hbm_peak_gbs_measured"] = mp.get("hbm_gbs")
        except Exception:
            pass

    cpu_baseline, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp_

        global _CPU_W
        _CPU_W = w
        cores = os.cpu_count() or 1
        k, combs = cpu_sample(w)
        mask_cpu = cpu_sample_accepted(w, k, combs)
        with mp_.get_context("fork").Pool(cores) as pool:
            E_cpu, F_cpu, t = cpu_reference_pass(k, combs, pool, cores)
        n = int(mask_cpu.sum())
        cpu_baseline = {
            "value": n / t, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
            "sample": sample_text(w, len(combs), n) + f"; {t:.1f} s; " +
                      ("unmodified reference predict_gdml (oracle/_ref), one process per core" if cpu_kind() == "reference"
                       else "NumPy oracle port, one process per core"),
        }
        # parity on the benchmark workload itself: the same sampled tuples through the GPU path (explicit combs)
        from mbgdml_b200.predictors import accept_mask

        order = len(w["families"][k])
        sc = [s_ for s_ in scalers if s_.order == order]
        E_gpu, F_gpu = mb.predict_gdml(Z, w["R"][0], eids, combs, models[k], cell, alchemy_scalers=sc)
        mask_gpu = accept_mask(Z, w["R"][0], eids, combs, models[k], cell)[0] == 1
        parity = {
            "checked_against": cpu_baseline["kind"], "n_candidates": int(len(combs)), "n_fragments": n,
            "accept_mask_equal": bool(np.array_equal(mask_gpu, mask_cpu)),
            "max_rel_err_E": float(abs(E_gpu - E_cpu) / max(abs(E_cpu), 1e-300)),
            "max_rel_err_F": float(np.abs(F_gpu - F_cpu).max() / max(np.abs(F_cpu).max(), 1e-300)),
            "tolerance": 1e-9,
        }
        parity["ok"] = bool(parity["accept_mask_equal"] and parity["max_rel_err_E"] <= 1e-9 and parity["max_rel_err_F"] <= 1e-9)

    cand_gen = None
    if world == 1 and not args.no_cull_record:
        cand_gen = candidate_generation_record(ctx)

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
        "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "multi_gpu_check": mg_check,
        "strong": strong, "graph": graph, "candidate_generation": cand_gen,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(R_host.nbytes), "d2h_bytes_per_step": int(8 * n_frames * (1 + 3 * n_atoms))},
        "gpu_launches": int(nl.item()), "clocks": clocks,
        "fp64_peak_tflops_measured": {"fma_pipe": peak_fma, "tensor_pipe": peak_dmma},
        "contraction_launches": [
            {"n_frag_atoms": r["n_frag_atoms"], "n_perms": r["n_perms"], "n_fragments": r["n_fragments"], "ms": round(r["ms"], 4),
             "tflops": round(r["n_fragments"] * rec_flops(r) / (r["ms"] * 1e-3) / 1e12, 2) if r["ms"] > 0 else None}
            for r in recs],
        "sm_clock_mhz_implied_by_fma_peak": clk_est, "contraction": {"auto": "dmma"}.get(args.contraction, args.contraction),
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
    ap.add_argument("--no-cull-record", action="store_true", help="skip the candidate-generation sub-record (1000-H2O frame)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record (one 64-H2O structure)")
    ap.add_argument("--contraction", default="auto", choices=["auto", "fma", "dmma", "dmma_wave"],
                    help="contraction kernel: FP64 tensor pipe, persistent per-warp workers (dmma, default), the same "
                         "arithmetic wave-launched (dmma_wave), or the FMA pipe (fma)")
    args = ap.parse_args()
    if args.frames is None:
        args.frames = {"box140": 8, "trimers64": 4, "monomers": 1, "mixed140": 4}[args.workload]
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
