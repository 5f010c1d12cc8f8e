#!/usr/bin/env python
"""bench.py -- gene-pair x delay CMI evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload), at EVERY N: BASELINE.json configs[2], the one the metric is quoted on at 1/2/4/8 GPUs --
"dynamics-coupling I(x; velocity_y | y), 500 genes x 10,000 cells" = 249,500 ordered (TF, target) pairs, one cmi evaluation
each on the cells kept by the reference's zero-cell mask (~4,500 of 10,000).  STRONG scaling: the pair list is fixed and
sharded across the ranks; one all-gather of the result vector is the only collective.

A "step" is one pass of the hot path over the whole pair list.
  value : evaluations/s with the normalised layers already resident in HBM (pair table -> results on every rank, all-gather
          and read-back included), CUDA events, max over ranks.
  e2e   : the same through the public API, causal_net_dynamics_coupling(adata): the raw layers go host -> device, are
          normalised there, every pair is evaluated, and the result DataFrame lands in adata.uns (device -> host read).
  roofline     : the dominant kernel (ksg_sweepw_kernel) against the ALU roofline north_star names: algorithmic FP32-pipe
                 lane-ops of the point-pair visits it makes (6 per kNN visit + 10 per count visit, D = 3) /
                 kernel time / (SMs x 128 lanes x clock).
  cpu_baseline : the UNMODIFIED reference (baseline/_ref, kind "reference"; oracle/ref_port.py, kind "port", if absent) on a
                 bounded sample of the same pairs, all host cores, with the parity of the GPU results on that sample.
  extra_workloads (1 GPU): BASELINE configs[1] (C2, RDI 100 genes x 2,000 cells x 3 delays) and a target slab of configs[4]
                 (C5, uniformised RDI 1,000 genes x 20,000 cells), each with its own value and roofline.
  invariance_ok (N > 1): sharded == unsharded, bit for bit, on a sub-panel of the workload.
--impl reference times the CPU arm alone on the same workload (per step: a bounded sample of its pairs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "baseline"))

OUT = sys.stdout
C3_GENES, C3_CELLS = 500, 10000
C2_DELAYS = [1, 5, 10]
WORKLOAD = "dynamics-coupling I(x; velocity_y | y), %d genes x %d cells, %d ordered pairs, k=5, drop_zero_cells (BASELINE configs[2])"
METRIC = "gene-pair x delay CMI evaluations/sec"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic_r2.json), or None."""
    for name in ("traffic_r2.json", "r1/traffic_r1.json"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel)
            if v is not None:
                return v
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_uuid=None):
        self.rows, self.proc, self.uuid = [], None, gpu_uuid

    def start(self):
        cmd = ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"]
        if self.uuid:
            cmd += ["-i", self.uuid]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        busy = [s for s in sm if s > 500] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------------
# workloads (synthetic, seeded; generators shared with the parity tests: tests/golden/recipes.py)
# ----------------------------------------------------------------------------------------------------------------
def reference_goldens():
    """outputs of the UNMODIFIED reference on the named configs (tests/golden/golden_configs.npz, made by
    tests/golden/make_golden_configs.py in the build container), or {} if the file is absent"""
    try:
        return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_configs.npz")))
    except Exception:
        return {}


def c3_adata():
    import recipes
    spl, vel, genes = recipes.coupling_panel(C3_GENES, C3_CELLS, 3)
    return recipes.FakeAnnData({"spliced": spl, "velocity": vel}, genes), spl, vel, genes


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (baseline/_ref) on all host cores; the reference-shaped port if it is absent
# ----------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(spl, vel, genes):
    _CPU["spl"], _CPU["vel"], _CPU["genes"] = spl, vel, genes


def _cpu_c3_worker(task):
    """one TF against a few targets through the reference's own driver (Scribe.py:10-139), or the port of its inner loop"""
    tf, targets = task
    import recipes
    warnings.simplefilter("ignore")
    spl, vel, genes = _CPU["spl"], _CPU["vel"], _CPU["genes"]
    cols = [tf] + list(targets)
    names = [genes[c] for c in cols]
    import load_ref
    ref = load_ref.load()
    if ref is not None:
        devnull = open(os.devnull, "w")
        old = sys.stderr
        sys.stderr = devnull                                      # the reference prints a tqdm bar per call
        try:
            ad = recipes.FakeAnnData({"spliced": spl[:, cols].copy(), "velocity": vel[:, cols].copy()}, names)
            ref[1].causal_net_dynamics_coupling(ad, TFs=[names[0]], Targets=names[1:])
        finally:
            sys.stderr = old
        M = ad.uns["causal_net"]["RDI"]
        return [float(M.loc[names[0], t]) for t in names[1:]]
    from oracle import ref_port
    s = (spl[:, cols] - spl[:, cols].mean(0)) / (spl[:, cols].max(0) - spl[:, cols].min(0))
    v = (vel[:, cols] - vel[:, cols].mean(0)) / (vel[:, cols].max(0) - vel[:, cols].min(0))
    out = []
    for j in range(1, len(cols)):
        x, z = s[:, 0], s[:, j]; y = s[:, j] + v[:, j]
        keep = ((x + y) + z) > 0
        out.append(ref_port.cmi(x[keep], y[keep], z[keep]))
    return out


def cpu_kind():
    import load_ref
    return "reference" if load_ref.available() else "port"


def c3_tasks(n_tasks, per_task, seed):
    rng = np.random.default_rng(seed)
    tasks = []
    for _ in range(n_tasks):
        g = rng.choice(C3_GENES, per_task + 1, replace=False)
        tasks.append((int(g[0]), [int(v) for v in g[1:]]))
    return tasks


_POOL = {}


def _close_pools():
    for p in _POOL.values():
        p.terminate()
    _POOL.clear()


def cpu_pool(cores, spl, vel, genes):
    """One fork pool per process, created outside any timed region (the layers are shared copy-on-write)."""
    import atexit
    from multiprocessing import get_context
    _cpu_init(spl, vel, genes)
    if cores > 1 and cores not in _POOL:
        if not _POOL:
            atexit.register(_close_pools)
        _POOL[cores] = get_context("fork").Pool(cores)
        _POOL[cores].map(abs, range(cores))          # spin the workers up
    return _POOL.get(cores)


def run_cpu_c3(tasks, pool):
    t0 = time.perf_counter()
    res = pool.map(_cpu_c3_worker, tasks, chunksize=1) if pool is not None else [_cpu_c3_worker(t) for t in tasks]
    return res, time.perf_counter() - t0


def _cpu_rdi_worker(task):
    """one (source row, target row, delay, uniformised) evaluation as causal_model.rdi issues it (causal_network.py:145-156)"""
    ea, eb, d, uniform = task
    import load_ref
    ref = load_ref.load()
    if ref is not None:
        T = len(eb)
        x = [[v] for v in ea[:T - d]]; y = [[v] for v in eb[d:]]; z = [[v] for v in eb[d - 1:T - 1]]
        return float(ref[0].cumi(x, y, z) if uniform else ref[0].cmi(x, y, z))
    from oracle import ref_port
    return ref_port.rdi_job((ea, eb, d, uniform))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ad, spl, vel, genes = c3_adata()
    cores = os.cpu_count() or 1
    per_task = 4
    pool = cpu_pool(cores, spl, vel, genes)
    for w in range(args.warmup):
        run_cpu_c3(c3_tasks(max(cores // 4, 1), 1, 100 + w), pool)
    total_t, total_jobs = 0.0, 0
    for s in range(args.steps):
        tasks = c3_tasks(cores, per_task, s)
        _, dt = run_cpu_c3(tasks, pool)
        total_t += dt; total_jobs += cores * per_task
    v = total_jobs / total_t
    n_pairs = C3_GENES * (C3_GENES - 1)
    kind = cpu_kind()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % (C3_GENES, C3_CELLS, n_pairs),
                   "sample": "each step = %d seeded (TF, target) pairs of the workload (%d TFs x %d targets)" % (cores * per_task, cores, per_task)},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": kind,
                         "sample": "%d steps x %d pairs; %s, one process per core, each running %s"
                                   % (args.steps, cores * per_task,
                                      "unmodified reference files from baseline/_ref" if kind == "reference" else "oracle/ref_port.py",
                                      "Scribe.causal_net_dynamics_coupling(adata, TFs=[tf], Targets=[4 genes])" if kind == "reference"
                                      else "the port of its per-pair loop")},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def roofline_of(st_sum, n_steps, kernel, sm_count, peaks, peak_src, weighted=False, step_ms=None, traffic=None):
    """ALU roofline of the estimator kernel, PER LAUNCH, from the library's per-call counters summed over `n_steps` steps
    (a step of the coupling workload is several launches: the pair list is evaluated in chunks bounded by workspace)."""
    sm_clock_hz = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    alu_peak = sm_count * 128 * sm_clock_hz
    n_launch = max(int(st_sum.get("estimator_launches", 0)), 1)
    if weighted:
        n_launch = max(n_launch // 2, 1)                         # the KDE sweep and the estimator sweep count as one pair of launches
    k_ms = st_sum["estimator_ms"] / n_launch
    vk, vc, vd = (st_sum[k] / n_launch for k in ("visited_knn", "visited_count", "visited_kde"))
    pairs = st_sum["point_pairs"] / n_launch
    ops = 6.0 * vk + 10.0 * vc + (6.0 * vd if weighted else 0.0)
    achieved = ops / (k_ms * 1e-3)
    dense = (22.0 if weighted else 16.0) * pairs / (k_ms * 1e-3)
    r = {"bound": "alu", "achieved": achieved / 1e12, "peak": alu_peak / 1e12, "unit": "Tlane-op/s", "frac": achieved / alu_peak,
         "traffic": traffic, "kernel": kernel, "kernel_ms": k_ms, "launches_per_step": n_launch / max(n_steps, 1),
         "algorithmic_ops": "6 per kNN-pass visit + 10 per count-pass visit" + (" + 6 per KDE visit" if weighted else "")
                            + " (SURVEY 8d, D = 3), counted on the point pairs the sorted sweep visits",
         "visits_per_launch": {"knn": vk, "count": vc, "kde": vd, "dense_pairs_N2": pairs,
                               "visited_fraction": (vk + vc) / max(2.0 * pairs, 1.0)},
         "dense_equivalent": {"value": dense / 1e12, "frac_of_peak": dense / alu_peak,
                              "note": "%d x N^2 per job / kernel time: what a dense brute force would have to sustain for the same "
                                      "evals/s -- a speed-up figure, not a utilisation" % (22 if weighted else 16)},
         "peak_source": "SMs x 128 FP32 lanes x sm_max_mhz from MEASURED_PEAKS.json (%s); FADD issue rate confirmed 3.89/4 "
                        "warp-inst/clk/SM by tools/microbench.cu" % peak_src}
    if step_ms:
        r["kernel_share_of_step"] = st_sum["estimator_ms"] / max(n_steps, 1) / step_ms
    return r


def add_stats(acc, st):
    for k in ("estimator_ms", "visited_knn", "visited_count", "visited_kde", "point_pairs", "kernel_launches", "estimator_launches",
              "h2d_bytes", "d2h_bytes"):
        acc[k] = acc.get(k, 0) + st[k]


def ours_arm(args):
    import pandas as pd
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    os.environ["SCRIBE_B200_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    import recipes
    ad, spl, vel, genes = c3_adata()
    if world == 1 and not args.no_cpu:
        cpu_pool(os.cpu_count() or 1, spl, vel, genes)              # fork the CPU-arm workers before this process touches CUDA
    import scribe_py_b200 as S
    from scribe_py_b200 import Scribe as SC
    from scribe_py_b200 import distributed as D
    h = S._lib.default_handle()
    plan = SC._coupling_plan(ad)
    n_jobs = len(plan["src"])
    SC._upload_layers(h, ad, plan["genes"], "spliced", "velocity", True)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def l2_flush():
        flush.fill_(1)
        torch.cuda.current_stream().synchronize()        # the library runs on its own stream: the flush must be complete first

    def step_resident():
        l2_flush()
        vals = SC._coupling_execute(plan)
        return vals, h.stats()

    # the e2e step's inputs live in PINNED host memory (bench contract): the layers are copied once, outside the timed region,
    # into page-locked buffers that the library's cudaMemcpyAsync then reads at PCIe speed
    spl_pin = torch.from_numpy(spl).pin_memory().numpy()
    vel_pin = torch.from_numpy(vel).pin_memory().numpy()

    def step_e2e():
        l2_flush()
        a = recipes.FakeAnnData({"spliced": spl_pin, "velocity": vel_pin}, genes)
        S.causal_net_dynamics_coupling(a)
        return a.uns["causal_net"]["RDI"], h.stats()

    # ---- sharded == unsharded, bit for bit (N > 1): a sub-panel of the workload, all drivers of the path
    invariance = None
    if world > 1:
        sub = recipes.FakeAnnData({"spliced": spl[:, :24].copy(), "velocity": vel[:, :24].copy()}, genes[:24])
        S.causal_net_dynamics_coupling(sub)
        got = sub.uns["causal_net"]["RDI"].to_numpy(dtype=float)
        E = recipes.rdi_synthetic(10, 600, 5)
        fr = pd.DataFrame(E, index=pd.Index(["g%02d" % i for i in range(10)], name="GENE_ID"))
        m = S.causal_model(fr); r = m.rdi([1, 5]); c = m.crdi(L=1)
        real = D.world
        D.world = lambda: (0, 1)
        try:
            sub1 = recipes.FakeAnnData({"spliced": spl[:, :24].copy(), "velocity": vel[:, :24].copy()}, genes[:24])
            S.causal_net_dynamics_coupling(sub1)
            m1 = S.causal_model(fr); r1 = m1.rdi([1, 5]); c1 = m1.crdi(L=1)
        finally:
            D.world = real
        ok = np.array_equal(got, sub1.uns["causal_net"]["RDI"].to_numpy(dtype=float)) and \
            all(np.array_equal(r[d].to_numpy(), r1[d].to_numpy(), equal_nan=True) for d in (1, 5, "MAX")) and \
            np.array_equal(c.to_numpy(), c1.to_numpy(), equal_nan=True)
        t = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
        invariance = bool(t.item() == 1.0)
        SC._upload_layers(h, ad, plan["genes"], "spliced", "velocity", True)       # the workload's layers back on the device

    for _ in range(args.warmup):
        step_resident()
    props = torch.cuda.get_device_properties(local)
    sampler = ClockSampler("GPU-" + str(props.uuid) if getattr(props, "uuid", None) else None)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {}
    barrier()
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals, st = step_resident()
        add_stats(acc, st)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the public API (host AnnData layers -> result DataFrame)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    h2d = d2h = 0
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        res, st = step_e2e()
        h2d += st["h2d_bytes"] + 2 * spl.nbytes; d2h += st["d2h_bytes"] + (8 * n_jobs if world > 1 else 0)
    barrier()
    e2e_wall = time.perf_counter() - t1

    t = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, acc["estimator_ms"]], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(acc["kernel_launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, e2e_ms, est_ms_max = t.tolist()
    launches_all = tot.item()

    extras = []
    if not args.no_extra:
        extras = extra_workloads(args, h, S, l2_flush, rank, world)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        kernel = "ksg_sweepw_kernel<1,6,2,false>"
        step_ms = dev_ms / args.steps
        value = n_jobs * args.steps / (dev_ms * 1e-3)
        roof = roofline_of(acc, args.steps, kernel, h.sm_count, peaks, peak_src, step_ms=step_ms,
                           traffic=ncu_traffic("c3:" + kernel) if n_gpus == 1 else None)
        jobs_rank = n_jobs / world
        n_eff = (acc["point_pairs"] / args.steps / max(jobs_rank, 1)) ** 0.5          # rms cells kept per pair on this rank
        algo_bytes = jobs_rank * (28.0 * C3_CELLS + 48.0 * n_eff)
        roof["hbm"] = {"achieved_gbs": algo_bytes / (acc["estimator_ms"] / args.steps * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                       "note": "non-binding: per pair 3 columns + the sort order read (28 B x cells), 3 compacted columns written and "
                               "read once (48 B x cells kept, rms %.0f)" % n_eff}
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (C3_GENES, C3_CELLS, n_jobs),
                       "jobs_per_step": n_jobs, "cells": C3_CELLS, "genes": C3_GENES, "kernel_path": os.environ.get("SCRIBE_B200_PATH", "auto"),
                       "sharding": "pair list fixed; contiguous target-major slabs per rank; one all-gather of the float64 result vector",
                       "l2": "256 MiB buffer written between steps (L2 flush)",
                       "timing": "torch.cuda events around the K steps, max over ranks; wall clock agrees: %.2f ms/step" % (wall_ms / args.steps)},
            "clocks": clocks,
            "e2e": {"value": n_jobs * e2e_steps / (e2e_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(h2d / e2e_steps),
                    "d2h_bytes_per_step": int(d2h / e2e_steps), "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "scribe_py_b200.causal_net_dynamics_coupling(adata) -> adata.uns['causal_net']['RDI'] (raw layers uploaded "
                           "from pinned host memory and normalised on the device every step)"},
            "gpu_launches": int(launches_all),
            "roofline": roof,
        }
        gold = reference_goldens()
        if "c3_matrix" in gold:
            M3 = np.zeros((C3_GENES, C3_GENES))
            M3[plan["src"], plan["dst"]] = np.where(np.isnan(vals), 0.0, vals)
            sub = M3[np.ix_(recipes.C3_TFS, recipes.C3_TARGETS)]
            line["parity_vs_reference"] = {"max_abs_diff": float(np.abs(sub - gold["c3_matrix"]).max()), "pairs": int(sub.size),
                                           "source": "unmodified reference's causal_net_dynamics_coupling on 5 TFs x 8 targets of this "
                                                     "workload (tests/golden/golden_configs.npz: c3_matrix)"}
        if invariance is not None:
            line["invariance_ok"] = invariance
        if extras:
            line["extra_workloads"] = extras
        if n_gpus == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            pool = cpu_pool(cores, spl, vel, genes)
            tasks = c3_tasks(cores, 8, seed=0)
            cpu_vals, dt = run_cpu_c3(tasks, pool)
            M = np.full((C3_GENES, C3_GENES), np.nan)
            M[plan["src"], plan["dst"]] = vals
            gpu_vals = np.array([[M[tf, t] for t in tg] for tf, tg in tasks])
            kind = cpu_kind()
            line["cpu_baseline"] = {"value": cores * 8 / dt, "unit": "evals/s", "cores": cores, "kind": kind,
                                    "sample": "%d seeded (TF, target) pairs of the workload (%d TFs x 8 targets), %s, one process per core, %.1f s"
                                              % (cores * 8, cores, "unmodified reference from baseline/_ref: Scribe.causal_net_dynamics_coupling"
                                                 if kind == "reference" else "oracle/ref_port.py", dt),
                                    "parity_max_abs_diff_on_sample": float(np.abs(np.array(cpu_vals) - gpu_vals).max())}
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_workloads(args, h, S, l2_flush, rank=0, world=1):
    """C2 (one GPU only: BASELINE names it a 1-GPU config) and a C5 target slab (every GPU count, jobs sharded like the main
    workload): value (matrix resident), roofline, and on one GPU a small CPU sample each."""
    import pandas as pd
    import torch
    import torch.distributed as dist
    import recipes
    from scribe_py_b200 import distributed as D
    peaks, peak_src = measured_peaks()
    out = []
    steps = max(2, min(args.steps, 5))

    def timed(fn, n_steps):
        acc = {}
        for _ in range(2):
            fn()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); ev0.record()
        for _ in range(n_steps):
            l2_flush()
            vals = fn()
            add_stats(acc, h.stats())
        ev1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n_steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
        return vals, ms, acc

    cores = os.cpu_count() or 1
    if world == 1:
        out.append(extra_c2(args, h, S, timed, steps, peaks, peak_src, cores))
    out.append(extra_c5(args, h, S, D, timed, peaks, peak_src, cores, rank, world))
    if world == 1:
        out.append(extra_umi_knn(h, S))
    return out if rank == 0 else []


def extra_umi_knn(h, S):
    """estimator-level umi (Euclidean sweep vs the generic thread-per-query kernel) and the 625-query kNN grid regressor"""
    import recipes
    E = recipes.zscore_rows(recipes.ar_panel(16, 20000, 5))
    x, y = E[0, :-1, None], E[1, 1:, None]
    rec = {"workload": "umi(x, y) at N = 19 999 (information_estimators.py:209-277) and viz_causality's kNN grid regressor "
                       "(625 queries, plot/heatmaps.py:405-423) on one gene pair", "unit": "ms per call (host arrays in, value out)"}
    for path, key in ((S._lib.PATH_SWEEP, "umi_sweep_ms"), (S._lib.PATH_GENERIC, "umi_generic_ms")):
        opts = S._lib.make_opts(S._lib.EST_UMI, 5, "kde", 5, 0.01, path=path)
        v = h.estimate(x, y, None, opts)
        t0 = time.perf_counter()
        for _ in range(3):
            v = h.estimate(x, y, None, opts)
        rec[key] = (time.perf_counter() - t0) / 3 * 1e3
        rec[key.replace("_ms", "_value")] = v
    S.heatmaps.causality_grid(E[0], E[1])
    t0 = time.perf_counter()
    for _ in range(10):
        S.heatmaps.causality_grid(E[0], E[1])
    rec["knn_grid_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    return rec


def extra_c2(args, h, S, timed, steps, peaks, peak_src, cores):
    import pandas as pd
    import recipes
    # ---- C2: RDI all ordered pairs, 100 genes x 2000 cells, delays {1,5,10}
    G, T = 100, 2000
    E = recipes.rdi_synthetic(G, T, 2)
    frame = pd.DataFrame(E, index=pd.Index(["g%03d" % i for i in range(G)], name="GENE_ID"))
    model = S.causal_model(frame)
    plan = model._rdi_plan(C2_DELAYS)                              # uploads the matrix; it stays resident for the timed steps
    vals, ms, acc = timed(lambda: model._rdi_execute(plan), steps)
    n_jobs = len(plan["jobs"])
    t0 = time.perf_counter(); model2 = S.causal_model(frame); res = model2.rdi(C2_DELAYS); e2e_s = time.perf_counter() - t0
    rec = {"workload": "RDI all ordered pairs, 100 genes x 2,000 pseudotime-ordered cells, delays {1,5,10} (BASELINE configs[1])",
           "jobs_per_step": n_jobs, "value": n_jobs / (ms * 1e-3), "unit": "evals/s", "ms_per_step": ms, "steps": steps,
           "e2e": {"value": n_jobs / e2e_s, "unit": "evals/s", "api": "causal_model(DataFrame).rdi([1,5,10]), one call"},
           "roofline": roofline_of(acc, steps, "ksg_sweepw_kernel<1,6,2,false>", h.sm_count, peaks, peak_src, step_ms=ms,
                                   traffic=ncu_traffic("c2:ksg_sweepw_kernel<1,6,2,false>"))}
    gold = reference_goldens()
    if "c2full_rdi_1" in gold:
        per = plan["per"]
        diff, gmax, rmax = 0.0, np.full((G, G), -np.inf), np.full((G, G), -np.inf)
        for i, d in enumerate(C2_DELAYS):
            Md = np.full((G, G), np.nan); Md[plan["src"], plan["tgt"]] = vals[i * per:(i + 1) * per]
            ref = gold["c2full_rdi_%d" % d]
            diff = max(diff, float(np.nanmax(np.abs(Md - ref))))
            gmax = np.fmax(gmax, Md); rmax = np.fmax(rmax, ref)

        def top(M, n=100):
            a, b = np.nonzero(~np.eye(G, dtype=bool))
            o_ = np.lexsort((b, a, -M[a, b]))[:n]
            return list(zip(a[o_].tolist(), b[o_].tolist()))

        rec["parity_vs_reference"] = {"max_abs_diff": diff, "values": 3 * G * (G - 1), "top100_ranking_identical": top(gmax) == top(rmax),
                                      "source": "the FULL matrix of this workload from the unmodified reference (golden_configs.npz: c2full_rdi_*)"}
    if not args.no_cpu:
        rng = np.random.default_rng(0)
        sample = []
        while len(sample) < 2 * cores:
            a, b = rng.integers(0, G, 2)
            if a != b:
                sample.append((int(a), int(b), int(C2_DELAYS[len(sample) % 3])))
        pool = _POOL.get(cores)
        t0 = time.perf_counter()
        tasks = [(E[a], E[b], d, False) for a, b, d in sample]
        cpu = pool.map(_cpu_rdi_worker, tasks, chunksize=1) if pool else [_cpu_rdi_worker(t) for t in tasks]
        dt = time.perf_counter() - t0
        per = plan["per"]
        M = {d: np.full((G, G), np.nan) for d in C2_DELAYS}
        for i, d in enumerate(C2_DELAYS):
            M[d][plan["src"], plan["tgt"]] = vals[i * per:(i + 1) * per]
        gpu = np.array([M[d][a, b] for a, b, d in sample])
        rec["cpu_baseline"] = {"value": len(sample) / dt, "unit": "evals/s", "cores": cores, "kind": cpu_kind(),
                               "sample": "%d seeded jobs, %.1f s" % (len(sample), dt),
                               "parity_max_abs_diff_on_sample": float(np.abs(np.array(cpu) - gpu).max())}
    return rec


def extra_c5(args, h, S, D, timed, peaks, peak_src, cores, rank, world):
    # ---- C5 target slab: uniformised RDI (cumi, KDE bw = 0.01), 1000 genes x 20 000 cells, all sources -> 8 targets, 3 delays;
    # the same 23 976 jobs at every GPU count, sharded across the ranks (strong scaling)
    import recipes
    G, T, NT = 1000, 20000, 8
    E = recipes.zscore_rows(recipes.ar_panel(G, T, 5))
    h.upload_matrix(E)
    src = np.concatenate([np.delete(np.arange(G, dtype=np.int32), t) for t in range(NT)])
    dst = np.repeat(np.arange(NT, dtype=np.int32), G - 1)
    jobs = np.zeros(len(src) * 3, dtype=S._lib.JOB_DTYPE)
    for i, d in enumerate(C2_DELAYS):
        sl = slice(i * len(src), (i + 1) * len(src))
        jobs["src"][sl] = src; jobs["dst"][sl] = dst; jobs["delay"][sl] = d
    opts = S._lib.make_opts(S._lib.EST_CUMI, 5)

    def evaluate(lo, hi, out_dev_ptr=None):
        return h.lagged_jobs(jobs[lo:hi], opts, out_dev_ptr=out_dev_ptr)

    vals, ms, acc = timed(lambda: D.sharded_evaluate(len(jobs), evaluate, h), 2)
    rec = {"workload": "target slab of the uniformised-RDI scale sweep (BASELINE configs[4]): cumi (KDE bw=0.01), 1,000 genes x 20,000 cells, "
                       "all 999 sources -> %d targets x delays {1,5,10}, sharded over %d GPU(s); the full sweep is 125 such slabs" % (NT, world),
           "jobs_per_step": len(jobs), "value": len(jobs) / (ms * 1e-3), "unit": "evals/s", "ms_per_step": ms, "steps": 2, "n_gpus": world,
           "full_config_extrapolated_s": 2997000 / (len(jobs) / (ms * 1e-3)),
           "roofline": roofline_of(acc, 2, "ksg_sweepw_kernel<1,6,2,true> + kde_sweep_kernel<2,2>", h.sm_count, peaks, peak_src, weighted=True,
                                   step_ms=ms, traffic=ncu_traffic("c5:ksg_sweepw_kernel<1,6,2,true>"))}
    if world >= 8 and not args.no_full_c5:
        # north_star's target run: ALL ordered pairs x 3 delays of the 1000-gene panel (2 997 000 cumi evaluations) through the
        # public API, one call, on this box's GPUs; parity on the pairs the unmodified reference was run on
        import pandas as pd
        import torch
        import torch.distributed as dist
        frame = pd.DataFrame(E, index=pd.Index(["g%04d" % i for i in range(G)], name="GENE_ID"))
        model = S.causal_model(frame)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = model.rdi(C2_DELAYS, uniformization=True)
        dist.barrier(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        n_full = G * (G - 1) * 3
        full = {"jobs": n_full, "wall_s": wall, "value": n_full / wall, "unit": "evals/s", "n_gpus": world,
                "api": "causal_model(DataFrame 1000 x 20000).rdi([1,5,10], uniformization=True), one call, matrix upload included"}
        gold = reference_goldens()
        if "c5_urdi_1" in gold:
            gi = recipes.C5_GENES
            full["parity_vs_reference"] = {
                "max_abs_diff": max(float(np.nanmax(np.abs(res[d].to_numpy(dtype=float)[np.ix_(gi, gi)] - gold["c5_urdi_%d" % d])))
                                    for d in C2_DELAYS),
                "values": 36, "source": "unmodified reference's rdi(uniformization=True) on 4 genes of this panel (golden_configs.npz: c5_urdi_*)"}
        rec["full_config"] = full
    if not args.no_cpu and world == 1:
        # the reference's cumi at N = 20 000 takes ~10 s alone and far longer when every core runs one (memory-bound kd-tree
        # KDE): a handful of jobs keeps this leg bounded
        sample = [(int(src[i]), int(dst[i]), 1) for i in np.random.default_rng(1).choice(len(src), min(cores, 4), replace=False)]
        pool = _POOL.get(cores)
        t0 = time.perf_counter()
        tasks = [(E[a], E[b], d, True) for a, b, d in sample]
        cores = len(sample)
        cpu = pool.map(_cpu_rdi_worker, tasks, chunksize=1) if pool else [_cpu_rdi_worker(t) for t in tasks]
        dt = time.perf_counter() - t0
        look = {(int(jobs["src"][i]), int(jobs["dst"][i]), int(jobs["delay"][i])): vals[i] for i in range(len(src))}
        gpu = np.array([look[s_] for s_ in sample])
        rec["cpu_baseline"] = {"value": len(sample) / dt, "unit": "evals/s", "cores": cores, "kind": cpu_kind(),
                               "sample": "%d seeded cumi jobs at N = 19 999, %.1f s" % (len(sample), dt),
                               "parity_max_abs_diff_on_sample": float(np.abs(np.array(cpu) - gpu).max())}
    return rec


def _json_only_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr; return a file object on the real
    stdout so that the JSON line is the only thing the driver reads there."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra workloads (C2, C5 slab)")
    ap.add_argument("--no-full-c5", action="store_true", help="at 8 GPUs: skip the full C5 sweep (2,997,000 evaluations, ~15 s)")
    args = ap.parse_args()
    global OUT
    OUT = _json_only_stdout()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
