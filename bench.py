#!/usr/bin/env python
"""bench.py -- gene-pair x delay CMI evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1], "RDI all ordered pairs, 100 genes x 2,000
pseudotime-ordered cells, delays {1,5,10}" = 29,700 cmi evaluations per step at N = 1.  For N > 1 the gene panel
grows to round(100*sqrt(N)) genes so that the work per GPU stays ~29,700 evaluations (weak scaling); ordered pairs
are sharded across ranks and one all-gather of the result vector is the only collective.

A "step" is one pass of the hot path over the whole job table.
  value : evaluations/s with the expression matrix already resident in HBM (job table -> results, incl. the
          all-gather), timed with CUDA events, max over ranks.
  e2e   : the same through the public API, causal_model.rdi(delays), from a host DataFrame: host->device copy of
          the matrix and device->host read of the results inside the timed region.
  roofline     : the dominant kernel (ksg_sweepw_kernel) against the ALU roofline north_star names: algorithmic
                 FP32-pipe lane-ops (16 per point pair for D=3) / kernel time / (SMs x 128 lanes x clock).
  cpu_baseline : the reference-shaped CPU port (oracle/ref_port.py, same scipy calls and per-sample loops as the
                 reference) on a bounded sample of the same jobs, all host cores.
--impl reference times that CPU port alone (bench.py's one other use of oracle/).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

OUT = sys.stdout
DELAYS = [1, 5, 10]
CELLS = 2000
OPS_PER_PAIR = 16            # cmi/RDI, D = 3: 6 (kNN pass) + 10 (count pass) FP32-pipe lane-ops
FP64_OPS_PER_PAIR = 13       # executed on the FP64 pipe: 3 DADD + 3 DSETP (kNN) + 3 DADD + 4 DSETP (count)


def genes_for(n_gpus):
    return int(round(100 * np.sqrt(n_gpus)))


def make_expression(G, T, seed=2):
    import recipes
    return recipes.rdi_synthetic(G, T, seed)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic_r1.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json"))).get(kernel)
    except Exception:
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
# CPU arm: the reference-shaped port on all host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_jobs(E, n_jobs, seed=0):
    rng = np.random.default_rng(seed)
    G = E.shape[0]
    out = []
    while len(out) < n_jobs:
        a, b = rng.integers(0, G, 2)
        if a != b:
            out.append((int(a), int(b), int(DELAYS[len(out) % len(DELAYS)])))
    return out


_POOL = {}


def cpu_pool(cores):
    """One fork pool per process, created outside any timed region."""
    from multiprocessing import get_context
    if cores > 1 and cores not in _POOL:
        _POOL[cores] = get_context("fork").Pool(cores)
        _POOL[cores].map(abs, range(cores))          # spin the workers up
    return _POOL.get(cores)


def run_cpu_sample(E, jobs, cores):
    from oracle import ref_port
    args = [(E[a], E[b], d, False) for a, b, d in jobs]
    pool = cpu_pool(cores)
    t0 = time.perf_counter()
    if pool is not None:
        vals = pool.map(ref_port.rdi_job, args, chunksize=1)
    else:
        vals = [ref_port.rdi_job(a) for a in args]
    return np.array(vals), time.perf_counter() - t0


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    G = genes_for(args.gpus)
    E = make_expression(G, CELLS)
    cores = os.cpu_count() or 1
    per_step = max(4 * cores, 32)
    cpu_pool(cores)
    for w in range(args.warmup):
        run_cpu_sample(E, cpu_jobs(E, min(per_step, cores), seed=100 + w), cores)
    total_t, total_jobs = 0.0, 0
    for s in range(args.steps):
        _, dt = run_cpu_sample(E, cpu_jobs(E, per_step, seed=s), cores)
        total_t += dt; total_jobs += per_step
    v = total_jobs / total_t
    line = {
        "impl": "reference", "metric": "gene-pair x delay CMI evaluations/sec", "value": v, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "RDI all ordered pairs, %d genes x %d cells, delays {1,5,10} (BASELINE configs[1]); "
                               "each step = %d sampled (source,target,delay) jobs" % (G, CELLS, per_step)},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d seeded random jobs of the workload, oracle/ref_port.py "
                                   "(cKDTree per-sample loops as the reference), multiprocessing over %d cores"
                                   % (args.steps, per_step, cores)},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
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

    import scribe_py_b200 as S
    h = S._lib.default_handle()
    G = genes_for(n_gpus)
    E = make_expression(G, CELLS)
    genes = ["g%03d" % i for i in range(G)]
    frame = pd.DataFrame(E, index=pd.Index(genes, name="GENE_ID"))
    model = S.causal_model(frame)
    plan = model._rdi_plan(DELAYS)
    n_jobs = len(plan["jobs"])
    h.upload_matrix(plan["E"], plan["run_off"])

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.fill_(1)
        torch.cuda.current_stream().synchronize()        # the library runs on its own stream: the flush must be complete first
        vals = model._rdi_execute(plan)
        return vals, h.stats()

    def step_e2e():
        flush.fill_(1)
        torch.cuda.current_stream().synchronize()
        m = S.causal_model(frame)
        res = m.rdi(DELAYS)
        return res, h.stats()

    for _ in range(args.warmup):
        step_resident()
    props = torch.cuda.get_device_properties(local)
    sampler = ClockSampler("GPU-" + str(props.uuid) if getattr(props, "uuid", None) else None)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = est_launches = pairs = v_knn = v_cnt = 0
    est_ms = 0.0
    barrier()
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals, st = step_resident()
        launches += st["kernel_launches"]; est_launches += st["estimator_launches"]; est_ms += st["estimator_ms"]
        pairs += st["point_pairs"]; v_knn += st["visited_knn"]; v_cnt += st["visited_count"]
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the public API (host DataFrame -> result DataFrames)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    h2d = d2h = 0
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        res, st = step_e2e()
        h2d += st["h2d_bytes"] + plan["E"].nbytes; d2h += st["d2h_bytes"]
    barrier()
    e2e_wall = time.perf_counter() - t1

    t = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, est_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(launches), float(pairs)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, e2e_ms, est_ms_max = t.tolist()
    launches_all, pairs_all = tot.tolist()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sm_clock_hz = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
        alu_peak = h.sm_count * 128 * sm_clock_hz                  # FP32 lanes x clock, lane-ops/s per GPU
        fp64_peak = h.sm_count * 64 * sm_clock_hz                  # measured 64 DADD/DSETP lanes/clk/SM (microbench)
        # dominant kernel on THIS rank.  Algorithmic work = the point-pair visits the kernel actually makes (the
        # sorted sweep visits only the z-window of each query) x SURVEY 8(d)'s per-visit lane-ops: 6 in the kNN pass,
        # 10 in the count pass (D = 3).
        path = os.environ.get("SCRIBE_B200_PATH", "auto")
        v1 = os.environ.get("SCRIBE_B200_SWEEP_V1", "0") not in ("", "0")
        windowed = path in ("auto", "sweep") and not v1
        kernel = ("ksg_sweepw_kernel<1,6,2,false>" if windowed else "ksg_sweep_kernel<1,1,1,6,2,false>") \
            if path in ("auto", "sweep") else "ksg_fused_kernel<1,1,1,6,2,false>"
        k_ms = est_ms / max(est_launches, 1)
        pairs_per_launch = pairs / max(est_launches, 1)
        vk, vc = v_knn / max(est_launches, 1), v_cnt / max(est_launches, 1)
        achieved = (6.0 * vk + 10.0 * vc) / (k_ms * 1e-3)
        # executed on the FP64 pipe per visit: first sweep / dense 6 (kNN) + 7 (count); windowed sweep 0 + 4 (its kNN scan
        # is an FP32 pre-filter: 2 FADD + 2 FSETP, exact FP64 re-test only for marked candidates)
        achieved64 = ((0.0 * vk + 4.0 * vc) if windowed else (6.0 * vk + 7.0 * vc)) / (k_ms * 1e-3)
        dense_equiv = OPS_PER_PAIR * pairs_per_launch / (k_ms * 1e-3)
        algo_bytes = 8.0 * 3 * CELLS * (n_jobs / world) + 8.0 * n_jobs / world
        value = n_jobs * args.steps / (dev_ms * 1e-3)
        line = {
            "metric": "gene-pair x delay CMI evaluations/sec", "value": value, "unit": "evals/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "RDI all ordered pairs, %d genes x %d pseudotime-ordered cells, delays {1,5,10} "
                                   "(BASELINE configs[1]%s): %d cmi evaluations per step, k=5"
                                   % (G, CELLS, "" if n_gpus == 1 else ", gene panel scaled by sqrt(n_gpus) for weak scaling", n_jobs),
                       "jobs_per_step": n_jobs, "cells": CELLS, "genes": G, "delays": DELAYS, "kernel_path": os.environ.get("SCRIBE_B200_PATH", "auto"),
                       "l2": "256 MiB buffer written between steps (L2 flush); job working set itself is L2-resident by design",
                       "timing": "torch.cuda events around the K steps, max over ranks; wall clock agrees: %.2f ms/step" % (wall_ms / args.steps)},
            "clocks": clocks,
            "e2e": {"value": n_jobs * e2e_steps / (e2e_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(h2d / e2e_steps),
                    "d2h_bytes_per_step": int(d2h / e2e_steps), "steps": e2e_steps,
                    "api": "scribe_py_b200.causal_model(DataFrame).rdi([1,5,10]) -> dict of DataFrames"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "alu", "achieved": achieved / 1e12, "peak": alu_peak / 1e12, "unit": "Tlane-op/s",
                         "frac": achieved / alu_peak, "traffic": ncu_traffic(kernel) if n_gpus == 1 else None,
                         "kernel": kernel, "kernel_ms": k_ms, "kernel_share_of_step": est_ms / dev_ms,
                         "algorithmic_ops": "6 per kNN-pass visit + 10 per count-pass visit (SURVEY 8d, D=3); the windowed sweep "
                                            "executes 5 + 6 per visit (FP32 pre-filter; n_z from the rank window, no z test)",
                         "visits_per_launch": {"knn": vk, "count": vc, "dense_pairs_N2": pairs_per_launch,
                                               "visited_fraction": (vk + vc) / max(2.0 * pairs_per_launch, 1.0)},
                         "dense_equivalent": {"value": dense_equiv / 1e12, "frac_of_peak": dense_equiv / alu_peak,
                                              "note": "16 x N^2 per job / kernel time: what a dense brute force would have to "
                                                      "sustain for the same evals/s -- a speed-up figure, not a utilisation"},
                         "peak_source": "SMs x 128 FP32 lanes x sm_max_mhz from MEASURED_PEAKS.json (%s); FADD issue rate "
                                        "confirmed 3.89/4 warp-inst/clk/SM by tools/microbench.cu" % peak_src,
                         "fp64_pipe": {"achieved": achieved64 / 1e12, "peak": fp64_peak / 1e12, "frac": achieved64 / fp64_peak,
                                       "unit": "Tlane-op/s", "note": ("windowed sweep: FP32 pre-filter in the kNN pass (0 FP64 per visit, exact "
                                                "re-test of marked candidates only), 2 DADD + 2 DSETP per count visit"
                                                if windowed else
                                                "the kernel executes on the FP64 pipe (exact counts): 6 DADD/DSETP per kNN visit, "
                                                "7 per count visit") + "; peak = 64 lanes/clk/SM measured"},
                         "hbm": {"achieved_gbs": algo_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": "non-binding: 3 columns x 8 B x N read per job"}},
        }
        if n_gpus == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            sample = cpu_jobs(E, max(24 * cores, 192), seed=0)
            cpu_pool(cores)
            cpu_vals, dt = run_cpu_sample(E, sample, cores)
            M = {d: np.full((G, G), np.nan) for d in DELAYS}
            per = plan["per"]
            for i, d in enumerate(DELAYS):
                M[d][plan["src"], plan["tgt"]] = vals[i * per:(i + 1) * per]
            gpu_vals = np.array([M[d][a, b] for a, b, d in sample])
            line["cpu_baseline"] = {"value": len(sample) / dt, "unit": "evals/s", "cores": cores, "kind": "port",
                                    "sample": "%d seeded random (source,target,delay) jobs of the workload, oracle/ref_port.py "
                                              "(cKDTree per-sample loops as the reference), multiprocessing over %d cores, %.1f s"
                                              % (len(sample), cores, dt),
                                    "parity_max_abs_diff_on_sample": float(np.abs(cpu_vals - gpu_vals).max())}
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    global OUT
    OUT = _json_only_stdout()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
