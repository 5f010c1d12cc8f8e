#!/usr/bin/env python
"""Throughput and sampled parity of every named BASELINE.json config on one GPU.

    python tools/run_configs.py [c1 c2 c3 c4 c5] [--scale S]   (S < 1 shrinks the gene count of the big configs)

For each config: GPU wall time through the public API, evaluations/s, device time of the estimator kernels, the
fraction of dense point pairs the sorted sweep visits, and the max |difference| against the reference-shaped CPU port
(oracle/ref_port.py) on a few sampled jobs.  Writes one JSON line per config."""
import argparse, json, os, sys, time, warnings
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import recipes
import scribe_py_b200 as S
from scribe_py_b200 import _lib
from oracle import ref_port, ksg_oracle


ar_panel = recipes.ar_panel          # the C4 / C5 generator lives with the other seeded recipes (tests/golden/recipes.py)


def cpu_check(fn, argsets):
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:     # multi-GPU runs: parity is established at 1 GPU + tools/check_multi_gpu.py
        return np.full(len(argsets), np.nan)
    from multiprocessing import get_context
    with get_context("fork").Pool(min(len(argsets), os.cpu_count() or 1)) as p:
        return np.array(p.map(fn, argsets))


def _cmi_job(a):
    return ref_port.cmi(*a)


def _cumi_job(a):
    return ref_port.cumi(*a)


def report(name, n_jobs, wall, stats, extra):
    line = {"config": name, "jobs": int(n_jobs), "wall_s": wall, "evals_per_s": n_jobs / wall,
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1"))}
    line.update(extra)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(line), flush=True)
    return line


def barrier():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch, torch.distributed as dist
        dist.barrier(); torch.cuda.synchronize()


def stats_acc():
    return {"estimator_ms": 0.0, "point_pairs": 0, "visited_knn": 0, "visited_count": 0, "visited_kde": 0, "kernel_launches": 0}


def run_c1():
    spl, vel, genes = recipes.gv_c()
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, genes)
    S.causal_net_dynamics_coupling(ad)              # warm-up
    t = time.perf_counter(); S.causal_net_dynamics_coupling(ad); wall = time.perf_counter() - t
    M = ad.uns["causal_net"]["RDI"].to_numpy(dtype=float)
    ref = ksg_oracle.dynamics_coupling(spl, vel)
    st = _lib.default_handle().stats()
    return report("C1 dynamics coupling 20 genes x 500 cells", 380, wall, st,
                  {"parity_max_abs_diff": float(np.abs(M - ref).max()), "parity_jobs": 380, "estimator_ms": st["estimator_ms"]})


def run_rdi(name, G, T, delays, uniform, seed, n_check, L=None):
    E = ar_panel(G, T, seed)
    if uniform:
        E = (E - E.mean(1, keepdims=True)) / E.std(1, keepdims=True)
    genes = ["g%04d" % i for i in range(G)]
    m = S.causal_model(pd.DataFrame(E, index=pd.Index(genes, name="GENE_ID")))
    h = _lib.default_handle()
    barrier(); t = time.perf_counter(); r = m.rdi(delays, uniformization=uniform); barrier(); wall = time.perf_counter() - t
    st = h.stats()
    n_jobs = G * (G - 1) * len(delays)
    rng = np.random.default_rng(0)
    args, got = [], []
    for _ in range(n_check):
        a, b = rng.choice(G, 2, replace=False); d = delays[rng.integers(len(delays))]
        x, y, z = ksg_oracle.lagged_columns([E[a]], [E[b]], d)
        args.append((x, y, z)); got.append(r[d].iloc[a, b])
    ref = cpu_check(_cumi_job if uniform else _cmi_job, args)
    extra = {"parity_max_abs_diff": float(np.nanmax(np.abs(ref - np.array(got)))) if WORLD == 1 else None, "parity_jobs": n_check,
             "estimator_ms": st["estimator_ms"], "visited_fraction": (st["visited_knn"] + st["visited_count"]) / max(2 * st["point_pairs"], 1),
             "note": "stats of the last chunk-call only" if False else ""}
    out = [report(name, n_jobs, wall, st, extra)]
    if L:
        barrier(); t = time.perf_counter(); c = m.crdi(L=L, uniformization=uniform); barrier(); wall = time.perf_counter() - t
        st = h.stats()
        # parity of sampled cRDI jobs through the oracle's conditioning-set logic
        res = {d: r[d].to_numpy(dtype=float) for d in delays}
        val, arg = ksg_oracle.max_over_delays(res, delays)
        jobs = ksg_oracle.crdi_jobs(val, arg, L)
        expr = [[E[g]] for g in range(G)]
        args, got = [], []
        for _ in range(n_check):
            a, b = rng.choice(G, 2, replace=False)
            d, cn, cd = jobs[(int(a), int(b))]
            args.append(ksg_oracle.crdi_columns(expr, int(a), int(b), d, cn, cd)); got.append(c.iloc[a, b])
        ref = cpu_check(_cumi_job if uniform else _cmi_job, args)
        out.append(report(name + " -> crdi(L=%d)" % L, G * (G - 1), wall, st,
                          {"parity_max_abs_diff": float(np.abs(ref - np.array(got)).max()) if WORLD == 1 else None, "parity_jobs": n_check,
                           "estimator_ms": st["estimator_ms"],
                           "visited_fraction": (st["visited_knn"] + st["visited_count"]) / max(2 * st["point_pairs"], 1)}))
    return out


def run_c3(G, N, n_check):
    spl, vel, genes = recipes.coupling_panel(G, N, 3)
    rng = np.random.default_rng(3)
    ad = recipes.FakeAnnData({"spliced": spl, "velocity": vel}, genes)
    h = _lib.default_handle()
    S.causal_net_dynamics_coupling(recipes.FakeAnnData({"spliced": spl[:, :8], "velocity": vel[:, :8]}, genes[:8]))   # warm-up
    walls = []
    for _ in range(3):              # the first full-size call also pays for the growth of the device workspaces
        barrier(); t = time.perf_counter(); S.causal_net_dynamics_coupling(ad); barrier(); walls.append(time.perf_counter() - t)
    wall = float(np.median(walls))          # a drop-in user calls once: the first call (workspace growth included) is reported beside it
    st = h.stats()
    M = ad.uns["causal_net"]["RDI"].to_numpy(dtype=float)
    s = (spl - spl.mean(0)) / (spl.max(0) - spl.min(0)); v = (vel - vel.mean(0)) / (vel.max(0) - vel.min(0))
    args, got = [], []
    for _ in range(n_check):
        a, b = rng.choice(G, 2, replace=False)
        x, z = s[:, a], s[:, b]; y = s[:, b] + v[:, b]
        keep = ((x + y) + z) > 0
        args.append((x[keep], y[keep], z[keep])); got.append(M[a, b])
    ref = cpu_check(_cmi_job, args)
    return report("C3 dynamics coupling %d genes x %d cells" % (G, N), G * (G - 1), wall, st,
                  {"parity_max_abs_diff": float(np.abs(ref - np.array(got)).max()) if WORLD == 1 else None, "parity_jobs": n_check,
                   "estimator_ms_last_chunk": st["estimator_ms"], "mean_cells_kept": float(np.mean([len(a[0]) for a in args])),
                   "first_call_s": walls[0], "wall_s_each_call": walls})


def init_distributed():
    """under torchrun: one rank per GPU; the drivers then shard the gene pairs and all-gather the results"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch, torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        os.environ["SCRIBE_B200_DEVICE"] = str(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return int(os.environ.get("RANK", "0")), world


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    RANK, WORLD = init_distributed()
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sc = a.scale
    lines = []
    h = _lib.default_handle()
    if RANK == 0:
        print(json.dumps({"device": h.device_name, "sms": h.sm_count, "path": os.environ.get("SCRIBE_B200_PATH", "auto"), "n_gpus": WORLD}), flush=True)
    for c in a.configs:
        if c == "c1": lines.append(run_c1())
        if c == "c2": lines += run_rdi("C2 RDI 100 genes x 2000 cells x delays {1,5,10}", 100, 2000, [1, 5, 10], False, 2, 32)
        if c == "c3": lines.append(run_c3(max(8, int(500 * sc)), 10000, 16))
        if c == "c4": lines += run_rdi("C4 RDI+cRDI(L=2) %d genes x 5000 cells" % max(8, int(200 * sc)), max(8, int(200 * sc)), 5000, [1, 5, 10], False, 4, 16, L=2)
        if c == "c5": lines += run_rdi("C5 uniformised RDI (cumi) %d genes x 20000 cells x 3 delays" % max(8, int(1000 * sc)), max(8, int(1000 * sc)), 20000, [1, 5, 10], True, 5, 4)
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
    if a.out and RANK == 0:
        with open(a.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")
