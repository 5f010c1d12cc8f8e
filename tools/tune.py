#!/usr/bin/env python
"""Kernel timing sweep: python tools/tune.py  (env SCRIBE_B200_Q=1|2|4 overrides queries per thread)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import recipes
import scribe_py_b200 as S
from scribe_py_b200 import _lib

def run(G, T, delays, est, reps=2):
    E = recipes.rdi_synthetic(G, T, 2) if T <= 4000 else np.random.default_rng(0).standard_normal((G, T)).cumsum(axis=1) * 0.05 + np.random.default_rng(1).standard_normal((G, T))
    h = _lib.default_handle()
    h.upload_matrix(E)
    src, dst = np.nonzero(~np.eye(G, dtype=bool))
    jobs = np.zeros(len(src) * len(delays), dtype=_lib.JOB_DTYPE)
    for i, d in enumerate(delays):
        sl = slice(i * len(src), (i + 1) * len(src))
        jobs["src"][sl] = src; jobs["dst"][sl] = dst; jobs["delay"][sl] = d
    opts = _lib.make_opts(est, 5)
    best = None
    for _ in range(reps + 1):
        v = h.lagged_jobs(jobs, opts)
        st = h.stats()
        if best is None or st["estimator_ms"] < best["estimator_ms"]:
            best = st
    pairs = best["point_pairs"]
    clk = 1.965e9
    cps = best["estimator_ms"] * 1e-3 * clk * 148 * 4 / (pairs / 32)        # SMSP-cycles per warp-step (32 point pairs)
    vis = (best["visited_knn"] + best["visited_count"]) / (2.0 * pairs)
    print("path=%s vis=%.3f clk/visited-step=%.1f " % (os.environ.get("SCRIBE_B200_PATH", "auto"), vis, cps / max(vis, 1e-9)), end="")
    print("Q=%s est=%d G=%d T=%d jobs=%d  est_ms=%.2f total_ms=%.2f  evals/s=%.0f  pairs/s=%.3e  clk/warp-step=%.2f  checksum=%.12f"
          % (os.environ.get("SCRIBE_B200_Q", "auto"), est, G, T, len(jobs), best["estimator_ms"], best["total_ms"],
             len(jobs) / (best["total_ms"] * 1e-3), pairs / (best["estimator_ms"] * 1e-3), cps, float(np.nansum(v))), flush=True)

def run_mi(G, T, reps=2):
    """all ordered pairs of the batched MI network (other_estimators.mi -> scribe_mi_jobs)"""
    E = recipes.rdi_synthetic(G, T, 2) if T <= 4000 else np.random.default_rng(0).standard_normal((G, T)).cumsum(axis=1) * 0.05 + np.random.default_rng(1).standard_normal((G, T))
    h = _lib.default_handle()
    h.upload_matrix(E)
    src, dst = np.nonzero(~np.eye(G, dtype=bool))
    opts = _lib.make_opts(_lib.EST_MI, 5)
    best = None
    for _ in range(reps + 1):
        v = h.mi_jobs(src.astype(np.int32), dst.astype(np.int32), opts)
        st = h.stats()
        if best is None or st["estimator_ms"] < best["estimator_ms"]:
            best = st
    print("mi path=%s G=%d T=%d jobs=%d est_ms=%.2f total_ms=%.2f evals/s=%.0f visited=%.3f checksum=%.12f"
          % (os.environ.get("SCRIBE_B200_PATH", "auto"), G, T, len(src), best["estimator_ms"], best["total_ms"],
             len(src) / (best["total_ms"] * 1e-3), (best["visited_knn"] + best["visited_count"]) / (2.0 * best["point_pairs"]),
             float(np.nansum(v))), flush=True)


def run_panel(name, E, delays, est, reps=2):
    """all ordered pairs x delays of an expression panel through scribe_lagged_jobs"""
    G = E.shape[0]
    h = _lib.default_handle()
    h.upload_matrix(E)
    src, dst = np.nonzero(~np.eye(G, dtype=bool))
    jobs = np.zeros(len(src) * len(delays), dtype=_lib.JOB_DTYPE)
    for i, d in enumerate(delays):
        sl = slice(i * len(src), (i + 1) * len(src))
        jobs["src"][sl] = src; jobs["dst"][sl] = dst; jobs["delay"][sl] = d
    opts = _lib.make_opts(est, 5)
    best = None
    for _ in range(reps + 1):
        v = h.lagged_jobs(jobs, opts)
        st = h.stats()
        if best is None or st["estimator_ms"] < best["estimator_ms"]:
            best = st
    print("AB %-12s jobs=%d est_ms=%.3f total_ms=%.3f visited=%.4f checksum=%.12f"
          % (name, len(jobs), best["estimator_ms"], best["total_ms"],
             (best["visited_knn"] + best["visited_count"]) / max(2.0 * best["point_pairs"], 1), float(np.nansum(v))), flush=True)


def run_coupling(name, G, N, reps=2):
    """dynamics coupling (C3 generator), all ordered pairs of the first G genes"""
    spl, vel, genes = recipes.coupling_panel(G, N, 3)
    s = (spl - spl.mean(0)) / (spl.max(0) - spl.min(0)); v = (vel - vel.mean(0)) / (vel.max(0) - vel.min(0))
    h = _lib.default_handle()
    h.upload_coupling(np.ascontiguousarray(s.T), np.ascontiguousarray((s + v).T))
    src, dst = np.nonzero(~np.eye(G, dtype=bool))
    opts = _lib.make_opts(_lib.EST_CMI, 5)
    best = None
    for _ in range(reps + 1):
        out = h.coupling_jobs(src.astype(np.int32), dst.astype(np.int32), opts, True)
        st = h.stats()
        if best is None or st["estimator_ms"] < best["estimator_ms"]:
            best = st
    print("AB %-12s jobs=%d est_ms=%.3f total_ms=%.3f visited=%.4f checksum=%.12f"
          % (name, len(src), best["estimator_ms"], best["total_ms"],
             (best["visited_knn"] + best["visited_count"]) / max(2.0 * best["point_pairs"], 1), float(np.nansum(out))), flush=True)


def run_ab():
    """the three workloads the kernel variants are compared on (tools/ab.py)"""
    run_panel("c2", recipes.rdi_synthetic(100, 2000, 2), [1, 5, 10], _lib.EST_CMI)
    run_coupling("c3sub", 80, 10000)
    E5 = recipes.zscore_rows(recipes.ar_panel(1000, 20000, 5)[:16])
    run_panel("c5slab_cumi", E5, [1], _lib.EST_CUMI, reps=1)
    run_panel("c5slab_cmi", E5, [1], _lib.EST_CMI, reps=1)
    E4 = recipes.ar_panel(200, 5000, 4)[:40]
    run_panel("c4sub", E4, [1, 5, 10], _lib.EST_CMI, reps=1)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "cmi"
    if what == "ab":
        run_ab()
    elif what == "cmi":
        run(60, 2000, [1, 5, 10], _lib.EST_CMI)
        run(24, 20000, [1], _lib.EST_CMI)
    elif what == "cumi":
        run(40, 2000, [1, 5, 10], _lib.EST_CUMI)
        run(16, 20000, [1], _lib.EST_CUMI)
    elif what == "c5":
        run(16, 20000, [1], _lib.EST_CMI)
        run(16, 20000, [1], _lib.EST_CUMI)
        run(30, 5000, [1], _lib.EST_CMI)
        run(30, 5000, [1], _lib.EST_CUMI)
    elif what == "big":
        run(24, 20000, [1], _lib.EST_CMI, reps=1)
    elif what == "c5one":
        run(16, 20000, [1], _lib.EST_CUMI, reps=1)
    elif what == "mi":
        for T, G in ((2000, 60), (20000, 16)):
            run_mi(G, T, reps=1)
    elif what == "matrix":
        for T, G in ((1000, 60), (2000, 60), (5000, 30), (10000, 20), (20000, 16)):
            for est in (_lib.EST_CMI, _lib.EST_CUMI):
                run(G, T, [1, 5], est, reps=1)
