#!/usr/bin/env python
"""One short pass of a named workload for ncu: python tools/prof_case.py c2|c3|c5 [reps]
c2: RDI 100 genes x 2000 cells x 3 delays; c3: dynamics coupling, first 60 genes of the C3 panel (10 000 cells);
c5: uniformised RDI (cumi), 12 genes of the z-scored C5 panel (20 000 cells), delay 1."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import recipes
from scribe_py_b200 import _lib

what = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = _lib.default_handle()


def panel_jobs(G, delays):
    src, dst = np.nonzero(~np.eye(G, dtype=bool))
    jobs = np.zeros(len(src) * len(delays), dtype=_lib.JOB_DTYPE)
    for i, d in enumerate(delays):
        sl = slice(i * len(src), (i + 1) * len(src))
        jobs["src"][sl] = src; jobs["dst"][sl] = dst; jobs["delay"][sl] = d
    return jobs


if what == "c2":
    h.upload_matrix(recipes.rdi_synthetic(100, 2000, 2))
    jobs, opts = panel_jobs(100, [1, 5, 10]), _lib.make_opts(_lib.EST_CMI, 5)
    run = lambda: h.lagged_jobs(jobs, opts)
elif what == "c5":
    h.upload_matrix(recipes.zscore_rows(recipes.ar_panel(1000, 20000, 5)[:12]))
    jobs, opts = panel_jobs(12, [1]), _lib.make_opts(_lib.EST_CUMI, 5)
    run = lambda: h.lagged_jobs(jobs, opts)
else:
    G = 60
    spl, vel, genes = recipes.coupling_panel(G, 10000, 3)
    s = (spl - spl.mean(0)) / (spl.max(0) - spl.min(0)); v = (vel - vel.mean(0)) / (vel.max(0) - vel.min(0))
    h.upload_coupling(np.ascontiguousarray(s.T), np.ascontiguousarray((s + v).T))
    src, dst = np.nonzero(~np.eye(G, dtype=bool))
    opts = _lib.make_opts(_lib.EST_CMI, 5)
    run = lambda: h.coupling_jobs(src.astype(np.int32), dst.astype(np.int32), opts, True)
for _ in range(reps):
    out = run()
st = h.stats()
print("%s: jobs=%d est_ms=%.3f total_ms=%.3f launches=%d checksum=%.12f" % (what, st["jobs"], st["estimator_ms"], st["total_ms"], st["kernel_launches"], float(np.nansum(out))))
