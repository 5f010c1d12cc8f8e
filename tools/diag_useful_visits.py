import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import recipes
from scribe_py_b200 import _lib
h = _lib.default_handle()
def report(name, st):
    print("%s: count-pass visits %.4g (lane x candidate), inside the lane's own window %.4g -> useful %.1f%%; knn visits %.4g; dense pairs %.4g"
          % (name, st["visited_count"], st["visited_kde"], 100.0 * st["visited_kde"] / st["visited_count"], st["visited_knn"], st["point_pairs"]))
G = 100
h.upload_matrix(recipes.rdi_synthetic(G, 2000, 2))
src, dst = np.nonzero(~np.eye(G, dtype=bool))
jobs = np.zeros(len(src) * 3, dtype=_lib.JOB_DTYPE)
for i, d in enumerate([1, 5, 10]):
    sl = slice(i * len(src), (i + 1) * len(src)); jobs["src"][sl] = src; jobs["dst"][sl] = dst; jobs["delay"][sl] = d
h.lagged_jobs(jobs, _lib.make_opts(_lib.EST_CMI, 5)); report("c2", h.stats())
G = 60
spl, vel, genes = recipes.coupling_panel(G, 10000, 3)
s = (spl - spl.mean(0)) / (spl.max(0) - spl.min(0)); v = (vel - vel.mean(0)) / (vel.max(0) - vel.min(0))
h.upload_coupling(np.ascontiguousarray(s.T), np.ascontiguousarray((s + v).T))
src, dst = np.nonzero(~np.eye(G, dtype=bool))
h.coupling_jobs(src.astype(np.int32), dst.astype(np.int32), _lib.make_opts(_lib.EST_CMI, 5), True); report("c3sub", h.stats())
E5 = recipes.zscore_rows(recipes.ar_panel(1000, 20000, 5)[:12])
h.upload_matrix(E5)
src, dst = np.nonzero(~np.eye(12, dtype=bool))
jobs = np.zeros(len(src), dtype=_lib.JOB_DTYPE); jobs["src"] = src; jobs["dst"] = dst; jobs["delay"] = 1
h.lagged_jobs(jobs, _lib.make_opts(_lib.EST_CMI, 5)); report("c5 cmi", h.stats())
