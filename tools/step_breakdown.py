#!/usr/bin/env python
"""Where one resident bench step spends its time: Python wall, the library's call span (CUDA events), estimator kernels."""
import os, sys, time
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import recipes
import scribe_py_b200 as S
import torch

G, T, DELAYS = 100, 2000, [1, 5, 10]
E = recipes.rdi_synthetic(G, T, 2)
frame = pd.DataFrame(E, index=pd.Index(["g%03d" % i for i in range(G)], name="GENE_ID"))
model = S.causal_model(frame)
plan = model._rdi_plan(DELAYS)
h = S._lib.default_handle()
for _ in range(3):
    model._rdi_execute(plan)
torch.cuda.synchronize()
rows = []
for _ in range(10):
    t0 = time.perf_counter()
    model._rdi_execute(plan)
    t1 = time.perf_counter()
    st = h.stats()
    rows.append(((t1 - t0) * 1e3, st["total_ms"], st["estimator_ms"]))
r = np.median(np.array(rows), axis=0)
print("python wall %.2f ms | library call span (events) %.2f ms | estimator kernels %.2f ms | host-only gap %.2f ms"
      % (r[0], r[1], r[2], r[0] - r[1]))
