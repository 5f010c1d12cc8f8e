#!/usr/bin/env python
"""Where the non-kernel time of one causal_net_dynamics_coupling(adata) call goes (C3 shape): host wall-clock per phase."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import recipes
import scribe_py_b200 as S
from scribe_py_b200 import Scribe as SC, _lib
G = int(sys.argv[1]) if len(sys.argv) > 1 else 500
spl, vel, genes = recipes.coupling_panel(G, 10000, 3)
h = _lib.default_handle()
sub = 40 if len(sys.argv) <= 2 else int(sys.argv[2])          # pairs evaluated: all sources x `sub` targets (keeps the kernel part short)
for rep in range(3):
    t = [time.perf_counter()]
    ad = recipes.FakeAnnData({"spliced": spl, "velocity": vel}, genes); t.append(time.perf_counter())
    plan = SC._coupling_plan(ad); t.append(time.perf_counter())
    a, b = SC._dense_layers(ad, plan["genes"], "spliced", "velocity"); t.append(time.perf_counter())
    ok = h.upload_coupling_raw(a, b, True, True); t.append(time.perf_counter())
    probe = [0, G // 2, G - 1]
    Sp, Yp, _ = SC._normalised_layers_pandas(ad, [plan["genes"][i] for i in probe], "spliced", "velocity", True); t.append(time.perf_counter())
    rows = [h.read_coupling_row(w, g, a.shape[0]) for g in probe for w in (0, 1)]; t.append(time.perf_counter())
    n = sub * (G - 1)
    opts = _lib.make_opts(_lib.EST_CMI, 5)
    vals = h.coupling_jobs(plan["src"][:n], plan["dst"][:n], opts, True); st = h.stats(); t.append(time.perf_counter())
    M = np.zeros((G, G)); full = np.zeros(len(plan["src"])); M[plan["aa"], plan["bb"]] = np.where(np.isnan(full), 0.0, full)
    import pandas as pd
    df = pd.DataFrame(M, index=plan["TFs"], columns=plan["Targets"]); t.append(time.perf_counter())
    names = ["adata", "plan", "dense_layers", "upload_raw+normalise", "pandas probe", "read 6 rows", "coupling_jobs(%d)" % n, "assemble frame"]
    d = np.diff(t) * 1e3
    print("rep %d: " % rep + " | ".join("%s %.1f" % (k, v) for k, v in zip(names, d)) + " | kernels %.1f ms, call span %.1f ms" % (st["estimator_ms"], st["total_ms"]))
