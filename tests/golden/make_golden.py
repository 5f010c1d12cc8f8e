"""Generates tests/golden/golden.npz by running the UNMODIFIED reference (/root/reference) on the seeded
inputs of recipes.py.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference files are loaded by path under a synthetic `Scribe` package with stub matplotlib modules
(SURVEY.md Appendix B); nothing is copied into this repo.  causal_network.py cannot finish rdi() on a modern
numpy (np.int at :299), so for the driver-level vectors a temporary copy is made under /tmp whose ONLY
change is `dtype=np.int` -> `dtype=object`; estimator-level vectors come from the untouched file.
Per-point eps/count vectors are produced with the same scipy calls the reference makes
(information_estimators.py:160-172), since the reference does not expose them.
"""
import importlib
import os
import shutil
import sys
import tempfile
import types
import warnings

import numpy as np
import pandas as pd
import scipy.spatial as ss

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import recipes  # noqa: E402

REF = "/root/reference/Scribe"


def load_reference():
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    tmp = tempfile.mkdtemp(prefix="scribe_ref_")
    for f in ("information_estimators.py", "Scribe.py"):
        os.symlink(os.path.join(REF, f), os.path.join(tmp, f))
    src = open(os.path.join(REF, "causal_network.py")).read()
    assert src.count("dtype=np.int") == 1
    open(os.path.join(tmp, "causal_network.py"), "w").write(src.replace("dtype=np.int", "dtype=object"))
    pkg = types.ModuleType("Scribe"); pkg.__path__ = [tmp]; sys.modules["Scribe"] = pkg
    ie = importlib.import_module("Scribe.information_estimators")
    cn = importlib.import_module("Scribe.causal_network")
    sc = importlib.import_module("Scribe.Scribe")
    return ie, cn, sc, tmp


def scipy_counts(x, y, z, k=5):
    """eps_i and the four ball sizes minus self with the reference's own scipy calls."""
    P = np.concatenate((x, y, z), axis=1); xz = np.concatenate((x, z), axis=1); yz = np.concatenate((y, z), axis=1)
    t = [ss.cKDTree(a) for a in (P, xz, yz, z)]
    eps = np.array([t[0].query(p, k + 1, p=np.inf)[0][k] for p in P])
    cnt = np.array([[len(t[s].query_ball_point(a[i], eps[i], p=np.inf)) - 1 for i in range(len(P))]
                    for s, a in enumerate((P, xz, yz, z))])
    return eps, cnt


def scipy_counts_mi(x, y, k=5):
    P = np.concatenate((x, y), axis=1)
    t = [ss.cKDTree(a) for a in (P, x, y)]
    eps = t[0].query(P, k + 1, p=np.inf)[0][:, k]
    cnt = np.array([[len(t[s].query_ball_point(a[i], eps[i], p=np.inf)) - 1 for i in range(len(P))]
                    for s, a in enumerate((P, x, y))])
    return eps, cnt


def main():
    warnings.simplefilter("ignore")
    ie, cn, sc, tmp = load_reference()
    out = {}
    L = lambda a: a.tolist()

    for name, fn in (("a", recipes.gv_a), ("b", recipes.gv_b), ("poisson", recipes.gv_poisson), ("d5", recipes.gv_d5),
                     ("dims", recipes.gv_dims)):
        x, y, z = fn()
        out[f"{name}_cmi"] = ie.cmi(L(x), L(y), L(z))
        out[f"{name}_cmi_norm"] = ie.cmi(L(x), L(y), L(z), normalization=True)
        out[f"{name}_cmi_k3"] = ie.cmi(L(x), L(y), L(z), k=3)
        out[f"{name}_cmi_k10"] = ie.cmi(L(x), L(y), L(z), k=10)
        out[f"{name}_mi"] = ie.mi(L(x), L(y))
        out[f"{name}_mi_k3"] = ie.mi(L(x), L(y), k=3)
        out[f"{name}_cumi_kde"] = ie.cumi(L(x), L(y), L(z))
        out[f"{name}_cumi_kde_bw"] = ie.cumi(L(x), L(y), L(z), bw=0.05, k=4)
        out[f"{name}_cumi_knn"] = ie.cumi(L(x), L(y), L(z), density_estimation_method="knn")
        out[f"{name}_umi_kde"] = ie.umi(L(x), L(y))
        try:
            out[f"{name}_umi_knn"] = ie.umi(L(x), L(y), density_estimation_method="knn")
        except ValueError:
            out[f"{name}_umi_knn"] = np.inf          # marker: the reference raised "math domain error"
        eps, cnt = scipy_counts(x, y, z)
        out[f"{name}_eps"] = eps; out[f"{name}_cnt"] = cnt
        eps, cnt = scipy_counts_mi(x, y)
        out[f"{name}_mi_eps"] = eps; out[f"{name}_mi_cnt"] = cnt

    for n in (4, 6, 7, 12, 40):       # tiny inputs incl. n <= k (cKDTree pads with inf)
        x, y, z = recipes.gv_small(n)
        out[f"small{n}_cmi"] = ie.cmi(L(x), L(y), L(z))
        out[f"small{n}_mi"] = ie.mi(L(x), L(y))

    # ---- C1: causal_net_dynamics_coupling through the real driver
    spl, vel, genes = recipes.gv_c()
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, genes)
    sc.causal_net_dynamics_coupling(ad)
    out["c_matrix"] = ad.uns["causal_net"]["RDI"].loc[genes, genes].to_numpy(dtype=float)
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, genes)
    sc.causal_net_dynamics_coupling(ad, TFs=genes[:5], Targets=genes[3:9], drop_zero_cells=False, normalize=False,
                                    t1_key="velocity")
    out["c_matrix_nodrop"] = ad.uns["causal_net"]["RDI"].to_numpy(dtype=float)
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "unspliced": np.abs(vel) * 3}, genes)
    sc.causal_net_dynamics_coupling(ad, TFs=genes[:4], Targets=genes[:6], t1_key="unspliced")
    out["c_matrix_t1"] = ad.uns["causal_net"]["RDI"].to_numpy(dtype=float)

    # ---- GV-D: single-run rdi (+ uniformised, + differential)
    X, g4, delays = recipes.gv_d()
    def model_single():
        m = cn.causal_model()
        df = pd.DataFrame(X, index=pd.Index(g4, name="GENE_ID"))
        m.expression = df; m.expression_raw = df; m.node_ids = df.index
        return m
    for tag, kw in (("", {}), ("_u", {"uniformization": True}), ("_diff", {"differential_mode": True})):
        m = model_single(); r = m.rdi(delays, **kw)
        for d in delays:
            out[f"d_rdi{tag}_{d}"] = r[d].to_numpy(dtype=float)
        out[f"d_rdi{tag}_MAX"] = r["MAX"].to_numpy(dtype=float)

    # ---- GV-E: multi-run rdi, MAX, crdi(L=1,2), uniformised crdi
    runs, g6, delays = recipes.gv_e()
    def model_multi():
        m = cn.causal_model()
        df = recipes.multi_run_frame(runs, g6)
        m.expression = df; m.expression_raw = df; m.node_ids = df.index.levels[0]; m.run_ids = df.index.levels[1]
        return m
    m = model_multi(); r = m.rdi(delays)
    for d in delays:
        out[f"e_rdi_{d}"] = r[d].to_numpy(dtype=float)
    out["e_rdi_MAX"] = r["MAX"].to_numpy(dtype=float)
    out["e_crdi_L1"] = m.crdi(L=1).to_numpy(dtype=float)
    out["e_crdi_L2"] = m.crdi(L=2).to_numpy(dtype=float)
    m = model_multi(); m.rdi([1, 2], uniformization=True)
    out["e_crdi_L1_u"] = m.crdi(L=1, uniformization=True).to_numpy(dtype=float)

    # ---- a slice of the C2 bench workload (single-run, 8 genes x 400 cells)
    E = recipes.rdi_synthetic(8, 400, 2)
    m = cn.causal_model(); df = pd.DataFrame(E, index=pd.Index(["g%d" % i for i in range(8)], name="GENE_ID"))
    m.expression = df; m.expression_raw = df; m.node_ids = df.index
    r = m.rdi([1, 5, 10])
    for d in (1, 5, 10):
        out[f"c2_rdi_{d}"] = r[d].to_numpy(dtype=float)

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **{k: np.asarray(v) for k, v in out.items()})
    shutil.rmtree(tmp)
    print("wrote", len(out), "arrays")
    for k in ("a_mi", "a_cmi", "a_umi_kde", "a_umi_knn", "a_cumi_kde", "a_cumi_knn", "a_cmi_norm", "b_mi", "b_cmi",
              "b_cumi_kde", "b_umi_kde", "b_cumi_knn", "b_umi_knn"):
        print(k, repr(float(out[k])))
    print("a sums", out["a_eps"].sum(), out["a_cnt"].sum(axis=1), " b sums", out["b_eps"].sum(), out["b_cnt"].sum(axis=1))
    print("c sum", out["c_matrix"].sum(), out["c_matrix"][0, 1])
    print("d", out["d_rdi_1"][0, 1], out["d_rdi_1"][1, 0], out["d_rdi_2"][0, 1])
    print("e", out["e_rdi_MAX"][0, 1], out["e_crdi_L2"][0, 1])


if __name__ == "__main__":
    main()
