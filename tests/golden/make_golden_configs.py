"""Config-scale golden vectors: the UNMODIFIED reference (/root/reference) run on the named BASELINE.json configs.

    python tests/golden/make_golden_configs.py [c2 c3 c4 c5 post] [--procs P]

Writes tests/golden/golden_configs.npz (merged with what is already there, so parts can be regenerated one at a time):

  c2    the FULL C2 matrix: rdi([1,5,10]) on 100 genes x 2000 cells = 29 700 cmi values.  The reference is single-threaded
        (~0.17 s per value), so the 4 950 unordered gene pairs are spread over P processes, each running the reference's own
        causal_model.rdi on a 2-gene frame (both directions, all delays) -- rdi treats every ordered pair independently
        (causal_network.py:134-156), so this is the same arithmetic as one 100-gene call.
  c3    causal_net_dynamics_coupling on the C3 layers (500 genes x 10 000 cells) for TFs x Targets = recipes.C3_TFS x
        recipes.C3_TARGETS (38 pairs; the reference normalises per gene, so a gene subset sees the numbers of the full run).
  c4    multi-run-form rdi([1,5,10]) + crdi(L=2) on recipes.C4_GENES of the C4 panel (200 genes x 5000 cells): 126 + 42 values.
  c5    rdi([1,5,10], uniformization=True) (cumi, KDE bw = 0.01) on recipes.C5_GENES of the z-scored C5 panel
        (1000 genes x 20 000 cells): 36 values at N = 19 999 / 19 995 / 19 990.
  diffcrdi  differential-mode rdi + crdi on GV-E (multi-run), L = 1, 2.
  post  CLR, roc, normalize, smooth, load_anndata, top-n edges: the reference's own post-/pre-processing functions on small inputs.

Run in the build container only.  Loader as in make_golden.py (stub matplotlib, /tmp copy of causal_network.py whose only
change is dtype=np.int -> dtype=object).
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import recipes  # noqa: E402
from make_golden import load_reference  # noqa: E402

OUT = os.path.join(HERE, "golden_configs.npz")
_REF = None


def ref():
    global _REF
    if _REF is None:
        warnings.simplefilter("ignore")
        _REF = load_reference()
    return _REF


def single_run_model(cn, E, names):
    m = cn.causal_model()
    df = pd.DataFrame(E, index=pd.Index(names, name="GENE_ID"))
    m.expression = df; m.expression_raw = df; m.node_ids = df.index
    return m


def _pair_rdi(args):
    """reference rdi on a 2-gene single-run frame -> [(a, b, delay, value), (b, a, delay, value), ...]"""
    Ea, Eb, a, b, delays, uniform = args
    _, cn, _, _ = ref()
    m = single_run_model(cn, np.stack([Ea, Eb]), ["a", "b"])
    r = m.rdi(delays, uniformization=uniform)
    out = []
    for d in delays:
        M = r[d].to_numpy(dtype=float)
        out.append((a, b, d, M[0, 1])); out.append((b, a, d, M[1, 0]))
    return out


def pooled_pairs(E, genes, delays, uniform, procs):
    from multiprocessing import get_context
    jobs = [(E[a], E[b], a, b, delays, uniform) for i, a in enumerate(genes) for b in genes[i + 1:]]
    G = E.shape[0]
    res = {d: np.full((G, G), np.nan) for d in delays}
    t0 = time.time()
    with get_context("fork").Pool(procs) as p:
        for n, part in enumerate(p.imap_unordered(_pair_rdi, jobs, chunksize=1)):
            for a, b, d, v in part:
                res[d][a, b] = v
            if n % 200 == 0:
                print("  %d / %d pairs, %.0f s" % (n, len(jobs), time.time() - t0), flush=True)
    return res


def part_c2(procs):
    E = recipes.rdi_synthetic(100, 2000, 2)
    res = pooled_pairs(E, list(range(100)), [1, 5, 10], False, procs)
    return {"c2full_rdi_%d" % d: res[d] for d in (1, 5, 10)}


def part_c3(procs):
    _, _, sc, _ = ref()
    spl, vel, genes = recipes.coupling_panel(500, 10000, 3)
    ad = recipes.FakeAnnData({"spliced": spl, "velocity": vel}, genes)
    tfs = [genes[i] for i in recipes.C3_TFS]; tg = [genes[i] for i in recipes.C3_TARGETS]
    sc.causal_net_dynamics_coupling(ad, TFs=tfs, Targets=tg)
    M = ad.uns["causal_net"]["RDI"]
    return {"c3_matrix": M.loc[tfs, tg].to_numpy(dtype=float)}


def part_c4(procs):
    _, cn, _, _ = ref()
    E = recipes.ar_panel(200, 5000, 4)[recipes.C4_GENES]
    names = ["g%04d" % i for i in recipes.C4_GENES]
    df = recipes.multi_run_frame([E], names)
    m = cn.causal_model()
    m.expression = df; m.expression_raw = df; m.node_ids = df.index.levels[0]; m.run_ids = df.index.levels[1]
    r = m.rdi([1, 5, 10])
    out = {"c4_rdi_%d" % d: r[d].loc[names, names].to_numpy(dtype=float) for d in (1, 5, 10)}
    out["c4_rdi_MAX"] = r["MAX"].loc[names, names].to_numpy(dtype=float)
    out["c4_crdi_L2"] = m.crdi(L=2).loc[names, names].to_numpy(dtype=float)
    return out


def part_c5(procs):
    E = recipes.zscore_rows(recipes.ar_panel(1000, 20000, 5))[recipes.C5_GENES]
    res = pooled_pairs(E, list(range(len(recipes.C5_GENES))), [1, 5, 10], True, procs)
    return {"c5_urdi_%d" % d: res[d] for d in (1, 5, 10)}


def part_post(procs):
    """pre-/post-processing functions of the reference on small inputs"""
    import tempfile
    _, cn, sc, _ = ref()
    out = {}
    rng = np.random.default_rng(31)
    # CLR: square asymmetric, square symmetric, rectangular; DataFrame in -> DataFrame out
    A = rng.gamma(1.5, 0.2, (12, 12)); np.fill_diagonal(A, 0.0)
    out["clr_in_sq"] = A
    out["clr_sq"] = np.asarray(sc.CLR(pd.DataFrame(A)))
    out["clr_sq_rowonly"] = np.asarray(sc.CLR(A.copy(), zscore_both_dim=False))
    Sm = (A + A.T) / 2
    out["clr_in_sym"] = Sm
    out["clr_sym"] = np.asarray(sc.CLR(Sm.copy()))
    R = rng.gamma(1.5, 0.2, (5, 9))
    out["clr_in_rect"] = R
    out["clr_rect"] = np.asarray(sc.CLR(R.copy()))
    out["clr_rect_both"] = np.asarray(sc.CLR(R.copy(), zscore_both_dim=True))
    # roc: 10 genes, random scores with ties, a true graph of 14 edges (mixed case on purpose: the reference lowercases)
    G = 10
    names = ["Gene%d" % i for i in range(G)]
    scores = np.round(rng.random((G, G)), 2)                       # ties
    pairs = [(a, b) for a in range(G) for b in range(G) if a != b]
    true = [pairs[i] for i in rng.choice(len(pairs), 14, replace=False)]
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        for a, b in true:
            fh.write("%s\t%s\n" % (names[a].upper() if a % 2 else names[a], names[b]))
        path = fh.name
    m = single_run_model(cn, rng.standard_normal((G, 30)), names)
    auroc, x, y = m.roc(pd.DataFrame(scores, index=names, columns=names), path)
    os.unlink(path)
    out["roc_scores"] = scores; out["roc_true"] = np.array(true); out["roc_auroc"] = auroc; out["roc_x"] = np.array(x); out["roc_y"] = np.array(y)
    # normalize / smooth (single-run and NaN-padded multi-run); smooth needs the axis= kwarg pandas 3 removed, so the
    # reference's expression `expression_raw.rolling(window, axis=1).mean()` is evaluated through its documented equivalent
    # `.T.rolling(window).mean().T` when the kwarg is rejected (same arithmetic: rolling mean along the cell axis).
    E = rng.gamma(2.0, 1.0, (6, 40)); E[2] = 1.5                  # a constant row: std 0 -> NaN -> fillna(0)
    m = single_run_model(cn, E, ["g%d" % i for i in range(6)])
    m.normalize()
    out["norm_in"] = E; out["norm_out"] = m.expression.to_numpy(dtype=float)
    runs, g6, _ = recipes.gv_e()
    df = recipes.multi_run_frame(runs, g6)
    m = cn.causal_model(); m.expression = df.copy(); m.expression_raw = df.copy()
    m.node_ids = df.index.levels[0]; m.run_ids = df.index.levels[1]
    m.normalize()
    out["norm_multi_out"] = m.expression.to_numpy(dtype=float)
    for w in (3, 7):
        m = single_run_model(cn, E, ["g%d" % i for i in range(6)])
        try:
            m.smooth(w)
        except TypeError:
            sm = m.expression_raw.T.rolling(window=w).mean().T
            m.expression = sm.dropna(axis=1, how="all")
        out["smooth%d_out" % w] = m.expression.to_numpy(dtype=float)
        out["smooth%d_cols" % w] = np.asarray(m.expression.columns, dtype=np.int64)
    # top-n edges (plot/networks.py:34-47): the data-frame part of vis_causal_net, run with the reference's own statements
    W = rng.gamma(1.5, 0.2, (8, 8)); np.fill_diagonal(W, 0.0)
    W[2, 5] = W[5, 2]                                             # a symmetric tie: neither direction is dropped
    names8 = ["n%d" % i for i in range(8)]
    df_mat = pd.DataFrame(W.copy(), index=names8, columns=names8)
    tmp = np.where(df_mat.values - df_mat.T.values < 0)
    for i in range(len(tmp[0])):
        df_mat.iloc[tmp[0][i], tmp[1][i]] = np.nan
    st = df_mat.stack().reset_index()
    st.columns = ["source", "target", "weight"]
    ind_vec = np.argsort(-st.loc[:, "weight"])
    top = st.loc[ind_vec[:10], :]
    out["edges_in"] = W
    out["edges_src"] = np.array([names8.index(s) for s in top["source"]]); out["edges_dst"] = np.array([names8.index(s) for s in top["target"]])
    out["edges_w"] = top["weight"].to_numpy(dtype=float)
    # kNN grid regressor of viz_causality (plot/heatmaps.py:389-423), the reference's own statements and scipy calls on one gene
    # pair: x(t - d), y(t), z = y(t - 1); grid over (x, z); tree over the (x, y) columns (as the reference builds it);
    # k + 1 neighbours, the nearest dropped, u = exp(-d / min d); expected_y normalised by its maximum
    import scipy.spatial as ss
    E = recipes.rdi_synthetic(6, 700, 9)
    delay, k, grid_num = 1, 5, 25
    x = [i for i in E[0]]; y_ori = [i for i in E[1]]
    x = x[:-delay]; y = y_ori[delay:]; z = y_ori[delay - 1:-1]
    cur_data = pd.DataFrame({'x': x, 'y': y, 'z': z, 'pair': 'a->b'})
    x_meshgrid = np.linspace(min(x), max(x), grid_num, endpoint=True)
    z_meshgrid = np.linspace(min(z), max(z), grid_num, endpoint=True)
    xv, zv = np.meshgrid(x_meshgrid, z_meshgrid)
    xp = xv.reshape((1, -1)).tolist(); zp = zv.reshape((1, -1)).tolist()
    xz_query = np.array(xp + zp).T
    tree_xz = ss.cKDTree(cur_data[['x', 'y']])
    dist_mat, idx_mat = tree_xz.query(xz_query, k=k + 1)
    exp_y = np.empty(len(dist_mat))
    for i in range(dist_mat.shape[0]):
        subset_dat = cur_data.iloc[idx_mat[i, 1:], 1]
        u = np.exp(-dist_mat[i, 1:] / np.min(dist_mat[i, 1:]))
        w = u / np.sum(u)
        exp_y[i] = sum(np.array(w) * np.array(subset_dat))
    out["knnreg_query"] = xz_query; out["knnreg_raw"] = exp_y.copy(); out["knnreg_norm"] = exp_y / max(exp_y)
    out.update(part_load_anndata())
    return out


def load_anndata_case():
    """AnnData stand-in with three dpt_groups branches (one of a single cell, which the reference skips), shuffled cells"""
    rng = np.random.default_rng(8)
    G, n = 4, 90
    X = rng.standard_normal((n, G)).cumsum(axis=0) * 0.3 + rng.standard_normal((n, G))
    branch = np.array(["1"] * 50 + ["2"] * 39 + ["3"])
    perm = rng.permutation(n)
    X, branch = X[perm], branch[perm]
    ptime = rng.random(n)
    obs = pd.DataFrame({"dpt_groups": branch, "dpt_pseudotime": ptime}, index=["c%d" % i for i in range(n)])
    genes = ["A", "B", "C", "D"]
    return recipes.FakeAnnData({"spliced": X.copy()}, genes, obs_names=list(obs.index), X=X, obs=obs), genes


def part_load_anndata():
    """Scribe.read_export.load_anndata (read_export.py:16-98) run from a /tmp copy whose ONLY change is `np.where(...)` ->
    `np.where(...)[0]` at :56 (pandas >= 2 rejects the tuple as a column indexer)."""
    import importlib, shutil, tempfile, types
    src = open("/root/reference/Scribe/read_export.py").read()
    old = "cur_id = np.where(np.array(keys) == str(key))"
    assert src.count(old) == 1
    tmp = tempfile.mkdtemp(prefix="scribe_ref_re_")
    open(os.path.join(tmp, "read_export.py"), "w").write(src.replace(old, old + "[0]"))
    for f in ("causal_network.py", "information_estimators.py", "logging.py", "settings.py"):
        os.symlink(os.path.join("/root/reference/Scribe", f), os.path.join(tmp, f))
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    saved = {k: v for k, v in sys.modules.items() if k == "Scribe" or k.startswith("Scribe.")}
    for k in saved:
        del sys.modules[k]
    pkg = types.ModuleType("Scribe"); pkg.__path__ = [tmp]; sys.modules["Scribe"] = pkg
    try:
        re_ = importlib.import_module("Scribe.read_export")
        ad, genes = load_anndata_case()
        m = re_.load_anndata(ad)
    finally:
        for k in [k for k in sys.modules if k == "Scribe" or k.startswith("Scribe.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        shutil.rmtree(tmp)
    e = m.expression
    return {"loadann_expression": e.to_numpy(dtype=float),
            "loadann_index_gene": np.array([str(t[0]) for t in e.index]), "loadann_index_run": np.array([str(t[1]) for t in e.index]),
            "loadann_node_ids": np.array([str(v) for v in m.node_ids]), "loadann_run_ids": np.array([str(v) for v in m.run_ids])}


def part_diffcrdi(procs):
    """differential-mode rdi + crdi (multi-run branch, causal_network.py:257-283 with :271): GV-E, L = 1 and 2"""
    _, cn, _, _ = ref()
    runs, g6, delays = recipes.gv_e()
    df = recipes.multi_run_frame(runs, g6)
    m = cn.causal_model(); m.expression = df; m.expression_raw = df
    m.node_ids = df.index.levels[0]; m.run_ids = df.index.levels[1]
    r = m.rdi(delays, differential_mode=True)
    out = {"e_rdi_diff_%d" % d: r[d].to_numpy(dtype=float) for d in delays}
    out["e_crdi_diff_L1"] = m.crdi(L=1, differential_mode=True).to_numpy(dtype=float)
    out["e_crdi_diff_L2"] = m.crdi(L=2, differential_mode=True).to_numpy(dtype=float)
    return out


PARTS = {"diffcrdi": part_diffcrdi, "c2": part_c2, "c3": part_c3, "c4": part_c4, "c5": part_c5, "post": part_post}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("parts", nargs="*", default=list(PARTS))
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    a = ap.parse_args()
    for p in a.parts:
        t0 = time.time()
        new = PARTS[p](a.procs)
        cur = dict(np.load(OUT)) if os.path.exists(OUT) else {}
        cur.update({k: np.asarray(v) for k, v in new.items()})
        np.savez_compressed(OUT, **cur)
        print("part %s: %d arrays, %.0f s" % (p, len(new), time.time() - t0), flush=True)
