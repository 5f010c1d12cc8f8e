"""GPU parity: the CUDA path (through the C-ABI) against the reference-generated fixtures and the oracle.
Bars: neighbour counts and eps bit-exact; mi/cmi values <= 1e-9 (budget 1e-5); umi/cumi <= 1e-6 (budget 1e-5:
the KDE exponent is evaluated by the FP32 MUFU); top-100 edge ranking identical."""
import numpy as np
import pandas as pd
import pytest

import recipes
from oracle import ksg_oracle as o

pytestmark = pytest.mark.gpu

TOL_EXACT = 1e-9     # estimators whose per-point inputs are integers (counts): only summation order differs
TOL_KDE = 1e-6       # uniformised estimators
SETS = {"a": recipes.gv_a, "b": recipes.gv_b, "poisson": recipes.gv_poisson, "d5": recipes.gv_d5, "dims": recipes.gv_dims}


@pytest.fixture(scope="module")
def S():
    import scribe_py_b200 as s
    s._lib.default_handle()          # raises if the CUDA library / device is missing
    return s


def close(got, ref, tol):
    if np.isnan(ref):
        return np.isnan(got)
    return abs(got - ref) <= tol


@pytest.mark.parametrize("path", ["auto", "generic", "dense", "sweep", "strip"])
@pytest.mark.parametrize("name", list(SETS))
def test_counts_bit_exact(S, golden, name, path):
    x, y, z = SETS[name]()
    d = S.information_estimators.ksg_debug("cmi", x, y, z, path=path)
    assert np.array_equal(d["eps"], golden[f"{name}_eps"])
    assert np.array_equal(d["counts"], golden[f"{name}_cnt"])
    d = S.information_estimators.ksg_debug("mi", x, y, path=path)
    assert np.array_equal(d["eps"], golden[f"{name}_mi_eps"])
    assert np.array_equal(d["counts"][:3], golden[f"{name}_mi_cnt"])


@pytest.mark.parametrize("name", list(SETS))
def test_estimator_values(S, golden, name):
    x, y, z = SETS[name]()
    L = lambda a: a.tolist()
    assert close(S.cmi(L(x), L(y), L(z)), float(golden[f"{name}_cmi"]), TOL_EXACT)
    assert close(S.cmi(x, y, z, normalization=True), float(golden[f"{name}_cmi_norm"]), TOL_EXACT)
    assert close(S.cmi(x, y, z, k=3), float(golden[f"{name}_cmi_k3"]), TOL_EXACT)
    assert close(S.cmi(x, y, z, k=10), float(golden[f"{name}_cmi_k10"]), TOL_EXACT)
    assert close(S.mi(L(x), L(y)), float(golden[f"{name}_mi"]), TOL_EXACT)
    assert close(S.mi(x, y, k=3), float(golden[f"{name}_mi_k3"]), TOL_EXACT)


@pytest.mark.parametrize("name", list(SETS))
def test_uniformised_values(S, golden, name):
    x, y, z = SETS[name]()
    assert close(S.cumi(x, y, z), float(golden[f"{name}_cumi_kde"]), TOL_KDE)
    assert close(S.cumi(x, y, z, bw=0.05, k=4), float(golden[f"{name}_cumi_kde_bw"]), TOL_KDE)
    assert close(S.cumi(x, y, z, density_estimation_method="knn"), float(golden[f"{name}_cumi_knn"]), TOL_KDE)
    assert close(S.umi(x, y), float(golden[f"{name}_umi_kde"]), TOL_KDE)
    ref = float(golden[f"{name}_umi_knn"])
    if np.isinf(ref):
        with pytest.raises(ValueError):
            S.umi(x, y, density_estimation_method="knn")
    else:
        assert close(S.umi(x, y, density_estimation_method="knn"), ref, TOL_KDE)


def test_uniformised_per_point(S):
    """weights and weighted ball sums against the oracle's float64 brute force."""
    x, y, z = recipes.gv_a()
    d = S.information_estimators.ksg_debug("cumi", x, y, z)
    P = np.concatenate((x, y, z), axis=1)
    w = o.kde_weights(P[:, [0, 2]], 0.01)
    eps, cnt, W = o.knn_counts(P, [(0, 1, 2), (0, 2), (1, 2), (2,)], 5, weights=w)
    assert np.array_equal(d["eps"], eps) and np.array_equal(d["counts"], cnt)
    assert np.allclose(d["weights"], w, rtol=2e-6, atol=0)
    assert np.allclose(d["wsum"][0], W[2], rtol=2e-6) and np.allclose(d["wsum"][1], W[3], rtol=2e-6)


@pytest.mark.parametrize("n", [4, 6, 7, 12, 40])
def test_tiny_inputs(S, golden, n):
    x, y, z = recipes.gv_small(n)
    assert close(S.cmi(x, y, z), float(golden[f"small{n}_cmi"]), TOL_EXACT)
    assert close(S.mi(x, y), float(golden[f"small{n}_mi"]), TOL_EXACT)


def test_error_behaviour(S):
    x, y, z = recipes.gv_small(12)
    with pytest.raises(AssertionError, match="Lists should have same length"):
        S.cmi(x, y[:-1], z)
    with pytest.raises(AssertionError, match="Lists should have same length"):
        S.mi(x[:-1], y)
    with pytest.raises(AssertionError, match="Set k smaller"):
        S.umi(x, y, k=12)
    with pytest.raises(ValueError, match="not recognized"):
        S.cumi(x, y, z, density_estimation_method="histogram")
    with pytest.raises(ValueError, match="not recognized"):
        S.umi(x, y, density_estimation_method="histogram")
    assert S.vd(3) == pytest.approx(o.vd(3), abs=1e-15)


@pytest.mark.parametrize("path", ["dense", "sweep", "strip"])
def test_random_sizes_against_oracle(S, path):
    """ragged sizes around the chunk / tile / CTA boundaries, continuous x,y and heavily tied z"""
    rng = np.random.default_rng(5)
    for n in (1, 2, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1025, 1537, 2049, 4100):
        x = rng.standard_normal((n, 1)); z = np.round(rng.standard_normal((n, 1)), 1); y = x + z + rng.standard_normal((n, 1))
        d = S.information_estimators.ksg_debug("cmi", x, y, z, path=path)
        eps, cnt = o.cmi_debug(x, y, z)
        assert np.array_equal(d["eps"], eps), n
        assert np.array_equal(d["counts"], cnt), n
        got, ref = S.cmi(x, y, z), o.cmi(x, y, z)
        assert close(got, ref, TOL_EXACT) or (np.isinf(ref) and got == ref), n


@pytest.mark.parametrize("name", ["a", "b", "poisson", "d5"])
def test_sweep_equals_dense_uniformised(S, name):
    """the sorted sweep must reproduce the dense kernels: eps/counts bit for bit, weights and weighted sums to 1e-12
    (the KDE sweep drops terms below 2^-64 of a sum that is >= 1)"""
    x, y, z = SETS[name]()
    a = S.information_estimators.ksg_debug("cumi", x, y, z, path="dense")
    for path in ("sweep", "strip"):
        b = S.information_estimators.ksg_debug("cumi", x, y, z, path=path)
        assert np.array_equal(a["eps"], b["eps"]) and np.array_equal(a["counts"], b["counts"]), path
        assert np.allclose(a["weights"], b["weights"], rtol=1e-6, atol=0)
        assert np.allclose(a["wsum"], b["wsum"], rtol=1e-6, atol=0)


def test_dynamics_coupling_c1(S, golden):
    spl, vel, genes = recipes.gv_c()
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, genes)
    assert S.causal_net_dynamics_coupling(ad) is None
    M = ad.uns["causal_net"]["RDI"]
    assert list(M.index) == genes and list(M.columns) == genes
    got = M.to_numpy(dtype=float)
    assert np.abs(got - golden["c_matrix"]).max() <= TOL_EXACT
    flat = lambda A: np.argsort(-A, axis=None, kind="stable")[:100]
    assert np.array_equal(flat(got), flat(golden["c_matrix"]))          # top-100 edge ranking identical
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, genes)
    out = S.causal_net_dynamics_coupling(ad, TFs=genes[:5], Targets=genes[3:9], drop_zero_cells=False, normalize=False,
                                         copy=True)
    assert out is ad
    assert np.abs(ad.uns["causal_net"]["RDI"].to_numpy(dtype=float) - golden["c_matrix_nodrop"]).max() <= TOL_EXACT
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "unspliced": np.abs(vel) * 3}, genes)
    S.causal_net_dynamics_coupling(ad, TFs=genes[:4], Targets=genes[:6], t1_key="unspliced")
    assert np.abs(ad.uns["causal_net"]["RDI"].to_numpy(dtype=float) - golden["c_matrix_t1"]).max() <= TOL_EXACT
    with pytest.raises(Exception):
        S.causal_net_dynamics_coupling(ad, TFs=["nope"])


def _model_single(S, X, genes):
    df = pd.DataFrame(X, index=pd.Index(genes, name="GENE_ID"))
    return S.causal_model(df)


def test_rdi_single_run(S, golden):
    X, genes, delays = recipes.gv_d()
    for tag, kw, tol in (("", {}, TOL_EXACT), ("_u", {"uniformization": True}, TOL_KDE),
                         ("_diff", {"differential_mode": True}, TOL_EXACT)):
        m = _model_single(S, X, genes)
        r = m.rdi(delays, **kw)
        assert set(r.keys()) == set(delays) | {"MAX"}
        for d in delays:
            got = r[d].to_numpy(dtype=float)
            assert np.allclose(got, golden[f"d_rdi{tag}_{d}"], atol=tol, rtol=0, equal_nan=True), (tag, d)
            assert np.isnan(np.diag(got)).all()
        assert np.allclose(r["MAX"].to_numpy(dtype=float), golden[f"d_rdi{tag}_MAX"], atol=tol, rtol=0)
        assert list(r[delays[0]].index) == genes


def test_rdi_crdi_multi_run(S, golden):
    runs, genes, delays = recipes.gv_e()
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    r = m.rdi(delays)
    for d in delays:
        assert np.allclose(r[d].to_numpy(dtype=float), golden[f"e_rdi_{d}"], atol=TOL_EXACT, rtol=0, equal_nan=True)
    assert np.allclose(r["MAX"].to_numpy(dtype=float), golden["e_rdi_MAX"], atol=TOL_EXACT, rtol=0)
    assert np.allclose(m.crdi(L=1).to_numpy(dtype=float), golden["e_crdi_L1"], atol=TOL_EXACT, rtol=0, equal_nan=True)
    assert np.allclose(m.crdi(L=2).to_numpy(dtype=float), golden["e_crdi_L2"], atol=TOL_EXACT, rtol=0, equal_nan=True)
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    m.rdi([1, 2], uniformization=True)
    assert np.allclose(m.crdi(L=1, uniformization=True).to_numpy(dtype=float), golden["e_crdi_L1_u"], atol=TOL_KDE,
                       rtol=0, equal_nan=True)


def test_crdi_runs_rdi_when_missing(S):
    runs, genes, _ = recipes.gv_e()
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    with pytest.warns(UserWarning, match="RUNNING RDI NOW"):
        c = m.crdi(L=1)
    assert 1 in m.rdi_results and c.shape == (6, 6)


def test_c2_slice(S, golden):
    E = recipes.rdi_synthetic(8, 400, 2)
    m = _model_single(S, E, ["g%d" % i for i in range(8)])
    r = m.rdi([1, 5, 10])
    for d in (1, 5, 10):
        assert np.allclose(r[d].to_numpy(dtype=float), golden[f"c2_rdi_{d}"], atol=TOL_EXACT, rtol=0, equal_nan=True)


def test_full_size_properties(S):
    """C2-sized job (N = 2000): size-independent checks that need no CPU oracle at full scale --
    n_xyz >= k with equality off ties, ball nesting n_xyz <= n_xz, n_yz <= n_z <= N-1, permutation invariance,
    and agreement of the fused and the generic kernels."""
    E = recipes.rdi_synthetic(3, 2000, 9)
    x, y, z = E[0, :-1, None], E[1, 1:, None], E[1, :-1, None]
    d = S.information_estimators.ksg_debug("cmi", x, y, z)
    c = d["counts"]
    assert (c[0] >= 5).all() and (c[0] == 5).mean() > 0.99
    assert (c[0] <= c[1]).all() and (c[0] <= c[2]).all() and (c[1] <= c[3]).all() and (c[2] <= c[3]).all()
    assert (c[3] <= len(x) - 1).all()
    for path in ("generic", "dense", "sweep", "strip"):
        g = S.information_estimators.ksg_debug("cmi", x, y, z, path=path)
        assert np.array_equal(g["eps"], d["eps"]) and np.array_equal(g["counts"], c), path
    perm = np.random.default_rng(0).permutation(len(x))
    v0, v1 = S.cmi(x, y, z), S.cmi(x[perm], y[perm], z[perm])
    assert abs(v0 - v1) < 1e-12
    # sampled rows against the oracle's definition
    P = np.concatenate((x, y, z), axis=1)
    for i in (0, 1, 777, 1998):
        dist = np.abs(P - P[i]).max(axis=1)
        assert d["eps"][i] == np.sort(dist)[5]
        assert c[3][i] == (np.abs(P[:, 2] - P[i, 2]) <= d["eps"][i]).sum() - 1


def test_load_anndata_rdi_on_gpu(S):
    """AnnData -> causal_model (branches as runs, pseudotime order) -> rdi on the GPU == oracle on the same ordering"""
    rng = np.random.default_rng(8)
    G, n = 4, 300
    X = rng.standard_normal((n, G)).cumsum(axis=0) * 0.3 + rng.standard_normal((n, G))
    branch = np.array(["1"] * 170 + ["2"] * 130)
    perm = rng.permutation(n)
    X, branch = X[perm], branch[perm]
    ptime = rng.random(n)
    obs = pd.DataFrame({"dpt_groups": branch, "dpt_pseudotime": ptime}, index=["c%d" % i for i in range(n)])
    ad = recipes.FakeAnnData({}, ["A", "B", "C", "D"], obs_names=list(obs.index), X=X, obs=obs)
    m = S.load_anndata(ad)
    r = m.rdi([1, 3])
    runs = []
    for key in ("1", "2"):
        c = np.where(branch == key)[0]; c = c[np.argsort(ptime[c], kind="stable")]
        runs.append(X[c].T)
    ref = o.rdi([[run[g] for run in runs] for g in range(G)], [1, 3])
    for d in (1, 3):
        assert np.allclose(r[d].to_numpy(dtype=float), ref[d], atol=TOL_EXACT, rtol=0, equal_nan=True)


def test_stats_and_paths(S):
    """per-pass visit counters: the dense kernels visit N^2 pairs per pass, the sweep strictly fewer, same answer"""
    h = S._lib.default_handle()
    E = recipes.rdi_synthetic(3, 1500, 4)
    x, y, z = E[0, :-1, None], E[1, 1:, None], E[1, :-1, None]
    n = len(x)
    vals = {}
    for name, path in (("dense", S._lib.PATH_DENSE), ("sweep", S._lib.PATH_SWEEP), ("strip", S._lib.PATH_STRIP)):
        vals[name] = h.estimate(x, y, z, S._lib.make_opts(S._lib.EST_CMI, 5, path=path))
        st = h.stats()
        assert st["jobs"] == 1 and st["kernel_launches"] >= 2 and st["estimator_ms"] > 0
        if name == "dense":
            assert st["visited_knn"] == n * n and st["visited_count"] == n * n
        elif name == "sweep":
            assert 0 < st["visited_knn"] < n * n and 0 < st["visited_count"] < n * n
        else:                                   # strip kernels: every count is a range query, only candidates are tallied
            assert 0 < st["visited_knn"] < n * n
    assert abs(vals["dense"] - vals["sweep"]) < 1e-13 and abs(vals["dense"] - vals["strip"]) < 1e-13


@pytest.mark.parametrize("dz", [1, 2, 3])
@pytest.mark.parametrize("k", [3, 5, 10])
def test_sweep_matrix_against_oracle(S, dz, k):
    """sorted sweep across conditioning dimension and k (KCAP 6 / 16), continuous + tied key column, cmi and cumi"""
    rng = np.random.default_rng(100 * dz + k)
    n = 700 + 37 * dz
    z = rng.standard_normal((n, dz))
    z[:, -1] = np.round(z[:, -1], 1)                       # heavy ties in the sort key
    x = 0.5 * z[:, :1] + rng.standard_normal((n, 1))
    y = x + z[:, -1:] + 0.5 * rng.standard_normal((n, 1))
    P = np.concatenate((x, y, z), axis=1)
    ix, iy, iz = (0,), (1,), tuple(range(2, 2 + dz))
    eps, cnt = o.knn_counts(P, [ix + iy + iz, ix + iz, iy + iz, iz], k)
    for path in ("sweep", "dense", "strip"):
        d = S.information_estimators.ksg_debug("cmi", x, y, z, k=k, path=path)
        assert np.array_equal(d["eps"], eps), path
        assert np.array_equal(d["counts"], cnt), path
    h = S._lib.default_handle()
    v_sweep = h.estimate(x, y, z, S._lib.make_opts(S._lib.EST_CUMI, k, bw=0.2, path=S._lib.PATH_SWEEP))
    v_dense = h.estimate(x, y, z, S._lib.make_opts(S._lib.EST_CUMI, k, bw=0.2, path=S._lib.PATH_DENSE))
    v_strip = h.estimate(x, y, z, S._lib.make_opts(S._lib.EST_CUMI, k, bw=0.2, path=S._lib.PATH_STRIP))
    ref = o.cumi(x, y, z, k=k, bw=0.2)
    assert abs(v_sweep - ref) < TOL_KDE and abs(v_dense - ref) < TOL_KDE and abs(v_sweep - v_dense) < 1e-7
    assert abs(v_strip - v_sweep) < 1e-12


def test_sweep_zero_inflated_large(S):
    """C3-like zero-inflated columns at N = 6000 (mask keeps ~45 %): sweep == dense == generic, per-point"""
    rng = np.random.default_rng(12)
    n = 6000
    x = rng.gamma(2., 1., (n, 1)); x[rng.random((n, 1)) < 0.3] = 0
    z = rng.gamma(2., 1., (n, 1)); z[rng.random((n, 1)) < 0.3] = 0
    y = z + 0.1 * rng.standard_normal((n, 1)); y[rng.random((n, 1)) < 0.1] = 0
    a = S.information_estimators.ksg_debug("cmi", x, y, z, path="dense")
    for path in ("sweep", "strip"):
        b = S.information_estimators.ksg_debug("cmi", x, y, z, path=path)
        assert np.array_equal(a["eps"], b["eps"]) and np.array_equal(a["counts"], b["counts"]), path
    assert (a["eps"] == 0).mean() > 0.002                     # the tie case (eps = 0) is actually exercised
    P = np.concatenate((x, y, z), axis=1)
    for i in (0, 17, 2999, 5999):
        dist = np.abs(P - P[i]).max(axis=1)
        assert a["eps"][i] == np.sort(dist)[5]
        assert a["counts"][0][i] == (dist <= a["eps"][i]).sum() - 1


def test_mi_network(S):
    """other_estimators.mi: all ordered pairs in one submission == oracle mi per pair (symmetric by construction)"""
    E = recipes.rdi_synthetic(5, 400, 6)
    m = S.causal_model(pd.DataFrame(E, index=pd.Index(["g%d" % i for i in range(5)], name="GENE_ID")))
    M = S.other_estimators.mi(m).to_numpy(dtype=float)
    assert np.isnan(np.diag(M)).all()
    for a in range(5):
        for b in range(5):
            if a != b:
                assert abs(M[a, b] - o.mi(E[a], E[b])) < TOL_EXACT
    assert np.allclose(M, M.T, equal_nan=True, atol=1e-12)
    C = S.other_estimators.corr(m).to_numpy(dtype=float)
    assert abs(C[0, 1] - np.corrcoef(E[0], E[1])[0, 1]) < 1e-12


def test_rdi_network_top100_ranking(S):
    """a 12-gene RDI network (132 ordered pairs x 2 delays): values within tolerance and the top-100 edge ranking of
    the MAX matrix identical to the oracle's"""
    E = recipes.rdi_synthetic(12, 500, 11)
    genes = ["g%02d" % i for i in range(12)]
    r = S.causal_model(pd.DataFrame(E, index=pd.Index(genes, name="GENE_ID"))).rdi([1, 5])
    ref = o.rdi([[E[g]] for g in range(12)], [1, 5])
    for d in (1, 5):
        assert np.allclose(r[d].to_numpy(dtype=float), ref[d], atol=TOL_EXACT, rtol=0, equal_nan=True)
    got, want = r["MAX"].to_numpy(dtype=float), ref["MAX"]
    top = lambda A: np.argsort(-A, axis=None, kind="stable")[:100]
    assert np.array_equal(top(got), top(want))


def test_full_size_cumi_job(S):
    """one C5-sized job (N = 19 999, z-scored, KDE bw = 0.01) through the path AUTO picks at that size (strip kernels):
    eps and all four counts bit-identical to the float64 brute force, weights to 2e-6, the estimate to 1e-6"""
    E = recipes.rdi_synthetic(3, 20000, 5)
    E = (E - E.mean(axis=1, keepdims=True)) / E.std(axis=1, keepdims=True)
    x, y, z = E[0, :-1, None], E[1, 1:, None], E[1, :-1, None]
    P = np.concatenate((x, y, z), axis=1)
    w = o.kde_weights(P[:, [0, 2]], 0.01)
    eps, cnt, W = o.knn_counts(P, [(0, 1, 2), (0, 2), (1, 2), (2,)], 5, weights=w)
    d = S.information_estimators.ksg_debug("cumi", x, y, z)
    assert np.array_equal(d["eps"], eps) and np.array_equal(d["counts"], cnt)
    assert np.allclose(d["weights"], w, rtol=2e-6, atol=0)
    assert np.allclose(d["wsum"][0], W[2], rtol=2e-6) and np.allclose(d["wsum"][1], W[3], rtol=2e-6)
    from scipy.special import digamma
    ref = float(np.mean(w * digamma(cnt[0]) - w * digamma(cnt[1]) - w * digamma(W[2]) + w * digamma(W[3])))
    assert abs(float(S.cumi(x, y, z)) - ref) < TOL_KDE
    c = S.information_estimators.ksg_debug("cmi", x, y, z)              # the sweep at full size
    assert np.array_equal(c["eps"], eps) and np.array_equal(c["counts"], cnt)


def test_rdi_short_run_contributes_nothing(S):
    """a run not longer than the delay contributes no samples (e[:-delay] is empty, causal_network.py:167-171)"""
    rng = np.random.default_rng(3)
    G = 3
    runs = [rng.standard_normal((G, 400)).cumsum(axis=1) * 0.2 + rng.standard_normal((G, 400)), rng.standard_normal((G, 4))]
    genes = ["a", "b", "c"]
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    r = m.rdi([1, 5])
    ref = o.rdi([[run[g] for run in runs] for g in range(G)], [1, 5])
    only_long = o.rdi([[runs[0][g]] for g in range(G)], [5])
    for d in (1, 5):
        assert np.allclose(r[d].to_numpy(dtype=float), ref[d], atol=TOL_EXACT, rtol=0, equal_nan=True)
    assert np.allclose(r[5].to_numpy(dtype=float), only_long[5], atol=TOL_EXACT, rtol=0, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("dz", [1, 2, 3])
@pytest.mark.parametrize("scale", [1.0, 1e-42, 1e-30, 1e20, 1e35, 3e300])
def test_windowed_sweep_extreme_scales(S, dz, scale):
    """the FP32 pre-filter of the windowed sweep must stay a superset at any data scale: FP32-denormal (1e-42), FP32-overflow
    (1e35: absmax > 1e30 marks everything; 3e300) and in between -- eps and counts bit-identical to the dense FP64 kernel"""
    rng = np.random.default_rng(900 + dz)
    n = 1100
    z = rng.standard_normal((n, dz))
    x = 0.4 * z[:, :1] + rng.standard_normal((n, 1))
    y = x + z[:, -1:] + 0.3 * rng.standard_normal((n, 1))
    x, y, z = x * scale, y * scale, z * scale
    a = S.information_estimators.ksg_debug("cmi", x, y, z, path="dense")
    b = S.information_estimators.ksg_debug("cmi", x, y, z, path="sweep")
    assert np.array_equal(a["eps"], b["eps"]) and np.array_equal(a["counts"], b["counts"])
    assert np.isfinite(a["eps"]).all() and (a["counts"][0] >= 5).all()
    # and the dense kernel itself against the oracle at this scale (the comparison above is otherwise self-referential)
    P = np.concatenate((x, y, z), axis=1)
    iz = tuple(range(2, 2 + dz))
    eps, cnt = o.knn_counts(P, [(0, 1) + iz, (0,) + iz, (1,) + iz, iz], 5)
    assert np.array_equal(a["eps"], eps) and np.array_equal(a["counts"], cnt)


@pytest.mark.gpu
def test_windowed_sweep_fp32_collisions(S):
    """coordinates that differ only below FP32 resolution (1 + j*2^-40) and an offset cluster (1e6 + small): the FP32 images
    collide, the exact FP64 re-test must still order them; plus mixed magnitudes in one column"""
    rng = np.random.default_rng(77)
    n = 900
    x = 1.0 + rng.integers(0, 4000, (n, 1)) * 2.0 ** -40
    z = 1e6 + rng.standard_normal((n, 1)) * 1e-3
    y = z - 1e6 + x + rng.standard_normal((n, 1)) * 1e-9
    x[::7] *= 1e5                                              # mixed magnitudes
    a = S.information_estimators.ksg_debug("cmi", x, y, z, path="dense")
    b = S.information_estimators.ksg_debug("cmi", x, y, z, path="sweep")
    assert np.array_equal(a["eps"], b["eps"]) and np.array_equal(a["counts"], b["counts"])
    P = np.concatenate((x, y, z), axis=1)
    eps, cnt = o.knn_counts(P, [(0, 1, 2), (0, 2), (1, 2), (2,)], 5)
    assert np.array_equal(b["eps"], eps) and np.array_equal(b["counts"], cnt)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [257, 1000, 4097])
def test_windowed_sweep_vs_first_sweep_poisson(S, n):
    """integer-count data (eps = 0 for most points, huge windows of tied keys): windowed sweep == dense == oracle"""
    rng = np.random.default_rng(n)
    x = rng.poisson(1.0, (n, 1)).astype(float)
    z = rng.poisson(2.0, (n, 1)).astype(float)
    y = (z + rng.poisson(0.5, (n, 1))).astype(float)
    a = S.information_estimators.ksg_debug("cmi", x, y, z, path="dense")
    b = S.information_estimators.ksg_debug("cmi", x, y, z, path="sweep")
    assert np.array_equal(a["eps"], b["eps"]) and np.array_equal(a["counts"], b["counts"])
    if n <= 1000:
        P = np.concatenate((x, y, z), axis=1)
        eps, cnt = o.knn_counts(P, [(0, 1, 2), (0, 2), (1, 2), (2,)], 5)
        assert np.array_equal(b["eps"], eps) and np.array_equal(b["counts"], cnt)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [3, 5, 10])
@pytest.mark.parametrize("kind", ["continuous", "ties", "poisson"])
def test_mi_sweep_against_oracle(S, k, kind):
    """mi on the windowed sweep (sorted by y, x marginal by rank range): eps and the three counts bit-identical to the
    oracle and to the dense kernel, value equal to the dense kernel's"""
    rng = np.random.default_rng(31 * k + len(kind))
    n = 1237
    if kind == "poisson":
        x = rng.poisson(2.0, (n, 1)).astype(float); y = (x + rng.poisson(1.0, (n, 1))).astype(float)
    else:
        x = rng.standard_normal((n, 1)); y = 0.6 * x + rng.standard_normal((n, 1))
        if kind == "ties":
            x = np.round(x, 1); y = np.round(y, 1)
    P = np.concatenate((x, y), axis=1)
    eps, cnt = o.knn_counts(P, [(0, 1), (0,), (1,)], k)
    a = S.information_estimators.ksg_debug("mi", x, y, k=k, path="dense")
    b = S.information_estimators.ksg_debug("mi", x, y, k=k, path="sweep")
    assert np.array_equal(a["eps"], eps) and np.array_equal(b["eps"], eps)
    assert np.array_equal(np.asarray(a["counts"])[:3], cnt) and np.array_equal(np.asarray(b["counts"])[:3], cnt)
    h = S._lib.default_handle()
    xs, ys = np.ascontiguousarray(x), np.ascontiguousarray(y)
    v_d = h.estimate(xs, ys, None, S._lib.make_opts(S._lib.EST_MI, k, path=S._lib.PATH_DENSE))
    v_s = h.estimate(xs, ys, None, S._lib.make_opts(S._lib.EST_MI, k, path=S._lib.PATH_SWEEP))
    assert abs(v_d - v_s) < 1e-12 and close(v_s, o.mi(x, y, k=k), TOL_EXACT)


@pytest.mark.gpu
def test_mi_network_sweep_equals_dense(S):
    """scribe_mi_jobs through the sweep (per-gene argsort + sorted rows cached per upload) == dense kernel, every pair"""
    rng = np.random.default_rng(5)
    G, T = 9, 700
    E = rng.standard_normal((G, T)); E[3] = 0.7 * E[1] + 0.3 * rng.standard_normal(T); E[5] = np.round(E[5], 1)
    h = S._lib.default_handle()
    src, dst = np.nonzero(~np.eye(G, dtype=bool))
    src, dst = src.astype(np.int32), dst.astype(np.int32)
    h.upload_matrix(E)
    v_s = h.mi_jobs(src, dst, S._lib.make_opts(S._lib.EST_MI, 5, path=S._lib.PATH_SWEEP))
    v_d = h.mi_jobs(src, dst, S._lib.make_opts(S._lib.EST_MI, 5, path=S._lib.PATH_DENSE))
    assert np.abs(v_s - v_d).max() < 1e-12
    ref = np.array([o.mi(E[a][:, None], E[b][:, None]) for a, b in zip(src[:12], dst[:12])])
    assert np.abs(v_s[:12] - ref).max() < TOL_EXACT
    h.upload_matrix(E[::-1].copy())                       # a new upload must invalidate the cached order
    v2 = h.mi_jobs(src, dst, S._lib.make_opts(S._lib.EST_MI, 5, path=S._lib.PATH_SWEEP))
    v2d = h.mi_jobs(src, dst, S._lib.make_opts(S._lib.EST_MI, 5, path=S._lib.PATH_DENSE))
    assert np.abs(v2 - v2d).max() < 1e-12 and np.abs(v2 - v_s).max() > 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("est", ["cmi", "cumi"])
def test_rdi_fast_path_equals_generic_job_loop(S, est):
    """scribe_lagged_jobs: the device-built descriptors of the plain-RDI fast path (full pair x delay table) give the same
    values, bit for bit, as the generic host job loop (taken for a table with fewer jobs than gene x delay columns, for a
    table that mixes in a conditioned job, and for multi-run input)"""
    lib = S._lib
    for runs in (None, [0, 300, 520]):
        rng = np.random.default_rng(3)
        G, T = 12, 520
        E = np.cumsum(rng.standard_normal((G, T)), axis=1) * 0.1 + rng.standard_normal((G, T))
        h = lib.default_handle()
        h.upload_matrix(E, None if runs is None else np.array(runs, dtype=np.int64))
        src, dst = np.nonzero(~np.eye(G, dtype=bool))
        delays = [1, 3, 7]
        jobs = np.zeros(len(src) * len(delays), dtype=lib.JOB_DTYPE)
        for i, d in enumerate(delays):
            sl = slice(i * len(src), (i + 1) * len(src))
            jobs["src"][sl] = src; jobs["dst"][sl] = dst; jobs["delay"][sl] = d
        opts = lib.make_opts(lib.EST_CUMI if est == "cumi" else lib.EST_CMI, 5, bw=0.2)
        fast = h.lagged_jobs(jobs, opts)                                  # 396 jobs >= 12 genes x 3 delays: fast path
        pick = np.array([0, 5, 131, 132, 200, 263, 264, 300, 395])
        slow = h.lagged_jobs(jobs[pick], opts)                            # 9 jobs < 36 columns: generic loop
        assert np.array_equal(fast[pick], slow)
        mixed = jobs[pick].copy()
        mixed = np.concatenate((mixed, mixed[:1]))
        mixed["n_cond"][-1] = 1; mixed["cond_gene"][-1, 0] = 2; mixed["cond_delay"][-1, 0] = 2
        got = h.lagged_jobs(mixed, opts)                                  # a conditioned job in the table: generic loop
        assert np.array_equal(got[:-1], slow)
        x, y, z = o.lagged_columns([E[0]] if runs is None else [E[0, a:b] for a, b in zip(runs[:-1], runs[1:])],
                                   [E[1]] if runs is None else [E[1, a:b] for a, b in zip(runs[:-1], runs[1:])], 1, False)
        ref = o.cumi(x, y, z, bw=0.2) if est == "cumi" else o.cmi(x, y, z)
        j01 = int(np.nonzero((jobs["src"] == 0) & (jobs["dst"] == 1) & (jobs["delay"] == 1))[0][0])
        assert close(fast[j01], ref, TOL_KDE if est == "cumi" else TOL_EXACT)


@pytest.mark.gpu
def test_mi_sweep_ragged_sizes(S):
    """mi on the windowed sweep for sizes around the chunk boundaries and n <= k (eps = inf): == dense kernel == oracle"""
    rng = np.random.default_rng(8)
    for n in (1, 2, 5, 6, 7, 31, 32, 33, 64, 65, 255, 257, 513):
        x = np.round(rng.standard_normal((n, 1)), 1); y = x + rng.standard_normal((n, 1))
        a = S.information_estimators.ksg_debug("mi", x, y, path="dense")
        b = S.information_estimators.ksg_debug("mi", x, y, path="sweep")
        assert np.array_equal(a["eps"], b["eps"]), n
        assert np.array_equal(np.asarray(a["counts"])[:3], np.asarray(b["counts"])[:3]), n
        eps, cnt = o.knn_counts(np.concatenate((x, y), axis=1), [(0, 1), (0,), (1,)], 5)
        assert np.array_equal(b["eps"], eps) and np.array_equal(np.asarray(b["counts"])[:3], cnt), n


# ---------------------------------------------------------------------------------------------------------------
# config-scale parity: the named BASELINE configs at their full sizes against outputs of the UNMODIFIED reference
# (tests/golden/make_golden_configs.py -> golden_configs.npz)
# ---------------------------------------------------------------------------------------------------------------
def _top_edges(M, n):
    """the n strongest ordered pairs of a score matrix (diagonal excluded), strongest first; ties broken by (source, target)"""
    G = M.shape[0]
    a, b = np.nonzero(~np.eye(G, dtype=bool))
    order = np.lexsort((b, a, -M[a, b]))
    return list(zip(a[order[:n]].tolist(), b[order[:n]].tolist()))


def test_config_c2_full_matrix(S, golden_configs):
    """BASELINE configs[1] in full: all 29 700 values of rdi([1,5,10]) on 100 genes x 2000 cells, and the top-100 edge ranking
    of the MAX matrix (9 900 edges)"""
    E = recipes.rdi_synthetic(100, 2000, 2)
    m = S.causal_model(pd.DataFrame(E, index=pd.Index(["g%03d" % i for i in range(100)], name="GENE_ID")))
    r = m.rdi([1, 5, 10])
    ref_max = np.full((100, 100), -np.inf)
    for d in (1, 5, 10):
        got, ref = r[d].to_numpy(dtype=float), golden_configs["c2full_rdi_%d" % d]
        assert np.isnan(np.diag(got)).all()
        assert np.nanmax(np.abs(got - ref)) <= TOL_EXACT, d
        ref_max = np.fmax(ref_max, ref)
    got_max = r["MAX"].to_numpy(dtype=float)
    off = ~np.eye(100, dtype=bool)
    assert np.abs(got_max[off] - ref_max[off]).max() <= TOL_EXACT
    assert _top_edges(got_max, 100) == _top_edges(ref_max, 100)


def test_config_c3_sampled_pairs(S, golden_configs):
    """BASELINE configs[2]: 38 (TF, target) pairs of the 500 genes x 10 000 cells dynamics-coupling panel through the public
    driver (device normalisation of the full-width layers included)"""
    spl, vel, genes = recipes.coupling_panel(500, 10000, 3)
    ad = recipes.FakeAnnData({"spliced": spl, "velocity": vel}, genes)
    tfs = [genes[i] for i in recipes.C3_TFS]; tg = [genes[i] for i in recipes.C3_TARGETS]
    S.causal_net_dynamics_coupling(ad, TFs=tfs, Targets=tg)
    got = ad.uns["causal_net"]["RDI"].loc[tfs, tg].to_numpy(dtype=float)
    assert np.abs(got - golden_configs["c3_matrix"]).max() <= TOL_EXACT
    assert (got[[1, 3], [1, 4]] == 0).all()                     # g0077 / g0251 are TF and target: skipped pairs read 0


def test_config_c4_rdi_crdi(S, golden_configs):
    """BASELINE configs[3]: multi-run-form rdi([1,5,10]) + crdi(L=2) (D = 5) on 7 genes x 5000 cells of the C4 panel"""
    E = recipes.ar_panel(200, 5000, 4)[recipes.C4_GENES]
    names = ["g%04d" % i for i in recipes.C4_GENES]
    m = S.causal_model(recipes.multi_run_frame([E], names))
    r = m.rdi([1, 5, 10])
    for d in (1, 5, 10):
        assert np.nanmax(np.abs(r[d].loc[names, names].to_numpy(dtype=float) - golden_configs["c4_rdi_%d" % d])) <= TOL_EXACT
    assert np.array_equal(np.isneginf(r["MAX"].to_numpy(dtype=float)), np.isneginf(golden_configs["c4_rdi_MAX"]))
    c = m.crdi(L=2).loc[names, names].to_numpy(dtype=float)
    assert np.nanmax(np.abs(c - golden_configs["c4_crdi_L2"])) <= TOL_EXACT


def test_config_c5_uniformised(S, golden_configs):
    """BASELINE configs[4]: 36 uniformised-RDI (cumi, KDE bw = 0.01) values at N = 19 999 / 19 995 / 19 990"""
    E = recipes.zscore_rows(recipes.ar_panel(1000, 20000, 5))[recipes.C5_GENES]
    m = S.causal_model(pd.DataFrame(E, index=pd.Index(["g%d" % i for i in range(len(E))], name="GENE_ID")))
    r = m.rdi([1, 5, 10], uniformization=True)
    for d in (1, 5, 10):
        assert np.nanmax(np.abs(r[d].to_numpy(dtype=float) - golden_configs["c5_urdi_%d" % d])) <= TOL_KDE, d


def test_crdi_differential_mode_gpu(S, golden_configs):
    """differential-mode rdi + crdi (multi-run branch of the reference, causal_network.py:257-283 with :271)"""
    runs, genes, delays = recipes.gv_e()
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    r = m.rdi(delays, differential_mode=True)
    for d in delays:
        assert np.allclose(r[d].to_numpy(dtype=float), golden_configs["e_rdi_diff_%d" % d], atol=TOL_EXACT, rtol=0, equal_nan=True)
    for L in (1, 2):
        c = m.crdi(L=L, differential_mode=True).to_numpy(dtype=float)
        assert np.allclose(c, golden_configs["e_crdi_diff_L%d" % L], atol=TOL_EXACT, rtol=0, equal_nan=True), L


def test_device_layer_normalisation_bit_identical(S):
    """scribe_upload_coupling_raw: the device-normalised S / Y rows are the same float64 as the reference's pandas expressions
    (Scribe.py:101-106,:121) for every gene; a NaN in the t0 layer is reported so the driver takes the host path"""
    from scribe_py_b200 import Scribe as SC
    rng = np.random.default_rng(5)
    for n_cells, n_genes in ((500, 20), (1037, 7), (130, 3), (10000, 12)):
        spl = rng.gamma(2., 1., (n_cells, n_genes)); spl[rng.random((n_cells, n_genes)) < 0.2] = 0
        vel = 0.1 * rng.standard_normal((n_cells, n_genes)); vel[rng.random((n_cells, n_genes)) < 0.01] = np.nan
        genes = ["g%02d" % i for i in range(n_genes)]
        ad = recipes.FakeAnnData({"spliced": spl, "velocity": vel}, genes)
        h = S._lib.default_handle()
        for normalize, t1 in ((True, "velocity"), (False, "velocity"), (True, "other")):
            ad.layers["other"] = np.abs(vel)
            Sp, Yp, _ = SC._normalised_layers_pandas(ad, genes, "spliced", t1, normalize)
            assert h.upload_coupling_raw(spl, ad.layers[t1], normalize, t1 == "velocity")
            for g in range(n_genes):
                assert np.array_equal(h.read_coupling_row(0, g, n_cells), Sp[g], equal_nan=True)
                assert np.array_equal(h.read_coupling_row(1, g, n_cells), Yp[g], equal_nan=True)
    bad = spl.copy(); bad[3, 1] = np.nan
    assert not h.upload_coupling_raw(bad, np.nan_to_num(vel), True, True)


def test_post_processing_on_device(S, golden_configs, tmp_path):
    """CLR (incl. the square-matrix broadcast quirk), roc and top-n edge extraction on the GPU against outputs of the reference's
    own functions (Scribe.py:142-176, causal_network.py:337-389, plot/networks.py:34-47)"""
    g = golden_configs
    assert np.allclose(S.CLR(pd.DataFrame(g["clr_in_sq"])).to_numpy(), g["clr_sq"], atol=1e-12, rtol=0)
    assert np.allclose(S.CLR(g["clr_in_sq"], zscore_both_dim=False), g["clr_sq_rowonly"], atol=1e-12, rtol=0)
    assert np.allclose(S.CLR(g["clr_in_sym"]), g["clr_sym"], atol=1e-12, rtol=0)
    assert np.allclose(S.CLR(g["clr_in_rect"]), g["clr_rect"], atol=1e-12, rtol=0)
    assert np.allclose(S.CLR(g["clr_in_rect"], zscore_both_dim=True), g["clr_rect_both"], atol=1e-12, rtol=0)
    names = ["Gene%d" % i for i in range(10)]
    path = tmp_path / "true.txt"
    path.write_text("".join("%s\t%s\n" % (names[a].upper() if a % 2 else names[a], names[b]) for a, b in g["roc_true"]))
    m = S.causal_model(pd.DataFrame(np.zeros((10, 5)), index=pd.Index(names, name="GENE_ID")))
    auroc, x, y = m.roc(pd.DataFrame(g["roc_scores"], index=names, columns=names), str(path))
    assert auroc == pytest.approx(float(g["roc_auroc"]), abs=1e-15)
    assert np.array_equal(x, g["roc_x"]) and np.array_equal(y, g["roc_y"])
    with pytest.raises(ZeroDivisionError):
        (tmp_path / "none.txt").write_text("nobody\tnothing\n")
        m.roc(pd.DataFrame(g["roc_scores"], index=names, columns=names), str(tmp_path / "none.txt"))
    n8 = ["n%d" % i for i in range(8)]
    top = S.top_n_edges(pd.DataFrame(g["edges_in"], index=n8, columns=n8), top_n_edges=10)
    assert np.array_equal(top["weight"].to_numpy(), g["edges_w"])
    got = sorted(zip(top["weight"], top["source"], top["target"]))
    ref = sorted(zip(g["edges_w"], [n8[i] for i in g["edges_src"]], [n8[i] for i in g["edges_dst"]]))
    assert got == ref
    # a larger matrix against the numpy restatement of the same three functions (tests/fake_device.py)
    from fake_device import FakeHandle
    rng = np.random.default_rng(3)
    M = np.round(rng.gamma(1.5, 0.1, (300, 300)), 3); np.fill_diagonal(M, 0.0)        # rounded: plenty of ties
    h, f = S._lib.default_handle(), FakeHandle()
    assert np.allclose(h.clr(M, True), f.clr(M, True), atol=1e-12, rtol=0, equal_nan=True)
    a, b, w = h.top_edges(M, 100); fa, fb, fw = f.top_edges(M, 100)
    assert np.array_equal(a, fa) and np.array_equal(b, fb) and np.array_equal(w, fw)
    truth = (rng.random((300, 300)) < 0.02).astype(np.uint8)
    ga, gx, gy = h.roc(M, truth); ra, rx, ry = f.roc(M, truth)
    assert np.array_equal(gx, rx) and np.array_equal(gy, ry) and ga == pytest.approx(ra, abs=1e-15)


def test_normalize_smooth_on_device(S, golden_configs):
    """causal_model.normalize / .smooth on the matrix resident on the GPU: bit-identical to the reference's pandas results
    (causal_network.py:97-109); the NaN-padded multi-run frame takes the pandas path and gives the same numbers"""
    g = golden_configs
    E = g["norm_in"]
    names = ["g%d" % i for i in range(6)]
    h = S._lib.default_handle()
    m = S.causal_model(pd.DataFrame(E, index=pd.Index(names, name="GENE_ID")))
    m.normalize()
    assert m._expr_stale                                       # ran on the device: the frame is rebuilt on first read
    assert np.array_equal(m.expression.to_numpy(dtype=float), g["norm_out"])
    for w in (3, 7):
        m = S.causal_model(pd.DataFrame(E, index=pd.Index(names, name="GENE_ID")))
        m.smooth(w)
        assert m._expr_stale
        assert np.array_equal(m.expression.to_numpy(dtype=float), g["smooth%d_out" % w])
        assert np.array_equal(np.asarray(m.expression.columns, dtype=np.int64), g["smooth%d_cols" % w])
    runs, g6, _ = recipes.gv_e()
    mm = S.causal_model(recipes.multi_run_frame(runs, g6))
    mm.normalize()
    assert np.allclose(mm.expression.to_numpy(dtype=float), g["norm_multi_out"], atol=0, rtol=0, equal_nan=True)
    # larger panel: every row against pandas, then rdi on the resident normalised matrix == rdi on the pandas-normalised frame
    E = recipes.ar_panel(40, 3000, 11)[:12]
    frame = pd.DataFrame(E, index=pd.Index(["g%02d" % i for i in range(12)], name="GENE_ID"))
    want = frame.sub(frame.mean(axis=1), axis=0)
    want = want.div(want.std(axis=1), axis=0).fillna(0)
    m = S.causal_model(frame)
    m.normalize()
    assert m._expr_stale
    r = m.rdi([1, 5])
    st = h.stats()
    assert st["h2d_bytes"] < 64 * 1024                          # job table only: the matrix was uploaded once, before normalize()
    assert np.array_equal(m.expression.to_numpy(dtype=float), want.to_numpy(dtype=float))
    r2 = S.causal_model(want).rdi([1, 5])
    for d in (1, 5):
        assert np.array_equal(r[d].to_numpy(dtype=float), r2[d].to_numpy(dtype=float), equal_nan=True)


@pytest.mark.parametrize("density", ["kde", "knn"])
@pytest.mark.parametrize("k", [3, 5, 10])
def test_umi_sweep_against_oracle_and_generic(S, k, density):
    """umi on the Euclidean variant of the windowed sweep (no count scan: n_x and W_y are rank ranges) against the oracle and
    against the generic thread-per-query kernel, continuous and tied data, N on both sides of the chunk boundaries"""
    for n, seed, ties in ((300, 1, False), (1000, 2, False), (4097, 3, False), (2500, 4, True)):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((n, 1)); y = 0.6 * x + 0.8 * rng.standard_normal((n, 1))
        if ties:
            x = np.round(x, 1); y = np.round(y, 2)
        kw = dict(k=k, density_estimation_method=density, bw=0.2)
        try:
            ref = o.umi(x, y, **kw)
        except ValueError:
            ref = None
        for path in ("sweep", "generic"):
            opts_path = {"sweep": S._lib.PATH_SWEEP, "generic": S._lib.PATH_GENERIC}[path]
            h = S._lib.default_handle()
            opts = S._lib.make_opts(S._lib.EST_UMI, k, density, 5, 0.2, path=opts_path)
            if ref is None:
                with pytest.raises(ValueError):
                    h.estimate(x, y, None, opts)
            else:
                assert abs(h.estimate(x, y, None, opts) - ref) <= TOL_KDE, (n, path)


def test_post_processing_edge_cases(S):
    """NaN entries, degenerate sizes and empty results of the device post-processing entry points"""
    from fake_device import FakeHandle
    h, f = S._lib.default_handle(), FakeHandle()
    rng = np.random.default_rng(12)
    M = rng.random((7, 7)); M[2, 3] = np.nan; M[5, 5] = np.nan
    with np.errstate(invalid="ignore"):
        want = f.clr(M, True)
    assert np.allclose(h.clr(M, True), want, atol=1e-12, rtol=0, equal_nan=True)
    one = np.array([[0.3, 0.7]])                                  # one row: the column std (ddof = 1) is NaN, the row std is not
    with np.errstate(invalid="ignore", divide="ignore"):
        assert np.allclose(h.clr(one, True), f.clr(one, True), equal_nan=True, atol=1e-12, rtol=0)
    a, b, w = h.top_edges(M, 1000)                                # more requested than exist: every kept entry, NaN dropped
    fa, fb, fw = f.top_edges(M, 1000)
    assert np.array_equal(a, fa) and np.array_equal(b, fb) and np.array_equal(w, fw) and len(w) < 49
    a, b, w = h.top_edges(np.full((4, 4), np.nan), 5)
    assert len(a) == 0 and len(w) == 0
    scores = rng.random((6, 6)); scores[1, 4] = np.nan            # a NaN score is not ranked
    truth = (rng.random((6, 6)) < 0.3).astype(np.uint8); truth[0, 1] = 1; truth[1, 0] = 0
    au, x, y = h.roc(scores, truth)
    i, j = np.nonzero(~np.eye(6, dtype=bool) & ~np.isnan(scores))
    order = np.argsort(-scores[i, j], kind="stable"); hit = truth[i, j][order].astype(bool)
    tp = np.concatenate(([0], np.cumsum(hit))); fp = np.concatenate(([0], np.cumsum(~hit)))
    assert len(x) == len(tp) and np.array_equal(x, fp / fp[-1]) and np.array_equal(y, tp / tp[-1])
    # kNN regressor with fewer data points than k + 1: missing neighbours are at infinity and carry no weight
    out = h.knn_regress([0.0, 1.0, 2.0], [0.0, 0.0, 0.0], [5.0, 7.0, 9.0], [0.1], [0.0], k=5)
    d = np.array([0.9, 1.9]); u = np.exp(-d / d.min()); assert abs(out[0] - (u / u.sum() * np.array([7.0, 9.0])).sum()) < 1e-12


def test_normalize_smooth_multi_run_equal_lengths(S):
    """multi-run frames WITHOUT NaN padding (equal run lengths) also take the device path: every (gene, run) row is one
    segment; the frame rebuilt from the device keeps the (GENE_ID, RUN_ID) index"""
    rng = np.random.default_rng(21)
    runs = [rng.gamma(2.0, 1.0, (5, 120)), rng.gamma(2.0, 1.0, (5, 120))]
    genes = ["g%d" % i for i in range(5)]
    df = recipes.multi_run_frame(runs, genes)
    c = df.sub(df.mean(axis=1), axis=0)
    want = c.div(c.std(axis=1), axis=0).fillna(0)
    m = S.causal_model(df)
    m.normalize()
    assert m._expr_stale
    got = m.expression
    assert list(got.index) == list(df.index) and got.index.names == df.index.names
    assert np.array_equal(got.to_numpy(dtype=float), want.to_numpy(dtype=float))
    r = m.rdi([1, 2]); r2 = S.causal_model(want).rdi([1, 2])
    for d in (1, 2):
        assert np.array_equal(r[d].to_numpy(dtype=float), r2[d].to_numpy(dtype=float), equal_nan=True)
    m = S.causal_model(df)
    m.smooth(4)
    sm = df.T.rolling(window=4).mean().T.dropna(axis=1, how="all")
    assert np.array_equal(m.expression.to_numpy(dtype=float), sm.to_numpy(dtype=float))
    assert list(m.expression.columns) == list(sm.columns) and list(m.expression.index) == list(sm.index)


def test_knn_grid_regressor(S, golden_configs):
    """the 625-query kNN grid regressor of viz_causality (plot/heatmaps.py:389-428) against the reference's own scipy calls"""
    E = recipes.rdi_synthetic(6, 700, 9)
    t = S.heatmaps.causality_grid(E[0], E[1], delay=1, k=5, grid_num=25)
    assert np.array_equal(np.stack([t["x"], t["z"]], axis=1), golden_configs["knnreg_query"])
    assert np.allclose(t["expected_y"], golden_configs["knnreg_norm"], atol=1e-12, rtol=0)
    raw = S.heatmaps.causality_grid(E[0], E[1], delay=1, k=5, grid_num=25, normalized=False)
    assert np.allclose(raw["expected_y"], golden_configs["knnreg_raw"], atol=1e-12, rtol=0)
    # brute-force check of the neighbour selection on other k and a triple
    rng = np.random.default_rng(4)
    x, y, z = rng.standard_normal(900), rng.standard_normal(900), rng.standard_normal(900)
    g = S.heatmaps.comb_logic_grid(x, y, z, k=7, grid_num=11, normalized=False)
    d = np.sqrt((g["x"].to_numpy()[:, None] - x[None, :]) ** 2 + (g["y"].to_numpy()[:, None] - y[None, :]) ** 2)
    idx = np.argsort(d, axis=1, kind="stable")[:, 1:8]
    dd = np.take_along_axis(d, idx, axis=1)
    u = np.exp(-dd / dd.min(axis=1, keepdims=True)); w = u / u.sum(axis=1, keepdims=True)
    assert np.allclose(g["expected_z"], (w * z[idx]).sum(axis=1), atol=1e-12, rtol=0)


def test_multi_gpu_invariance_torchrun():
    """sharded == unsharded, bit for bit, on every GPU count this box offers (tools/check_multi_gpu.py under torchrun)"""
    import os, subprocess, sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for world in sorted({2, n}):
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                            "--master-addr", "127.0.0.1", "--master-port", str(29610 + world),
                            os.path.join(root, "tools", "check_multi_gpu.py")], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and "OK bit-identical" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])
