"""Pins oracle/ksg_oracle.py against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import recipes
from oracle import ksg_oracle as o
from oracle import ref_port as rp

SETS = {"a": recipes.gv_a, "b": recipes.gv_b, "poisson": recipes.gv_poisson, "d5": recipes.gv_d5, "dims": recipes.gv_dims}


@pytest.mark.parametrize("name", list(SETS))
def test_counts_bit_exact(golden, name):
    x, y, z = SETS[name]()
    eps, cnt = o.cmi_debug(x, y, z)
    assert np.array_equal(eps, golden[f"{name}_eps"])
    assert np.array_equal(cnt, golden[f"{name}_cnt"])
    dx, dy = x.shape[1], y.shape[1]
    P = np.concatenate((x, y), axis=1)
    ix, iy = tuple(range(dx)), tuple(range(dx, dx + dy))
    eps, cnt = o.knn_counts(P, [ix + iy, ix, iy], 5)
    assert np.array_equal(eps, golden[f"{name}_mi_eps"])
    assert np.array_equal(cnt, golden[f"{name}_mi_cnt"])


@pytest.mark.parametrize("name", list(SETS))
def test_cmi_mi_values(golden, name):
    x, y, z = SETS[name]()
    assert o.cmi(x, y, z) == pytest.approx(float(golden[f"{name}_cmi"]), abs=1e-12)
    assert o.cmi(x, y, z, normalization=True) == pytest.approx(float(golden[f"{name}_cmi_norm"]), abs=1e-12)
    assert o.cmi(x, y, z, k=3) == pytest.approx(float(golden[f"{name}_cmi_k3"]), abs=1e-12)
    assert o.cmi(x, y, z, k=10) == pytest.approx(float(golden[f"{name}_cmi_k10"]), abs=1e-12)
    assert o.mi(x, y) == pytest.approx(float(golden[f"{name}_mi"]), abs=1e-12)
    assert o.mi(x, y, k=3) == pytest.approx(float(golden[f"{name}_mi_k3"]), abs=1e-12)


@pytest.mark.parametrize("name", list(SETS))
def test_uniformised_values(golden, name):
    x, y, z = SETS[name]()
    for key, val in (("cumi_kde", lambda: o.cumi(x, y, z)), ("cumi_kde_bw", lambda: o.cumi(x, y, z, bw=0.05, k=4)),
                     ("cumi_knn", lambda: o.cumi(x, y, z, density_estimation_method="knn")),
                     ("umi_kde", lambda: o.umi(x, y))):
        ref = float(golden[f"{name}_{key}"])
        if np.isnan(ref):
            assert np.isnan(val())
        else:
            assert val() == pytest.approx(ref, abs=1e-9), key
    ref = float(golden[f"{name}_umi_knn"])
    if np.isinf(ref):                       # the reference raised ValueError("math domain error")
        with pytest.raises(ValueError):
            o.umi(x, y, density_estimation_method="knn")
    else:
        assert o.umi(x, y, density_estimation_method="knn") == pytest.approx(ref, abs=1e-9)


@pytest.mark.parametrize("n", [4, 6, 7, 12, 40])
def test_tiny_inputs(golden, n):
    x, y, z = recipes.gv_small(n)
    for got, ref in ((o.cmi(x, y, z), float(golden[f"small{n}_cmi"])), (o.mi(x, y), float(golden[f"small{n}_mi"]))):
        if np.isnan(ref):
            assert np.isnan(got)
        else:
            assert got == pytest.approx(ref, abs=1e-12)


def test_dynamics_coupling(golden):
    spl, vel, genes = recipes.gv_c()
    M = o.dynamics_coupling(spl, vel)
    assert np.allclose(M, golden["c_matrix"], atol=1e-12, rtol=0)
    M2 = o.dynamics_coupling(spl, vel, tf_idx=range(5), target_idx=range(3, 9), normalize=False, drop_zero_cells=False)
    assert np.allclose(M2, golden["c_matrix_nodrop"], atol=1e-12, rtol=0)
    M3 = o.dynamics_coupling(spl, np.abs(vel) * 3, tf_idx=range(4), target_idx=range(6), y_is_sum=False)
    assert np.allclose(M3, golden["c_matrix_t1"], atol=1e-12, rtol=0)


def test_rdi_single_run(golden):
    X, genes, delays = recipes.gv_d()
    runs = [[X[g]] for g in range(len(genes))]
    for tag, kw in (("", {}), ("_u", {"uniformization": True}), ("_diff", {"differential_mode": True})):
        r = o.rdi(runs, delays, **kw)
        for d in delays:
            assert np.allclose(r[d], golden[f"d_rdi{tag}_{d}"], atol=1e-9, rtol=0, equal_nan=True), (tag, d)
        assert np.allclose(r["MAX"], golden[f"d_rdi{tag}_MAX"], atol=1e-9, rtol=0)


def test_rdi_crdi_multi_run(golden):
    runs, genes, delays = recipes.gv_e()
    er = [[r[g] for r in runs] for g in range(len(genes))]
    res = o.rdi(er, delays)
    for d in delays:
        assert np.allclose(res[d], golden[f"e_rdi_{d}"], atol=1e-12, rtol=0, equal_nan=True)
    assert np.allclose(res["MAX"], golden["e_rdi_MAX"], atol=1e-12, rtol=0)
    assert np.allclose(o.crdi(er, res, delays, L=1), golden["e_crdi_L1"], atol=1e-12, rtol=0, equal_nan=True)
    assert np.allclose(o.crdi(er, res, delays, L=2), golden["e_crdi_L2"], atol=1e-12, rtol=0, equal_nan=True)
    ru = o.rdi(er, [1, 2], uniformization=True)
    assert np.allclose(o.crdi(er, ru, [1, 2], L=1, uniformization=True), golden["e_crdi_L1_u"], atol=1e-9, rtol=0,
                       equal_nan=True)


# ---------------------------------------------------------------------------------------------------------------
# oracle/ref_port.py (the CPU arm of every bench ratio) against the same reference-generated fixtures
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["a", "b", "d5"])
def test_ref_port_values(golden, name):
    x, y, z = SETS[name]()
    assert rp.cmi(x, y, z) == pytest.approx(float(golden[f"{name}_cmi"]), abs=1e-12)
    assert rp.cmi(x, y, z, normalization=True) == pytest.approx(float(golden[f"{name}_cmi_norm"]), abs=1e-12)
    assert rp.cmi(x, y, z, k=3) == pytest.approx(float(golden[f"{name}_cmi_k3"]), abs=1e-12)
    assert rp.mi(x, y) == pytest.approx(float(golden[f"{name}_mi"]), abs=1e-12)
    assert rp.cumi(x, y, z) == pytest.approx(float(golden[f"{name}_cumi_kde"]), abs=1e-9)
    assert rp.cumi(x, y, z, bw=0.05, k=4) == pytest.approx(float(golden[f"{name}_cumi_kde_bw"]), abs=1e-9)
    ref = float(golden[f"{name}_cumi_knn"])
    got = rp.cumi(x, y, z, density_estimation_method="knn")
    assert np.isnan(got) if np.isnan(ref) else got == pytest.approx(ref, abs=1e-9)
    assert rp.umi(x, y) == pytest.approx(float(golden[f"{name}_umi_kde"]), abs=1e-9)
    ref = float(golden[f"{name}_umi_knn"])
    if np.isinf(ref):
        with pytest.raises(ValueError):
            rp.umi(x, y, density_estimation_method="knn")
    else:
        assert rp.umi(x, y, density_estimation_method="knn") == pytest.approx(ref, abs=1e-9)


def test_ref_port_rdi_job(golden):
    X, genes, delays = recipes.gv_d()
    for d in delays:
        for a, b in ((0, 1), (1, 0), (3, 2)):
            assert rp.rdi_job((X[a], X[b], d, False)) == pytest.approx(float(golden[f"d_rdi_{d}"][a, b]), abs=1e-12)
            assert rp.rdi_job((X[a], X[b], d, True)) == pytest.approx(float(golden[f"d_rdi_u_{d}"][a, b]), abs=1e-9)


# ---------------------------------------------------------------------------------------------------------------
# config-scale pins (BASELINE configs C2..C5 at their full sizes): sampled jobs, both restatements
# ---------------------------------------------------------------------------------------------------------------
def test_config_c2_sampled(golden_configs):
    E = recipes.rdi_synthetic(100, 2000, 2)
    rng = np.random.default_rng(0)
    for _ in range(6):
        a, b = rng.choice(100, 2, replace=False); d = int(rng.choice([1, 5, 10]))
        ref = float(golden_configs["c2full_rdi_%d" % d][a, b])
        x, y, z = o.lagged_columns([E[a]], [E[b]], d)
        assert o.cmi(x, y, z) == pytest.approx(ref, abs=1e-12)
        assert rp.rdi_job((E[a], E[b], d, False)) == pytest.approx(ref, abs=1e-12)


def test_config_c3_sampled(golden_configs):
    spl, vel, genes = recipes.coupling_panel(500, 10000, 3)
    tf, tg = recipes.C3_TFS[:2], recipes.C3_TARGETS[:3]
    M = o.dynamics_coupling(spl, vel, tf_idx=tf, target_idx=tg)
    assert np.allclose(M, golden_configs["c3_matrix"][:2, :3], atol=1e-12, rtol=0)


def test_config_c4_sampled(golden_configs):
    E = recipes.ar_panel(200, 5000, 4)[recipes.C4_GENES]
    expr = [[E[g]] for g in range(len(E))]
    res = {d: golden_configs["c4_rdi_%d" % d] for d in (1, 5, 10)}
    for a, b, d in ((0, 3, 1), (5, 2, 10)):
        x, y, z = o.lagged_columns([E[a]], [E[b]], d)
        assert o.cmi(x, y, z) == pytest.approx(float(res[d][a, b]), abs=1e-12)
    val, arg = o.max_over_delays(res, [1, 5, 10])
    assert np.array_equal(val, golden_configs["c4_rdi_MAX"])
    jobs = o.crdi_jobs(val, arg, 2)
    for a, b in ((1, 4), (6, 0)):
        d, cn, cd = jobs[(a, b)]
        x, y, z = o.crdi_columns(expr, a, b, d, cn, cd)
        assert o.cmi(x, y, z) == pytest.approx(float(golden_configs["c4_crdi_L2"][a, b]), abs=1e-12)


def test_config_c5_sampled(golden_configs):
    """one uniformised job at N = 19 999 through the reference-shaped port (13 s); the oracle's brute force is pinned on
    the same estimator at smaller N above"""
    E = recipes.zscore_rows(recipes.ar_panel(1000, 20000, 5))[recipes.C5_GENES]
    assert rp.rdi_job((E[2], E[1], 1, True)) == pytest.approx(float(golden_configs["c5_urdi_1"][2, 1]), abs=1e-9)
