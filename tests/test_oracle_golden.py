"""Pins oracle/ksg_oracle.py against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import recipes
from oracle import ksg_oracle as o

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
