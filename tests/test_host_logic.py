"""CPU tests of the host side of the drop-in: with the device replaced by tests/fake_device.FakeHandle (which
evaluates jobs with the oracle) the Python drivers must reproduce the reference-generated fixtures -- this
checks job construction, lag conventions, frames, MAX / argmax, top-incoming selection, cRDI conditioning sets
and the dynamics-coupling preprocessing.  Plus the sharding arithmetic and a world_size-2 gloo all-gather."""
import os
import socket

import numpy as np
import pandas as pd
import pytest

import recipes
from fake_device import FakeHandle


@pytest.fixture()
def S(monkeypatch):
    import scribe_py_b200 as s
    fake = FakeHandle()
    monkeypatch.setattr(s._lib, "default_handle", lambda device=None: fake)
    s._fake = fake
    return s


def test_rdi_single_run_host_logic(S, golden):
    X, genes, delays = recipes.gv_d()
    for tag, kw in (("", {}), ("_diff", {"differential_mode": True})):
        m = S.causal_model(pd.DataFrame(X, index=pd.Index(genes, name="GENE_ID")))
        r = m.rdi(delays, **kw)
        for d in delays:
            assert np.allclose(r[d].to_numpy(dtype=float), golden[f"d_rdi{tag}_{d}"], atol=1e-12, rtol=0, equal_nan=True)
        assert np.allclose(r["MAX"].to_numpy(dtype=float), golden[f"d_rdi{tag}_MAX"], atol=1e-12, rtol=0)
        assert np.isneginf(np.diag(r["MAX"].to_numpy(dtype=float))).all()


def test_rdi_crdi_multi_run_host_logic(S, golden):
    runs, genes, delays = recipes.gv_e()
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    r = m.rdi(delays)
    for d in delays:
        assert np.allclose(r[d].to_numpy(dtype=float), golden[f"e_rdi_{d}"], atol=1e-12, rtol=0, equal_nan=True)
    assert np.allclose(m.crdi(L=1).to_numpy(dtype=float), golden["e_crdi_L1"], atol=1e-12, rtol=0, equal_nan=True)
    assert np.allclose(m.crdi(L=2).to_numpy(dtype=float), golden["e_crdi_L2"], atol=1e-12, rtol=0, equal_nan=True)


def test_nan_padding_and_ragged_runs(S):
    runs, genes, _ = recipes.gv_e()
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    ids, E, ro = m._dense()
    assert list(ro) == [0, 150, 260] and E.shape == (6, 260) and not np.isnan(E).any()
    assert np.array_equal(E[2, 150:], runs[1][2])
    bad = recipes.multi_run_frame(runs, genes)
    bad.iloc[3, 5] = np.nan                      # one gene loses a cell -> lengths differ -> the reference's assertion
    with pytest.raises(AssertionError, match="same length"):
        S.causal_model(bad)._dense()


def test_dynamics_coupling_host_logic(S, golden):
    spl, vel, genes = recipes.gv_c()
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, genes)
    S.causal_net_dynamics_coupling(ad, TFs=genes[:3], Targets=genes[:4])
    M = ad.uns["causal_net"]["RDI"]
    assert M.shape == (3, 4)
    assert np.allclose(M.to_numpy(dtype=float), golden["c_matrix"][:3, :4], atol=1e-12, rtol=0)
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "unspliced": np.abs(vel) * 3}, genes)
    S.causal_net_dynamics_coupling(ad, TFs=genes[:4], Targets=genes[:6], t1_key="unspliced")
    assert np.allclose(ad.uns["causal_net"]["RDI"].to_numpy(dtype=float), golden["c_matrix_t1"], atol=1e-12, rtol=0)
    from scipy import sparse
    ad = recipes.FakeAnnData({"spliced": sparse.csr_matrix(spl), "velocity": vel.copy()}, genes)
    S.causal_net_dynamics_coupling(ad, TFs=genes[:2], Targets=genes[:3])
    assert np.allclose(ad.uns["causal_net"]["RDI"].to_numpy(dtype=float), golden["c_matrix"][:2, :3], atol=1e-12, rtol=0)


def test_clr_matches_definition(S):
    rng = np.random.default_rng(0)
    A = pd.DataFrame(rng.random((5, 5)), index=list("abcde"), columns=list("abcde"))
    out = S.CLR(A)
    m = A.values
    zr = (np.round(m, 10) - m.mean(axis=1)) / m.std(axis=1, ddof=1); zr[zr < 0] = 0
    assert np.allclose(out.values, np.sqrt(zr ** 2))
    assert list(out.index) == list("abcde")


def test_post_processing_against_reference(S, golden_configs, tmp_path):
    """CLR / roc / top-n edges: the shims' argument handling and the formulas (here through the test double) against outputs
    of the reference's own functions"""
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
    assert auroc == pytest.approx(float(g["roc_auroc"]), abs=1e-12)
    assert np.allclose(x, g["roc_x"], atol=1e-15, rtol=0) and np.allclose(y, g["roc_y"], atol=1e-15, rtol=0)
    n8 = ["n%d" % i for i in range(8)]
    top = S.top_n_edges(pd.DataFrame(g["edges_in"], index=n8, columns=n8), top_n_edges=10)
    assert np.array_equal(top["weight"].to_numpy(), g["edges_w"])
    got = sorted(zip(top["weight"], top["source"], top["target"]))
    ref = sorted(zip(g["edges_w"], [n8[i] for i in g["edges_src"]], [n8[i] for i in g["edges_dst"]]))
    assert got == ref                                     # equal weights (the reciprocal tie) may come in either order


def test_matrix_stays_resident_between_calls(S):
    """rdi() uploads the matrix once; a second rdi()/crdi() on the unchanged model reuses it; assigning a frame re-uploads"""
    X, genes, delays = recipes.gv_d()
    m = S.causal_model(pd.DataFrame(X, index=pd.Index(genes, name="GENE_ID")))
    fake = S._fake
    m.rdi(delays); n1 = fake.uploads
    m.rdi([1]); m.crdi(L=1)
    assert fake.uploads == n1
    m.expression = pd.DataFrame(X * 2.0, index=pd.Index(genes, name="GENE_ID"))
    m.rdi([1])
    assert fake.uploads == n1 + 1
    other = S.causal_model(pd.DataFrame(X, index=pd.Index(genes, name="GENE_ID")))
    other.rdi([1])                                        # another model takes the device over ...
    m.rdi([1])                                            # ... so this one uploads again
    assert fake.uploads == n1 + 3
    m.normalize()                                         # host path with the test double (no device normalisation)
    assert np.allclose(m.expression.to_numpy().mean(axis=1), 0, atol=1e-12)


def test_normalize_smooth_host_path_against_reference(S, golden_configs):
    """the pandas paths of normalize()/smooth() (taken with the test double, and on the GPU for NaN-padded frames)"""
    g = golden_configs
    names = ["g%d" % i for i in range(6)]
    m = S.causal_model(pd.DataFrame(g["norm_in"], index=pd.Index(names, name="GENE_ID")))
    m.normalize()
    assert np.array_equal(m.expression.to_numpy(dtype=float), g["norm_out"])
    for w in (3, 7):
        m = S.causal_model(pd.DataFrame(g["norm_in"], index=pd.Index(names, name="GENE_ID")))
        m.smooth(w)
        assert np.array_equal(m.expression.to_numpy(dtype=float), g["smooth%d_out" % w])
        assert np.array_equal(np.asarray(m.expression.columns, dtype=np.int64), g["smooth%d_cols" % w])
    runs, g6, _ = recipes.gv_e()
    mm = S.causal_model(recipes.multi_run_frame(runs, g6))
    mm.normalize()
    assert np.array_equal(mm.expression.to_numpy(dtype=float), g["norm_multi_out"], equal_nan=True)


def test_load_anndata_against_reference(S, golden_configs):
    """load_anndata against the reference's own read_export.load_anndata (one-token pandas fix, see make_golden_configs.py):
    same rows, same (gene, run) order, same pseudotime ordering of every branch, same NaN padding"""
    import make_golden_configs as mg
    ad, genes = mg.load_anndata_case()
    m = S.load_anndata(ad)
    g = golden_configs
    e = m.expression
    assert [str(t[0]) for t in e.index] == list(g["loadann_index_gene"]) and [str(t[1]) for t in e.index] == list(g["loadann_index_run"])
    assert np.array_equal(e.to_numpy(dtype=float), g["loadann_expression"], equal_nan=True)
    assert [str(v) for v in m.node_ids] == list(g["loadann_node_ids"]) and [str(v) for v in m.run_ids] == list(g["loadann_run_ids"])


def test_slab_partition():
    from scribe_py_b200.distributed import slab
    for n in (0, 1, 7, 8, 29700, 249500):
        for ws in (1, 2, 3, 4, 8):
            parts = [slab(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_crdi_differential_mode(S, golden_configs):
    """differential-mode rdi + crdi against the unmodified reference (multi-run branch; per-job np.std scales computed on the host)"""
    from oracle import ksg_oracle as o
    runs, genes, delays = recipes.gv_e()
    m = S.causal_model(recipes.multi_run_frame(runs, genes))
    r = m.rdi(delays, differential_mode=True)
    for d in delays:
        assert np.allclose(r[d].to_numpy(dtype=float), golden_configs["e_rdi_diff_%d" % d], atol=1e-12, rtol=0, equal_nan=True)
    for L in (1, 2):
        c = m.crdi(L=L, differential_mode=True).to_numpy(dtype=float)
        assert np.allclose(c, golden_configs["e_crdi_diff_L%d" % L], atol=1e-12, rtol=0, equal_nan=True), L
    # and the oracle's own differential cRDI against the same fixture
    er = [[rr[g] for rr in runs] for g in range(len(genes))]
    res = {d: golden_configs["e_rdi_diff_%d" % d] for d in delays}
    assert np.allclose(o.crdi(er, res, delays, L=2, differential_mode=True), golden_configs["e_crdi_diff_L2"], atol=1e-12, rtol=0, equal_nan=True)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scribe_py_b200 as s
    from scribe_py_b200.distributed import sharded_evaluate
    seen = []

    def evaluate(lo, hi):
        seen.append((lo, hi))
        return np.arange(lo, hi, dtype=np.float64) ** 2 + 0.5

    out = sharded_evaluate(101, evaluate)
    ok = bool(np.array_equal(out, np.arange(101, dtype=np.float64) ** 2 + 0.5)) and len(seen) == 1
    # a rank that fails must not leave the others waiting in the all-gather: every rank raises
    from scribe_py_b200.distributed import ShardError

    def failing(lo, hi):
        if rank == 1:
            raise ValueError("delay leaves no samples")
        return np.zeros(hi - lo)

    try:
        sharded_evaluate(40, failing)
        ok = False
    except ValueError:
        ok = ok and rank == 1
    except ShardError as e:
        ok = ok and rank == 0 and "[1]" in str(e)
    # the full rdi() driver, sharded: every rank must end up with the complete matrices
    fake = FakeHandle()
    s._lib.default_handle = lambda device=None: fake
    X, genes, delays = recipes.gv_d()
    r = s.causal_model(pd.DataFrame(X, index=pd.Index(genes, name="GENE_ID"))).rdi(delays)
    # cRDI = two sharded phases with the all-gather in between (conditioning sets come from the gathered RDI matrix),
    # and the dynamics-coupling driver: both must give every rank the complete result
    runs, genes_e, delays_e = recipes.gv_e()
    m = s.causal_model(recipes.multi_run_frame(runs, genes_e))
    m.rdi(delays_e)
    crdi = m.crdi(L=2).to_numpy(dtype=float)
    spl, vel, genes_c = recipes.gv_c()
    ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, genes_c)
    s.causal_net_dynamics_coupling(ad, TFs=genes_c[:3], Targets=genes_c[:4])
    coup = ad.uns["causal_net"]["RDI"].to_numpy(dtype=float)
    q.put((rank, ok, seen, fake.calls[:1], r[1].to_numpy(dtype=float), r[2].to_numpy(dtype=float), crdi, coup))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(golden):
    import torch.multiprocessing as mp
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0)); port = s_.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][2] == [(0, 51)] and res[1][2] == [(51, 101)]
    for rank, ok, seen, calls, r1, r2, crdi, coup in res:
        assert ok
        assert calls == [12]                              # 24 jobs (12 ordered pairs x 2 delays) split evenly
        assert np.allclose(r1, golden["d_rdi_1"], atol=1e-12, rtol=0, equal_nan=True)
        assert np.allclose(r2, golden["d_rdi_2"], atol=1e-12, rtol=0, equal_nan=True)
        assert np.allclose(crdi, golden["e_crdi_L2"], atol=1e-12, rtol=0, equal_nan=True)
        assert np.allclose(coup, golden["c_matrix"][:3, :4], atol=1e-12, rtol=0)


def test_load_anndata_branches_as_runs(S):
    """branches become runs ordered by pseudotime; rdi on the loaded model == oracle on that ordering"""
    from oracle import ksg_oracle as o
    rng = np.random.default_rng(8)
    G, n = 4, 90
    X = rng.standard_normal((n, G)).cumsum(axis=0) * 0.3 + rng.standard_normal((n, G))
    branch = np.array(["1"] * 50 + ["2"] * 39 + ["3"])                    # branch "3" has one cell -> dropped
    perm = rng.permutation(n)
    X, branch = X[perm], branch[perm]
    ptime = rng.random(n)
    obs = pd.DataFrame({"dpt_groups": branch, "dpt_pseudotime": ptime}, index=["c%d" % i for i in range(n)])
    genes = ["A", "B", "C", "D"]
    ad = recipes.FakeAnnData({"spliced": X.copy()}, genes, obs_names=list(obs.index), X=X, obs=obs)
    m = S.load_anndata(ad)
    assert list(m.node_ids) == genes and list(m.run_ids) == ["1", "2"]
    assert m.expression.shape == (8, 50) and m.spliced.shape == (n, G) and m.velocity is None
    ids, E, ro = m._dense()
    assert list(ro) == [0, 50, 89]
    runs = []
    for key in ("1", "2"):
        c = np.where(branch == key)[0]; c = c[np.argsort(ptime[c], kind="stable")]
        runs.append(X[c].T)
    assert np.array_equal(E, np.concatenate(runs, axis=1))
    r = m.rdi([1, 2])
    ref = o.rdi([[run[g] for run in runs] for g in range(G)], [1, 2])
    for d in (1, 2):
        assert np.allclose(r[d].to_numpy(dtype=float), ref[d], atol=1e-12, rtol=0, equal_nan=True)
    # no branch annotation -> one run in the given cell order
    ad2 = recipes.FakeAnnData({}, genes, obs_names=list(obs.index), X=X, obs=pd.DataFrame(index=obs.index))
    m2 = S.load_anndata(ad2)
    assert m2.expression.shape == (G, n) and list(m2.run_ids) == [0]
    assert np.array_equal(m2._dense()[1], X.T)


def test_pair_delay_job_table_slices():
    """the lazily materialised rdi job table: any slab equals the same slice of the full delay-major table"""
    from scribe_py_b200.causal_network import _PairDelayJobs
    G = 7
    tgt, src = np.meshgrid(np.arange(G, dtype=np.int32), np.arange(G, dtype=np.int32), indexing="ij")
    off = tgt != src
    tgt, src = tgt[off], src[off]
    delays = [1, 5, 10]
    t = _PairDelayJobs(src, tgt, delays)
    per = G * (G - 1)
    assert len(t) == per * len(delays)
    full = t[0:len(t)]
    assert (full["src"] == np.tile(src, 3)).all() and (full["dst"] == np.tile(tgt, 3)).all()
    assert (full["delay"] == np.repeat(delays, per)).all() and (full["n_cond"] == 0).all()
    rng = np.random.default_rng(0)
    for _ in range(50):
        lo, hi = sorted(rng.integers(0, len(t) + 1, 2))
        part = t[lo:hi]
        assert len(part) == hi - lo and (part == full[lo:hi]).all()
    assert t[5:9] is t[5:9]                                  # a repeated slab request is served from the one-entry cache
    with pytest.raises(TypeError):
        t[3]


def test_coupling_normalisation_fast_path_is_bit_identical():
    """the gene-major numpy normalisation of causal_net_dynamics_coupling must equal the reference's pandas expressions
    bit for bit (it verifies itself on a few genes and otherwise falls back); odd inputs take the pandas path"""
    import scipy.sparse as sp
    from scribe_py_b200 import Scribe as sc
    rng = np.random.default_rng(11)
    N, G = 700, 23
    spl = rng.gamma(2., 1., (N, G)); spl[rng.random((N, G)) < 0.25] = 0; spl[:, 5] = 2.0          # one constant gene: 0/0
    vel = 0.1 * rng.standard_normal((N, G)); vel[rng.random((N, G)) < 0.01] = np.nan
    genes = ["g%02d" % i for i in range(G)]
    want = np.unique(genes[3:20])
    cases = {
        "c-order": (spl, vel), "f-order": (np.asfortranarray(spl), np.asfortranarray(vel)),
        "float32": (spl.astype(np.float32), vel.astype(np.float32)),
        "sparse": (sp.csr_matrix(spl), sp.csr_matrix(np.nan_to_num(vel))),
    }
    for name, (a, b) in cases.items():
        for t1 in ("velocity", "other"):
            ad = recipes.FakeAnnData({"spliced": a, t1: b}, genes)
            with np.errstate(all="ignore"):
                ref = sc._normalised_layers_pandas(ad, want, "spliced", t1, True)
                got = sc._normalised_layers(ad, want, "spliced", t1, True)
            assert got[2] == ref[2], name
            assert np.array_equal(got[0], ref[0], equal_nan=True) and np.array_equal(got[1], ref[1], equal_nan=True), (name, t1)
            assert got[0].flags["C_CONTIGUOUS"] and got[0].shape == (len(want), N)
    ad = recipes.FakeAnnData({"spliced": spl, "velocity": vel}, genes)
    raw = sc._normalised_layers(ad, want, "spliced", "velocity", False)                     # normalize=False: pandas path
    assert np.array_equal(raw[0], spl[:, 3:20].T)


def test_top_incoming_matches_reference_loop(S):
    """the all-targets-at-once form of the streaming top-(L+1) selection == the reference's double loop
    (causal_network.py:316-333), slot for slot, on matrices with ties, -inf and NaN diagonals"""
    def loop_version(V, Dl, node_ids, k_nodes):
        top_nodes, top_delays, top_values = {}, {}, {}
        for b_i, b in enumerate(node_ids):
            nodes = [None] * (k_nodes + 1); dls = [np.nan] * (k_nodes + 1); vals = [-np.inf] * (k_nodes + 1)
            for a_i, a in enumerate(node_ids):
                if a_i == b_i:
                    continue
                lo = min(vals)
                if V[a_i, b_i] > lo:
                    s = vals.index(lo)
                    nodes[s] = a; dls[s] = Dl[a_i, b_i]; vals[s] = V[a_i, b_i]
            top_nodes[b], top_delays[b], top_values[b] = nodes, dls, vals
        return top_nodes, top_delays, top_values

    rng = np.random.default_rng(21)
    for G, L, ties in ((9, 1, False), (17, 2, True), (6, 4, True), (3, 3, False)):
        genes = ["g%02d" % i for i in range(G)]
        V = rng.standard_normal((G, G))
        if ties:
            V = np.round(V, 0)                                   # many equal values, also equal to the running minimum
        np.fill_diagonal(V, -np.inf)
        Dl = rng.integers(1, 4, (G, G)).astype(object)
        m = S.causal_model(pd.DataFrame(rng.standard_normal((G, 30)), index=pd.Index(genes, name="GENE_ID")))
        fn = getattr(m, "_causal_model__extract_top_incoming_nodes_delays")
        got = fn(pd.DataFrame(V, index=genes, columns=genes), pd.DataFrame(Dl, index=genes, columns=genes, dtype=object), L)
        ref = loop_version(V, Dl, genes, L)
        for gd, rd in zip(got, ref):
            for b in genes:
                assert len(gd[b]) == L + 1
                for x, y in zip(gd[b], rd[b]):
                    assert (x == y) or (isinstance(x, float) and isinstance(y, float) and np.isnan(x) and np.isnan(y)), (G, L, b, gd[b], rd[b])
