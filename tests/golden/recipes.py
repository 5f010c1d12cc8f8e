"""Seeded synthetic inputs shared by make_golden.py (which runs the real reference on them) and the tests."""
import numpy as np


def gv_a():
    """continuous, N=500 (SURVEY GV-A)"""
    rng = np.random.default_rng(0); N = 500
    x = rng.random((N, 1)); z = rng.random((N, 1)); y = x + 0.3 * z + 0.1 * rng.standard_normal((N, 1))
    return x, y, z


def gv_b():
    """zero-inflated gamma -> many exact ties, N=800 (SURVEY GV-B)"""
    rng = np.random.default_rng(7); N = 800
    x = np.where(rng.random((N, 1)) < .3, 0, rng.gamma(2, 1, (N, 1)))
    z = np.where(rng.random((N, 1)) < .3, 0, rng.gamma(2, 1, (N, 1)))
    y = x + 0.3 * z + 0.1 * rng.standard_normal((N, 1))
    return x, y, z


def gv_poisson():
    """integer counts: eps = 0 for most points, heavy ties at every distance, N=600"""
    rng = np.random.default_rng(11); N = 600
    x = rng.poisson(1.5, (N, 1)).astype(float); z = rng.poisson(2.0, (N, 1)).astype(float)
    y = (x + rng.poisson(1.0, (N, 1))).astype(float)
    return x, y, z


def gv_d5():
    """cRDI-shaped job: dx=dy=1, dz=3, N=700"""
    rng = np.random.default_rng(21); N = 700
    z = rng.standard_normal((N, 3)); x = 0.5 * z[:, :1] + rng.standard_normal((N, 1))
    y = x + z[:, 1:2] - 0.5 * z[:, 2:3] + 0.3 * rng.standard_normal((N, 1))
    return x, y, z


def gv_dims():
    """multi-column x and y (generic kernel): dx=2, dy=2, dz=2, N=300"""
    rng = np.random.default_rng(22); N = 300
    z = rng.standard_normal((N, 2)); x = rng.standard_normal((N, 2)) + z
    y = x[:, ::-1] * 0.7 + 0.4 * rng.standard_normal((N, 2))
    return x, y, z


def gv_small(n):
    rng = np.random.default_rng(100 + n)
    x = rng.standard_normal((n, 1)); z = rng.standard_normal((n, 1)); y = x + z + 0.5 * rng.standard_normal((n, 1))
    return x, y, z


def gv_c():
    """dynamics coupling, 20 genes x 500 cells (SURVEY GV-C == BASELINE config C1)"""
    rng = np.random.default_rng(0); G, N = 20, 500
    spl = rng.gamma(2., 1., (N, G)); spl[rng.random((N, G)) < 0.2] = 0
    vel = 0.1 * rng.standard_normal((N, G)); vel[:, 1] += 0.5 * (spl[:, 0] - spl[:, 0].mean())
    genes = ["g%02d" % i for i in range(G)]
    return spl, vel, genes


def gv_d():
    """single-run RDI, 4 genes x 200 cells, delays [1, 2] (SURVEY GV-D)"""
    rng = np.random.default_rng(1); G, T = 4, 200
    X = np.cumsum(rng.standard_normal((G, T)), axis=1); X[1, 1:] += 0.8 * X[0, :-1]
    return X, ["g%d" % i for i in range(G)], [1, 2]


def gv_e():
    """multi-run RDI + cRDI, 6 genes, runs of 150 and 110 cells (SURVEY GV-E)"""
    rng = np.random.default_rng(4); G = 6; lens = [150, 110]
    runs = []
    for T in lens:
        X = 0.2 * np.cumsum(rng.standard_normal((G, T)), 1) + rng.standard_normal((G, T))
        X[1, 2:] += 0.9 * X[0, :-2]; X[2, 1:] += 0.7 * X[1, :-1]
        runs.append(X)
    return runs, ["g%d" % i for i in range(G)], [1, 2, 3]


def multi_run_frame(runs, genes):
    """NaN-padded MultiIndex (GENE_ID, RUN_ID) x cells DataFrame, as read_expression_file would build."""
    import pandas as pd
    Tmax = max(r.shape[1] for r in runs)
    rows, idx = [], []
    for gi, g in enumerate(genes):
        for ri, r in enumerate(runs):
            v = np.full(Tmax, np.nan); v[:r.shape[1]] = r[gi]
            rows.append(v); idx.append((g, "r%d" % ri))
    return pd.DataFrame(np.array(rows), index=pd.MultiIndex.from_tuples(idx, names=["GENE_ID", "RUN_ID"]))


def rdi_synthetic(G, T, seed):
    """C2-style generator: sparse regulator graph, AR(1) + tanh coupling at delays {1,5,10}."""
    rng = np.random.default_rng(seed)
    parents = [rng.choice([g for g in range(G) if g != t], size=2, replace=False) for t in range(G)]
    pdel = [rng.choice([1, 5, 10], size=2) for _ in range(G)]
    coef = rng.uniform(0.4, 0.9, (G, 2)) * rng.choice([-1, 1], (G, 2))
    E = np.zeros((G, T + 50)); E[:, :10] = rng.standard_normal((G, 10))
    noise = 0.3 * rng.standard_normal((G, T + 50))
    for t in range(10, T + 50):
        for g in range(G):
            v = 0.9 * E[g, t - 1] + noise[g, t]
            for p, d, a in zip(parents[g], pdel[g], coef[g]):
                v += a * np.tanh(E[p, t - d])
            E[g, t] = v
    return E[:, 50:]


def ar_panel(G, T, seed):
    """cheap autocorrelated expression panel with sparse lagged couplings (vectorised over genes): the C4 / C5 generator
    (BASELINE configs[3], configs[4]; C5 z-scores each gene afterwards, as causal_model.normalize() would)"""
    rng = np.random.default_rng(seed)
    E = np.zeros((G, T + 60)); E[:, :12] = rng.standard_normal((G, 12))
    par = rng.integers(0, G, (G, 2)); dl = rng.choice([1, 5, 10], (G, 2)); co = rng.uniform(0.4, 0.9, (G, 2)) * rng.choice([-1, 1], (G, 2))
    noise = 0.3 * rng.standard_normal((G, T + 60))
    for t in range(12, T + 60):
        E[:, t] = 0.9 * E[:, t - 1] + noise[:, t] + co[:, 0] * np.tanh(E[par[:, 0], t - dl[:, 0]]) + co[:, 1] * np.tanh(E[par[:, 1], t - dl[:, 1]])
    return E[:, 60:]


def zscore_rows(E):
    return (E - E.mean(1, keepdims=True)) / E.std(1, keepdims=True)


def coupling_panel(G, N, seed=3):
    """C3 generator (BASELINE configs[2]): spliced = gamma(2,1) with 20 % zeros, velocity = sparse linear coupling of two
    parents - 0.2 * spliced + noise.  Returns cells x genes layers and the gene names."""
    rng = np.random.default_rng(seed)
    spl = rng.gamma(2., 1., (N, G)); spl[rng.random((N, G)) < 0.2] = 0
    par = rng.integers(0, G, (G, 2)); co = rng.uniform(-0.5, 0.5, (G, 2))
    vel = co[:, 0] * (spl[:, par[:, 0]] - spl[:, par[:, 0]].mean(0)) + co[:, 1] * (spl[:, par[:, 1]] - spl[:, par[:, 1]].mean(0)) \
        - 0.2 * spl + 0.1 * rng.standard_normal((N, G))
    # (fancy indexing leaves `vel` in a non-C memory order; an AnnData layer is a plain C-contiguous cells x genes array)
    return np.ascontiguousarray(spl), np.ascontiguousarray(vel), ["g%04d" % i for i in range(G)]


# gene subsets of the named configs on which the UNMODIFIED reference was run (tests/golden/make_golden_configs.py)
C3_TFS, C3_TARGETS = [3, 77, 140, 251, 402], [9, 77, 118, 205, 251, 333, 420, 499]
C4_GENES = [5, 23, 61, 99, 134, 170, 198]
C5_GENES = [12, 345, 678, 901]


class FakeAnnData:
    """Duck-typed AnnData: var_names / obs_names (pandas Index), layers (cells x genes), uns, [:, genes]."""

    def __init__(self, layers, var_names, obs_names=None, uns=None, X=None, obs=None):
        import pandas as pd
        self.X = X
        self.obs = obs
        self.var_names = pd.Index(var_names)
        n = next(iter(layers.values())).shape[0] if layers else X.shape[0]
        self.obs_names = pd.Index(obs_names if obs_names is not None else ["c%d" % i for i in range(n)])
        self.layers = layers
        self.uns = {} if uns is None else uns

    def obs_keys(self):
        return [] if self.obs is None else list(self.obs.columns)

    def __getitem__(self, key):
        _, genes = key
        idx = [self.var_names.get_loc(g) for g in genes]
        return FakeAnnData({k: v[:, idx] for k, v in self.layers.items()}, [self.var_names[i] for i in idx],
                           self.obs_names, self.uns)
