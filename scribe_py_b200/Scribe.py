"""Drop-in for Scribe/Scribe.py: `causal_net_dynamics_coupling` (Scribe.py:10-139), plus CLR / check_symmetric.

AnnData in, `adata.uns['causal_net'] = {'RDI': DataFrame}` out.  The per-pair Python loop of the reference
(Scribe.py:110-135) becomes one batched submission: both layers are normalised with the same pandas
expressions, uploaded once, and every (TF, target) pair -- zero-cell mask, compaction and cmi -- is evaluated
on the GPU(s).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.sparse import issparse

from . import _lib
from . import distributed as _dist

_DEVICE_NORMALISATION_VERIFIED = set()      # (normalize, y_is_sum) combinations whose device result matched pandas on probe genes
CLR_DDOF = 1      # the reference's module constant (Scribe.py:8); scribe_clr uses ddof = 1

__all__ = ["causal_net_dynamics_coupling", "CLR", "check_symmetric"]


def _layer_frame(adata, genes, key):
    sub = adata[:, genes]
    mat = sub.layers[key]
    if issparse(mat):
        mat = mat.todense()
    df = pd.DataFrame(np.asarray(mat))
    df.index = adata.obs_names
    df.columns = sub.var_names
    return df


def _normalised_layers_pandas(adata, genes, t0_key, t1_key, normalize):
    """Scribe.py:88-121 with the reference's own pandas expressions.  Returns S, Y gene-major (genes x cells) and the gene order."""
    spliced = _layer_frame(adata, genes, t0_key)
    velocity = _layer_frame(adata, genes, t1_key)
    velocity[pd.isna(velocity)] = 0
    if normalize:                                    # Scribe.py:104-106, same pandas expressions
        spliced = (spliced - spliced.mean()) / (spliced.max() - spliced.min())
        velocity = (velocity - velocity.mean()) / (velocity.max() - velocity.min())
    y = (spliced + velocity) if t1_key == 'velocity' else velocity            # Scribe.py:121
    S = np.ascontiguousarray(spliced.to_numpy(dtype=np.float64).T)
    Y = np.ascontiguousarray(y.to_numpy(dtype=np.float64).T)
    return S, Y, list(spliced.columns)


def _normalised_layers(adata, genes, t0_key, t1_key, normalize):
    """The same result computed gene-major in numpy -- one transposed copy per layer instead of five frame-sized pandas
    temporaries -- when that is provably the same arithmetic: dense float64 layers without NaN in the t0 layer, and a
    check on a few genes that the pandas expressions give bit-identical numbers with the installed pandas (its column
    mean is a pairwise sum over a contiguous column; a different pandas build may sum differently, in which case
    everything goes through pandas).  Anything else takes the pandas path."""
    sub = adata[:, genes]
    a, b = sub.layers[t0_key], sub.layers[t1_key]
    ok = normalize and all(isinstance(m, np.ndarray) and m.dtype == np.float64 and m.ndim == 2 for m in (a, b)) and a.shape == b.shape \
        and a.shape[0] > 0
    if not ok:
        return _normalised_layers_pandas(adata, genes, t0_key, t1_key, normalize)
    At = np.ascontiguousarray(a.T)
    if np.isnan(At).any():
        return _normalised_layers_pandas(adata, genes, t0_key, t1_key, normalize)
    Bt = np.array(b.T, dtype=np.float64, order="C")
    Bt[np.isnan(Bt)] = 0

    def norm(M):
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = M.sum(axis=1) / M.shape[1]
            return (M - mean[:, None]) / (M.max(axis=1) - M.min(axis=1))[:, None]

    S, V = norm(At), norm(Bt)
    Y = (S + V) if t1_key == 'velocity' else V
    order = list(sub.var_names)
    probe = sorted(set([0, len(order) // 2, len(order) - 1]))
    Sp, Yp, _ = _normalised_layers_pandas(adata, [order[i] for i in probe], t0_key, t1_key, normalize)
    if not (np.array_equal(Sp, S[probe], equal_nan=True) and np.array_equal(Yp, Y[probe], equal_nan=True)):
        return _normalised_layers_pandas(adata, genes, t0_key, t1_key, normalize)
    return S, Y, order


def _dense_layers(adata, genes, t0_key, t1_key):
    """The two layers restricted to `genes` (in that order) as dense cells x genes float64 arrays, without a copy when the
    gene list is the AnnData's own column order (the reference goes through adata[:, genes] and a DataFrame, Scribe.py:88-99)."""
    idx = adata.var_names.get_indexer(list(genes))
    whole = len(idx) == len(adata.var_names) and np.array_equal(idx, np.arange(len(idx)))
    out = []
    for key in (t0_key, t1_key):
        mat = adata.layers[key]
        if issparse(mat):
            mat = np.asarray(mat.todense())
        mat = np.asarray(mat)
        if not whole:
            mat = mat[:, idx]
        out.append(np.ascontiguousarray(mat, dtype=np.float64))
    return out[0], out[1]


def _upload_layers(h, adata, genes, t0_key, t1_key, normalize):
    """Make S / Y of `genes` resident on the handle's device.  Preferred: the raw layers go up as they are and the device
    normalises them (scribe_upload_coupling_raw; numpy's summation order, so the same float64 as the pandas expressions of
    Scribe.py:104-106) -- checked on three probe genes against those pandas expressions; anything that does not fit (NaN in the
    t0 layer, a pandas build that sums differently) is normalised on the host and uploaded with scribe_upload_coupling."""
    order = list(genes)
    y_is_sum = t1_key == 'velocity'
    try:
        a, b = _dense_layers(adata, order, t0_key, t1_key)
        ok = a.ndim == 2 and a.shape == b.shape and a.shape[0] > 0 and hasattr(h, "upload_coupling_raw")
    except Exception:
        ok = False
    if ok and h.upload_coupling_raw(a, b, normalize, y_is_sum):
        # The device follows numpy's pairwise summation; whether the installed pandas sums a column the same way is a property
        # of the pandas / numpy build, not of the data, so the probe runs until it has succeeded once in this process.
        key = (bool(normalize), y_is_sum)
        if key in _DEVICE_NORMALISATION_VERIFIED:
            return order
        probe = sorted(set([0, len(order) // 2, len(order) - 1]))
        Sp, Yp, _ = _normalised_layers_pandas(adata, [order[i] for i in probe], t0_key, t1_key, normalize)
        same = all(np.array_equal(h.read_coupling_row(0, g, a.shape[0]), Sp[i], equal_nan=True) and
                   np.array_equal(h.read_coupling_row(1, g, a.shape[0]), Yp[i], equal_nan=True) for i, g in enumerate(probe))
        if same:
            if a.shape[0] >= 1024:                   # long columns exercise every branch of the pairwise order (blocks of 128, split points)
                _DEVICE_NORMALISATION_VERIFIED.add(key)
            return order
    S, Y, order = _normalised_layers(adata, genes, t0_key, t1_key, normalize)
    h.upload_coupling(S, Y)
    return order


def _coupling_plan(adata, TFs=None, Targets=None, t0_key='spliced', t1_key='velocity', normalize=True):
    """Argument handling of causal_net_dynamics_coupling (Scribe.py:66-86) and the job table: every (TF, target) pair with
    different names, target-major so that a rank's slab shares its targets' y / z columns."""
    if TFs is None:
        TFs = adata.var_names.tolist()
    else:
        TFs = adata.var_names.intersection(TFs).tolist()
        if len(TFs) == 0:
            raise Exception("The adata object has no gene names from .var_name that intersects with the TFs list you provided")
    if Targets is None:
        Targets = adata.var_names.tolist()
    else:
        Targets = adata.var_names.intersection(Targets).tolist()
        if len(Targets) == 0:
            raise Exception("The adata object has no gene names from .var_name that intersect with the Targets list you provided")
    genes = np.unique(TFs + Targets)
    col = {g: i for i, g in enumerate(genes)}
    ia = np.array([col[g] for g in TFs], dtype=np.int32)
    ib = np.array([col[g] for g in Targets], dtype=np.int32)
    bb, aa = np.meshgrid(np.arange(len(Targets)), np.arange(len(TFs)), indexing="ij")
    names_differ = ib[:, None] != ia[None, :]        # skip g_a == g_b (Scribe.py:112); names are unique, so compare their positions
    aa, bb = aa[names_differ], bb[names_differ]
    return {"TFs": TFs, "Targets": Targets, "genes": genes, "aa": aa, "bb": bb, "src": np.ascontiguousarray(ia[aa]),
            "dst": np.ascontiguousarray(ib[bb]), "t0_key": t0_key, "t1_key": t1_key, "normalize": normalize}


def _coupling_execute(plan, drop_zero_cells=True):
    """Evaluate the plan's pairs against the S / Y matrices resident on this rank's GPU; every rank gets every result."""
    h = _lib.default_handle()
    src, dst = plan["src"], plan["dst"]
    opts = _lib.make_opts(_lib.EST_CMI, 5)

    def evaluate(lo, hi, out_dev_ptr=None):
        return h.coupling_jobs(src[lo:hi], dst[lo:hi], opts, drop_zero_cells, out_dev_ptr=out_dev_ptr)

    return _dist.sharded_evaluate(len(src), evaluate, h)


def causal_net_dynamics_coupling(adata, TFs=None, Targets=None, guide_keys=None, t0_key='spliced', t1_key='velocity',
                                 normalize=True, drop_zero_cells=True, copy=False):
    """Infer a causal network from dynamics-coupled single-cell measurements: for every TF g_a and target g_b,
    I(spliced[g_a]; y[g_b] | spliced[g_b]) with y = spliced + velocity when t1_key == 'velocity', else the
    t1 layer itself.  Arguments and result are those of the reference (Scribe.py:10-64)."""
    plan = _coupling_plan(adata, TFs, Targets, t0_key, t1_key, normalize)
    if guide_keys is not None:                       # computed and unused, as in the reference (Scribe.py:80-86)
        guides = np.unique(adata.obs[guide_keys].tolist())
        guides = np.setdiff1d(guides, ['*', 'nan', 'neg'])
    _upload_layers(_lib.default_handle(), adata, plan["genes"], t0_key, t1_key, normalize)
    vals = _coupling_execute(plan, drop_zero_cells)
    TFs, Targets = plan["TFs"], plan["Targets"]
    M = np.zeros((len(TFs), len(Targets)))           # the skipped g_a == g_b entries: NaN -> 0 by fillna (Scribe.py:137)
    M[plan["aa"], plan["bb"]] = np.where(np.isnan(vals), 0.0, vals)
    adata.uns['causal_net'] = {"RDI": pd.DataFrame(M, index=TFs, columns=Targets)}
    return adata if copy else None


def CLR(causality_mat, zscore_both_dim=None):
    """Context likelihood of relatedness of a causality matrix (Scribe.py:142-176), computed on the GPU
    (scribe_clr: row / column statistics and the element-wise z-scores are three small kernels over the matrix).
    Keeps the reference's quirk for square input: the (n,) row statistics are broadcast along the last axis, so entry
    (i, j) is z-scored against row j (Scribe.py:151-154)."""
    zscore_both_dim = check_symmetric(causality_mat) if zscore_both_dim is None else zscore_both_dim
    if isinstance(causality_mat, pd.DataFrame):
        mat = causality_mat.values
    elif issparse(causality_mat):
        mat = causality_mat.toarray()
    else:
        mat = causality_mat
    res = _lib.default_handle().clr(np.asarray(mat, dtype=np.float64), bool(zscore_both_dim))
    if isinstance(causality_mat, pd.DataFrame):
        res = pd.DataFrame(res)
        res.index, res.columns = causality_mat.index, causality_mat.columns
    return res


def check_symmetric(a, rtol=1e-05, atol=1e-08):
    """(Scribe.py:179-192)"""
    if a.shape[0] != a.shape[1]:
        return False
    if isinstance(a, pd.DataFrame):
        a = a.fillna(-1).values
    if issparse(a):
        a.data = np.nan_to_num(a.data)
        return (abs(a - a.T) > atol).nnz == 0
    return np.allclose(a, a.T, rtol=rtol, atol=atol)
