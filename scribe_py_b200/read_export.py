"""Drop-in for Scribe/read_export.py: `load_anndata` -- the AnnData-in seam of the RDI path (read_export.py:16-98).

Cells are grouped into runs by branch (scanpy 'dpt_groups' / Monocle 'Branch') and ordered by pseudotime inside each
run; the result is the MultiIndex (GENE_ID, RUN_ID) x cells frame, NaN-padded to the longest run, that
`causal_model.rdi/.crdi` consume.  Differences from the reference, which cannot run on pandas >= 2: without branch
annotations the model gets a single-run genes x cells DataFrame (the reference stored a bare ndarray that `rdi`
cannot index, read_export.py:74-78), and a leading one-cell branch no longer breaks the concatenation (:51-52,:66).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.sparse import issparse

from .causal_network import causal_model

__all__ = ["load_anndata"]


def _dense(m):
    return m.toarray() if issparse(m) else np.asarray(m)


def load_anndata(anndata, keys=None):
    """Convert an AnnData object to a `causal_model` (same arguments and attributes as the reference)."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = causal_model()
    X = _dense(anndata.X)                                   # cells x genes
    genes = list(anndata.var_names)
    obs_keys = list(anndata.obs_keys()) if hasattr(anndata, "obs_keys") else list(anndata.obs.columns)
    if "dpt_groups" in obs_keys:
        branch_key, pseudotime_key = "dpt_groups", "dpt_pseudotime"
    elif "Branch" in obs_keys:
        branch_key, pseudotime_key = "Branch", "Pseudotime"
    else:
        branch_key = None

    if branch_key is not None:
        branch = np.array([str(b) for b in anndata.obs[branch_key].tolist()])
        ptime = np.asarray(anndata.obs[pseudotime_key].values, dtype=float)
        runs = []
        for key in sorted(set(branch)):
            cells = np.where(branch == key)[0]
            if len(cells) == 1:                              # read_export.py:51-52
                continue
            cells = cells[np.argsort(ptime[cells], kind="stable")]
            runs.append((key, X[cells].T))                   # genes x cells of this run, pseudotime-ordered
        if not runs:
            raise ValueError("no branch with more than one cell")
        width = max(r.shape[1] for _, r in runs)
        rows, index = [], []
        for key, block in runs:
            pad = np.full((len(genes), width), np.nan)
            pad[:, :block.shape[1]] = block
            rows.append(pad)
            index += [(g, key) for g in genes]
        frame = pd.DataFrame(np.concatenate(rows, axis=0), index=pd.MultiIndex.from_tuples(index, names=("GENE_ID", "RUN_ID")))
        model.node_ids = frame.index.levels[0]
        model.run_ids = frame.index.levels[1]
    else:
        frame = pd.DataFrame(X.T, index=pd.Index(genes, name="GENE_ID"))
        model.node_ids = frame.index
        model.run_ids = [0]

    model.X = pd.DataFrame(X.T, index=anndata.var_names, columns=anndata.obs_names)
    model.expression_raw = frame
    model.expression = frame
    layers = getattr(anndata, "layers", {})
    for name in ("unspliced", "spliced", "velocity"):
        setattr(model, name, _dense(layers[name]) if name in layers.keys() else None)
    return model
