"""Drop-in for Scribe/other_estimators.py: pairwise Pearson and mutual-information networks over a causal_model
(other_estimators.py:14-83).  In the reference both are dead code (the module-level helpers are shadowed by the public
functions of the same name, :10-11 vs :50, and are called with one tuple argument, :40,:75); here `mi` is one batched
submission of all ordered gene pairs to the dense KSG kernel and `corr` is numpy's correlation matrix."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import _lib
from . import distributed as _dist

__all__ = ["corr", "mi"]


def _rows(self):
    node_ids = list(self.node_ids)
    conc = getattr(self, "expression_concatenated", None)
    if conc is None:                      # models built by load_anndata / the constructor: concatenate the runs
        _, E, _ = self._dense()
        return node_ids, E
    E = np.ascontiguousarray(conc.loc[node_ids].to_numpy(dtype=np.float64))
    return node_ids, E


def corr(self, number_of_processes=1):
    """Pairwise Pearson correlation network; NaN diagonal as in the reference's result frame."""
    node_ids, E = _rows(self)
    M = np.corrcoef(E)
    np.fill_diagonal(M, np.nan)
    self.corr_results = pd.DataFrame(M, index=node_ids, columns=node_ids)
    return self.corr_results


def mi(self, number_of_processes=1):
    """Pairwise KSG mutual information network (k = 5) over all cells; rows = id1, columns = id2, NaN diagonal."""
    node_ids, E = _rows(self)
    assert not np.isnan(E).any(), "Lists should have same length"      # NaN padding cannot enter the estimator
    G = len(node_ids)
    a, b = np.nonzero(~np.eye(G, dtype=bool))
    h = _lib.default_handle()
    h.upload_matrix(E)
    opts = _lib.make_opts(_lib.EST_MI, 5)
    vals = _dist.sharded_evaluate(len(a), lambda lo, hi: h.mi_jobs(a[lo:hi], b[lo:hi], opts))
    M = np.full((G, G), np.nan)
    M[a, b] = vals
    self.mi_results = pd.DataFrame(M, index=node_ids, columns=node_ids)
    return self.mi_results
