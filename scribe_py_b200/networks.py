"""Top-n edge extraction of the inferred network: the data-frame half of Scribe.plot.networks.vis_causal_net
(Scribe/plot/networks.py:34-47).  The drawing itself (networkx + matplotlib) is out of scope; this returns the edge table it
would draw: the weaker direction of every reciprocal pair is dropped, the remaining edges are ranked by weight and the
top_n_edges strongest are kept.  The ranking runs on the GPU (scribe_top_edges)."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import _lib

__all__ = ["top_n_edges"]


def top_n_edges(adata, key='RDI', top_n_edges=10):
    """DataFrame with columns source / target / weight, strongest edge first (plot/networks.py:34-47).

    adata: an AnnData-like object with uns['causal_net'][key], or the score DataFrame itself."""
    if isinstance(adata, pd.DataFrame):
        df_mat = adata
    else:
        if 'causal_net' not in adata.uns.keys():
            raise Exception('causal_net is not a key in uns slot. Please first run causal network inference with Scribe.')
        df_mat = adata.uns['causal_net'][key]
    M = df_mat.to_numpy(dtype=np.float64)
    n_top = M.size if top_n_edges is None else int(top_n_edges)
    src, dst, w = _lib.default_handle().top_edges(M, n_top)
    return pd.DataFrame({"source": np.asarray(df_mat.index, dtype=object)[src], "target": np.asarray(df_mat.columns, dtype=object)[dst],
                         "weight": w})
