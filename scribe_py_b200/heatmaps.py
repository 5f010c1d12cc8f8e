"""Data half of the response / causality heat maps: the k-nearest-neighbour grid regressors of
Scribe.plot.heatmaps.viz_causality (Scribe/plot/heatmaps.py:389-428) and viz_comb_logic (:560-600).  The drawing
(matplotlib / seaborn facets) is out of scope; these functions return the tables the reference would draw.

For a gene pair the reference lays a grid_num x grid_num mesh over two coordinates, asks a cKDTree for the k + 1 nearest
cells of every mesh point, drops the nearest, weights the rest with exp(-d / min d) and reports the weighted mean of a
third coordinate.  Here the 625 queries of a pair are one kernel launch (scribe_knn_regress: a warp per mesh point)."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import _lib

__all__ = ["causality_grid", "comb_logic_grid"]


def _mesh(a, b, grid_num):
    am = np.linspace(min(a), max(a), grid_num, endpoint=True)
    bm = np.linspace(min(b), max(b), grid_num, endpoint=True)
    av, bv = np.meshgrid(am, bm)
    return av.reshape(-1), bv.reshape(-1)


def causality_grid(x, y_ori, delay=1, k=5, grid_num=25, log=False, normalized=True):
    """Expected y(t) over a mesh of (x(t - delay), y(t - 1)) for one gene pair (plot/heatmaps.py:389-428).

    x, y_ori: the two genes' expression along the ordering.  Returns a DataFrame with columns x, z, expected_y (divided by
    its maximum when `normalized`, as the reference does unconditionally).  As in the reference the neighbour search runs in
    the (x, y) plane of the cells although the mesh spans (x, z) (:409)."""
    x = np.asarray(x, dtype=np.float64); y_ori = np.asarray(y_ori, dtype=np.float64)
    if log:
        x, y_ori = np.log(x + 1), np.log(y_ori + 1)
    if delay != 0:
        x, y, z = x[:-delay], y_ori[delay:], y_ori[delay - 1:-1]
    else:
        y, z = y_ori, y_ori
    qx, qz = _mesh(x, z, grid_num)
    exp_y = _lib.default_handle().knn_regress(x, y, y, qx, qz, k)
    if normalized:
        finite = exp_y[~np.isnan(exp_y)]
        max_val = finite.max() if len(finite) else np.nan
        if not np.isfinite(max_val):
            max_val = 1e10
        exp_y = exp_y / max_val
    return pd.DataFrame({"x": qx, "z": qz, "expected_y": exp_y})


def comb_logic_grid(x, y, z, k=5, grid_num=25, normalized=True):
    """Expected z over a mesh of (x, y) for a gene triple (plot/heatmaps.py:560-600): the combinatorial-logic map."""
    x, y, z = (np.asarray(v, dtype=np.float64) for v in (x, y, z))
    qx, qy = _mesh(x, y, grid_num)
    exp_z = _lib.default_handle().knn_regress(x, y, z, qx, qy, k)
    if normalized:
        finite = exp_z[~np.isnan(exp_z)]
        max_val = finite.max() if len(finite) else np.nan
        if not np.isfinite(max_val):
            max_val = 1e10
        exp_z = exp_z / max_val
    return pd.DataFrame({"x": qx, "y": qy, "expected_z": exp_z})
