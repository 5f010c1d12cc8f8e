"""scribe_py_b200 -- B200-native (sm_100a) implementation of Scribe-py's pairwise-causality hot path.

Module layout mirrors the reference package for the path it replaces (Scribe/__init__.py:15-18):
    information_estimators : vd, mi, cmi, umi, cumi
    causal_network         : causal_model (.rdi, .crdi, preprocessing pass-throughs)
    Scribe                 : causal_net_dynamics_coupling, CLR
    other_estimators       : mi, corr (pairwise networks over a causal_model)
    read_export            : load_anndata (AnnData -> causal_model, branches as runs ordered by pseudotime)
    networks               : top_n_edges (the edge table behind plot.networks.vis_causal_net)
    heatmaps               : causality_grid, comb_logic_grid (the kNN grid regressors behind plot.heatmaps)
All estimator arithmetic runs in libscribe_b200.so (CUDA, FP64); there is no CPU fallback.
"""
from . import _lib
from . import information_estimators
from . import causal_network
from . import Scribe
from . import read_export
from . import other_estimators
from . import networks
from . import heatmaps
from .information_estimators import vd, mi, cmi, umi, cumi
from .causal_network import causal_model
from .Scribe import causal_net_dynamics_coupling, CLR
from .read_export import load_anndata
from .networks import top_n_edges

__all__ = ["information_estimators", "causal_network", "Scribe", "vd", "mi", "cmi", "umi", "cumi", "causal_model",
           "causal_net_dynamics_coupling", "CLR", "load_anndata", "read_export", "other_estimators", "networks", "top_n_edges", "heatmaps"]
__version__ = "0.1.0"
