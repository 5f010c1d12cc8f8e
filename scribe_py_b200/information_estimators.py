"""Drop-in for Scribe/information_estimators.py: `vd`, `mi`, `cmi`, `umi`, `cumi`.

Same names, positional/keyword signatures, defaults, assertion and exception behaviour as the reference
(information_estimators.py:15-30, :59-109, :111-173, :209-277, :326-403); the arithmetic runs in the
sm_100a kernels behind libscribe_b200.so.  Inputs may be list-of-lists ``[[v], [v], ...]`` as the
reference's drivers pass them, or (N, d) arrays.
"""
from __future__ import annotations

from math import log, pi

import numpy as np
from scipy.special import gamma

from . import _lib

__all__ = ["vd", "mi", "cmi", "umi", "cumi", "ksg_debug"]


def _handle():
    return _lib.default_handle()


def _dim(a):
    return len(a[0])


def _as2d(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return a


def vd(d):
    """log-volume of the d-dimensional Euclidean unit ball (information_estimators.py:15-30)."""
    return 0.5 * d * log(pi) - log(gamma(0.5 * d + 1))


def mi(x_orig, y_orig, use_rank_order=False, k=5):
    """KSG mutual information (information_estimators.py:59-109); `use_rank_order` is ignored there too."""
    assert len(x_orig) == len(y_orig), "Lists should have same length"
    x, y = _as2d(x_orig), _as2d(y_orig)
    return np.float64(_handle().estimate(x, y, None, _lib.make_opts(_lib.EST_MI, k)))


def _normalise(x, y, z):
    # x /= np.std(x) etc.: population std over the whole array (information_estimators.py:150-153, :370-373)
    return x / np.std(x), y / np.std(y), z / np.std(z)


def cmi(x_orig, y_orig, z_orig, normalization=False, k=5):
    """KSG conditional mutual information I(x; y | z) (information_estimators.py:111-173)."""
    assert len(x_orig) == len(y_orig), "Lists should have same length"
    assert len(x_orig) == len(z_orig), "Lists should have same length"
    x, y, z = _as2d(x_orig), _as2d(y_orig), _as2d(z_orig)
    if normalization:
        x, y, z = _normalise(x, y, z)
    return np.float64(_handle().estimate(x, y, z, _lib.make_opts(_lib.EST_CMI, k)))


def umi(x, y, k=5, density_estimation_method="kde", k_density=5, bw=.01):
    """Uniformised mutual information, Euclidean balls (information_estimators.py:209-277)."""
    assert len(x) == len(y), "Lists should have same length"
    assert k <= len(x) - 1, "Set k smaller than num. samples - 1"
    opts = _lib.make_opts(_lib.EST_UMI, k, density_estimation_method, k_density, bw)
    return float(_handle().estimate(_as2d(x), _as2d(y), None, opts))


def cumi(x_orig, y_orig, z_orig, normalization=False, k=5, density_estimation_method="kde", k_density=5, bw=.01):
    """Uniformised conditional mutual information (information_estimators.py:326-403)."""
    assert len(x_orig) == len(y_orig), "Lists should have same length"
    assert len(x_orig) == len(z_orig), "Lists should have same length"
    x, y, z = _as2d(x_orig), _as2d(y_orig), _as2d(z_orig)
    if normalization:
        x, y, z = _normalise(x, y, z)
    opts = _lib.make_opts(_lib.EST_CUMI, k, density_estimation_method, k_density, bw)
    return np.float64(_handle().estimate(x, y, z, opts))


def ksg_debug(estimator, x, y, z=None, normalization=False, k=5, density_estimation_method="kde", k_density=5,
              bw=.01, path="auto"):
    """Per-point quantities behind an estimate (not in the reference): eps_i, the ball counts and, for the
    uniformised estimators, the weights and weighted sums.  Used by the bit-exactness tests."""
    est = {"mi": _lib.EST_MI, "cmi": _lib.EST_CMI, "umi": _lib.EST_UMI, "cumi": _lib.EST_CUMI}[estimator]
    x, y = _as2d(x), _as2d(y)
    z = None if z is None else _as2d(z)
    if normalization:
        x, y, z = _normalise(x, y, z)
    p = _lib._PATH_NAMES[path]
    opts = _lib.make_opts(est, k, density_estimation_method, k_density, bw, p)
    return _handle().debug_counts(x, y, z, opts)
