"""Reference-shaped CPU port of the KSG estimators -- TEST / BASELINE INFRASTRUCTURE ONLY.

Where ksg_oracle.py restates the *semantics* by brute force, this file restates the reference's *cost
structure*: the same third-party calls (scipy cKDTree build + one `query` and three or four
`query_ball_point` calls per sample, sklearn KernelDensity, scalar digamma) issued from per-sample Python
loops, as information_estimators.py:160-173, :243-277, :379-403 do.  bench.py times it on the GPU box's host
cores as `cpu_baseline` / `--impl reference` (kind "port": /root/reference is not on that box and the
reference is pure Python, so there is nothing to compile into oracle/_ref).  Checked against the
reference-generated fixtures in tests/test_oracle_golden.py.  Never imported by scribe_py_b200.
"""
from math import log

import numpy as np
from scipy.spatial import cKDTree
from scipy.special import digamma

from .ksg_oracle import vd, lagged_columns


def _cols(a):
    a = np.asarray(a, dtype=float)
    return a[:, None] if a.ndim == 1 else a


def _ball(tree, pts, i, r, p=np.inf):
    return len(tree.query_ball_point(pts[i], r, p=p)) - 1


def cmi(x, y, z, normalization=False, k=5):
    x, y, z = _cols(x), _cols(y), _cols(z)
    if normalization:
        x = x / np.std(x); y = y / np.std(y); z = z / np.std(z)
    xyz, xz, yz = np.hstack((x, y, z)), np.hstack((x, z)), np.hstack((y, z))
    t_xyz, t_xz, t_yz, t_z = cKDTree(xyz), cKDTree(xz), cKDTree(yz), cKDTree(z)
    n = len(xyz)
    total = 0.0
    terms = []
    for i in range(n):
        r = t_xyz.query(xyz[i], k + 1, p=np.inf)[0][k]
        terms.append(digamma(_ball(t_xyz, xyz, i, r)) - digamma(_ball(t_xz, xz, i, r))
                     - digamma(_ball(t_yz, yz, i, r)) + digamma(_ball(t_z, z, i, r)))
    return float(np.mean(terms))


def mi(x, y, k=5):
    x, y = _cols(x), _cols(y)
    xy = np.hstack((x, y))
    t_xy, t_x, t_y = cKDTree(xy), cKDTree(x), cKDTree(y)
    r = t_xy.query(xy, k + 1, p=np.inf)[0][:, k]
    n = len(xy)
    terms = [digamma(n) + digamma(_ball(t_xy, xy, i, r[i])) - digamma(_ball(t_x, x, i, r[i]))
             - digamma(_ball(t_y, y, i, r[i])) for i in range(n)]
    return float(np.mean(terms))


def _weights(P, tree, method, k_density, bw):
    from sklearn.neighbors import KernelDensity
    m = method.lower()
    if m == "kde":
        dens = np.exp(KernelDensity(bandwidth=bw).fit(P).score_samples(P))
    elif m == "knn":
        r = np.array([tree.query(p, k_density + 1, p=np.inf)[0][k_density] for p in P])
        with np.errstate(divide="ignore"):
            dens = float(k_density) / len(P) / r ** P.shape[1]
    else:
        raise ValueError("The density estimation method is not recognized")
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dens
        return inv / np.mean(inv)


def cumi(x, y, z, normalization=False, k=5, density_estimation_method="kde", k_density=5, bw=.01):
    x, y, z = _cols(x), _cols(y), _cols(z)
    if normalization:
        x = x / np.std(x); y = y / np.std(y); z = z / np.std(z)
    xyz, xz, yz = np.hstack((x, y, z)), np.hstack((x, z)), np.hstack((y, z))
    t_xyz, t_xz, t_yz, t_z = cKDTree(xyz), cKDTree(xz), cKDTree(yz), cKDTree(z)
    w = _weights(xz, t_xz, density_estimation_method, k_density, bw)
    terms = []
    for i in range(len(xyz)):
        r = t_xyz.query(xyz[i], k + 1, p=np.inf)[0][k]
        w_yz = sum(w[j] for j in t_yz.query_ball_point(yz[i], r, p=np.inf)) - w[i]
        w_z = sum(w[j] for j in t_z.query_ball_point(z[i], r, p=np.inf)) - w[i]
        with np.errstate(invalid="ignore", divide="ignore"):
            terms.append(w[i] * (digamma(_ball(t_xyz, xyz, i, r)) - digamma(_ball(t_xz, xz, i, r))
                                 - digamma(w_yz) + digamma(w_z)))
    return float(np.mean(terms))


def umi(x, y, k=5, density_estimation_method="kde", k_density=5, bw=.01):
    x, y = _cols(x), _cols(y)
    n = len(x)
    assert k <= n - 1, "Set k smaller than num. samples - 1"
    xy = np.hstack((x, y))
    t_xy, t_x, t_y = cKDTree(xy), cKDTree(x), cKDTree(y)
    w = _weights(x, t_x, density_estimation_method, k_density, bw)
    dx, dy = x.shape[1], y.shape[1]
    ans = digamma(k) + 2 * log(n - 1) - digamma(n) + vd(dx) + vd(dy) - vd(dx + dy)
    for i in range(n):
        r = t_xy.query(xy[i], k + 1, p=2)[0][k]
        nx = _ball(t_x, x, i, r, p=2)
        wy = sum(w[j] for j in t_y.query_ball_point(y[i], r, p=2)) - w[i]
        ans -= w[i] * log(nx) / n
        ans -= w[i] * log(wy) / n
    return float(ans)


def rdi_job(args):
    """Module-level worker for multiprocessing: one (source, target, delay) evaluation."""
    ea, eb, delay, uniformization = args
    x, y, z = lagged_columns([ea], [eb], delay)
    return cumi(x, y, z) if uniformization else cmi(x, y, z)
