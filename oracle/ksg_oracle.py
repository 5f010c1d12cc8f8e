"""CPU oracle for the Scribe KSG hot path -- TEST INFRASTRUCTURE ONLY.

This module is a float64 brute-force restatement (numpy) of the algorithm that
aristoteleo/Scribe-py runs through scipy.spatial.cKDTree / sklearn KernelDensity /
scipy.special.digamma.  It exists to CHECK the CUDA path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it; the product package (scribe_py_b200/) never does and raises if its
CUDA library is missing.

Parity pin: every function here is checked in tests/test_oracle_golden.py against
fixtures under tests/golden/ that were produced by importing the UNMODIFIED
reference from /root/reference (script: tests/golden/make_golden.py).  The
reference ships no golden vectors or working tests of its own (its only test,
Scribe/test/test_psl.py, imports a module that does not exist), so those
reference-generated fixtures are the pin.

Third-party arithmetic restated here (not under /root/reference; versions are
lower-bounded only in the reference's setup.py:11 -- scipy>=1.0, scikit-learn>=0.19.1;
fixtures were generated with scipy 1.18.1 / scikit-learn 1.9.0):
  * cKDTree.query(p=inf / p=2)            -> exact k-th smallest distance, self included
  * cKDTree.query_ball_point(p=inf / p=2) -> closed ball; for p=2 scipy tests sum(diff^2) <= r*r
  * KernelDensity(gaussian).score_samples -> log of an exact Gaussian sum (atol=rtol=0)
  * scipy.special.digamma                 -> psi

Reference lines each function follows are cited as information_estimators.py:LINE (ie:),
causal_network.py:LINE (cn:) and Scribe.py:LINE (sc:).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import digamma as _psi

__all__ = [
    "vd", "mi", "cmi", "umi", "cumi", "knn_counts", "kde_weights", "knn_density_weights",
    "lagged_columns", "rdi", "max_over_delays", "top_incoming", "crdi_jobs", "crdi",
    "dynamics_coupling",
]

_CHUNK = 512


def _as2d(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return a


def vd(d):
    """log-volume of the d-dimensional Euclidean unit ball (ie:15-30)."""
    return 0.5 * d * math.log(math.pi) - math.log(math.gamma(0.5 * d + 1))


def _absdiff_rows(P, r0, r1):
    """|P[i,c]-P[j,c]| for i in [r0,r1), all j, all c -> (r1-r0, N, d) float64."""
    return np.abs(P[r0:r1, None, :] - P[None, :, :])


def knn_counts(cols, subspaces, k=5, metric="chebyshev", weights=None):
    """Brute-force restatement of the tree queries of mi/cmi/cumi/umi.

    cols      : (N, D) joint data.
    subspaces : list of column-index tuples; counts are returned in that order.
    Returns eps (N,), counts (len(subspaces), N) int64 [ball size minus self], and when
    `weights` is given wsum (len(subspaces), N) = sum of weights over the ball MINUS w_i.

    chebyshev : eps_i = sorted_j(max_c |P_ic-P_jc|)[k] (self included; ie:101,:165,:396);
                ball test max_{c in S}|.| <= eps_i (closed; ie:105-107,:169-172,:399-402).
    euclidean : d2 = sum_c diff^2; r_i = sqrt(sorted(d2)[k]); ball test sum_{c in S} diff^2 <= r_i*r_i
                (scipy's p=2 convention; ie:263,:268,:272-273).
    """
    P = _as2d(cols)
    N, D = P.shape
    ns = len(subspaces)
    eps = np.empty(N)
    counts = np.zeros((ns, N), dtype=np.int64)
    wsum = np.zeros((ns, N)) if weights is not None else None
    for r0 in range(0, N, _CHUNK):
        r1 = min(N, r0 + _CHUNK)
        A = _absdiff_rows(P, r0, r1)                     # (m, N, D)
        if metric == "chebyshev":
            joint = A.max(axis=2)
        else:
            A = A * A
            joint = A.sum(axis=2)
        if N > k:
            e = np.partition(joint, k, axis=1)[:, k]
        else:                                            # cKDTree pads missing neighbours with inf
            e = np.full(r1 - r0, np.inf)
        if metric == "euclidean":
            e = np.sqrt(e)
            thr = e * e
        else:
            thr = e
        eps[r0:r1] = e
        for s, S in enumerate(subspaces):
            S = list(S)
            if metric == "chebyshev":
                dS = A[:, :, S].max(axis=2) if len(S) > 1 else A[:, :, S[0]]
            else:
                # scipy accumulates the squared differences coordinate by coordinate
                dS = A[:, :, S[0]].copy()
                for c in S[1:]:
                    dS += A[:, :, c]
            inside = dS <= thr[:, None]
            counts[s, r0:r1] = inside.sum(axis=1) - 1
            if weights is not None:
                wsum[s, r0:r1] = (inside * weights[None, :]).sum(axis=1) - weights[r0:r1]
    if weights is not None:
        return eps, counts, wsum
    return eps, counts


def mi(x, y, use_rank_order=False, k=5):
    """psi(N) + mean_i[psi(n_xy) - psi(n_x) - psi(n_y)]  (ie:59-109)."""
    x = _as2d(x); y = _as2d(y)
    assert len(x) == len(y), "Lists should have same length"
    dx, dy = x.shape[1], y.shape[1]
    P = np.concatenate((x, y), axis=1)
    ix = tuple(range(dx)); iy = tuple(range(dx, dx + dy))
    _, c = knn_counts(P, [ix + iy, ix, iy], k)
    N = len(x)
    return float(np.mean(_psi(N) + _psi(c[0]) - _psi(c[1]) - _psi(c[2])))


def _std_scale(a):
    return a / np.std(a)          # population std over the flattened array (ie:150-153)


def _cmi_layout(x, y, z, normalization):
    x = _as2d(x); y = _as2d(y); z = _as2d(z)
    assert len(x) == len(y), "Lists should have same length"
    assert len(x) == len(z), "Lists should have same length"
    if normalization:
        x = _std_scale(x); y = _std_scale(y); z = _std_scale(z)
    dx, dy, dz = x.shape[1], y.shape[1], z.shape[1]
    P = np.concatenate((x, y, z), axis=1)
    ix = tuple(range(dx)); iy = tuple(range(dx, dx + dy)); iz = tuple(range(dx + dy, dx + dy + dz))
    return P, ix, iy, iz


def cmi(x, y, z, normalization=False, k=5):
    """mean_i[psi(n_xyz) - psi(n_xz) - psi(n_yz) + psi(n_z)]  (ie:111-173)."""
    P, ix, iy, iz = _cmi_layout(x, y, z, normalization)
    _, c = knn_counts(P, [ix + iy + iz, ix + iz, iy + iz, iz], k)
    return float(np.mean(_psi(c[0]) - _psi(c[1]) - _psi(c[2]) + _psi(c[3])))


def cmi_debug(x, y, z, normalization=False, k=5):
    """eps and the four count vectors of cmi (for bit-exactness tests)."""
    P, ix, iy, iz = _cmi_layout(x, y, z, normalization)
    return knn_counts(P, [ix + iy + iz, ix + iz, iy + iz, iz], k)


def kde_weights(P, bw):
    """w = (1/kde)/mean(1/kde) with kde_i proportional to sum_j exp(-|P_i-P_j|^2/(2 bw^2))
    (ie:248-251, :385-388; the Gaussian normalising constant and 1/N cancel)."""
    P = _as2d(P)
    N = len(P)
    S = np.empty(N)
    for r0 in range(0, N, _CHUNK):
        r1 = min(N, r0 + _CHUNK)
        d2 = ((P[r0:r1, None, :] - P[None, :, :]) ** 2).sum(axis=2)
        S[r0:r1] = np.exp(-d2 / (2.0 * bw * bw)).sum(axis=1)
    inv = 1.0 / S
    return inv / np.mean(inv)


def knn_density_weights(P, k_density):
    """w from the k_density-th Chebyshev neighbour distance (ie:254-256, :390-392)."""
    P = _as2d(P)
    N, d = P.shape
    r, _ = knn_counts(P, [tuple(range(d))], k_density)
    with np.errstate(divide="ignore", invalid="ignore"):
        dens = float(k_density) / N / r ** d
        inv = 1.0 / dens
        return inv / np.mean(inv)


def cumi(x, y, z, normalization=False, k=5, density_estimation_method="kde", k_density=5, bw=.01):
    """mean_i w_i[psi(n_xyz) - psi(n_xz) - psi(W_yz - w_i) + psi(W_z - w_i)]  (ie:326-403)."""
    P, ix, iy, iz = _cmi_layout(x, y, z, normalization)
    m = density_estimation_method.lower()
    if m == "kde":
        w = kde_weights(P[:, list(ix + iz)], bw)
    elif m == "knn":
        w = knn_density_weights(P[:, list(ix + iz)], k_density)
    else:
        raise ValueError("The density estimation method is not recognized")
    _, c, W = knn_counts(P, [ix + iy + iz, ix + iz, iy + iz, iz], k, weights=w)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = w * _psi(c[0]) - w * _psi(c[1]) - w * _psi(W[2]) + w * _psi(W[3])
    return float(np.mean(s))


def umi(x, y, k=5, density_estimation_method="kde", k_density=5, bw=.01):
    """Uniformised MI with Euclidean balls (ie:209-277)."""
    x = _as2d(x); y = _as2d(y)
    assert len(x) == len(y), "Lists should have same length"
    assert k <= len(x) - 1, "Set k smaller than num. samples - 1"
    N = len(x)
    dx, dy = x.shape[1], y.shape[1]
    m = density_estimation_method.lower()
    if m == "kde":
        w = kde_weights(x, bw)
    elif m == "knn":
        w = knn_density_weights(x, k_density)
    else:
        raise ValueError("The density estimation method is not recognized")
    P = np.concatenate((x, y), axis=1)
    ix = tuple(range(dx)); iy = tuple(range(dx, dx + dy))
    _, c, W = knn_counts(P, [ix, iy], k, metric="euclidean", weights=w)
    ans = _psi(k) + 2 * math.log(N - 1) - _psi(N) + vd(dx) + vd(dy) - vd(dx + dy)
    for i in range(N):
        ans += -w[i] * math.log(c[0][i]) / N      # math.log(0) raises ValueError like the reference
        ans += -w[i] * math.log(W[1][i]) / N
    return float(ans)


# ----------------------------------------------------------------------------------------
# drivers (causal_network.py / Scribe.py) on plain arrays
# ----------------------------------------------------------------------------------------

def lagged_columns(runs_a, runs_b, delay, differential_mode=False):
    """x, y, z of one RDI job (cn:145-173).  runs_* : list of 1-D arrays (one per run, NaN-free)."""
    xs, ys, zs = [], [], []
    for ea, eb in zip(runs_a, runs_b):
        T = len(eb)
        xs.append(ea[:max(0, len(ea) - delay)])               # e[:-delay]: empty when the run is not longer than the delay
        zs.append(eb[delay - 1:T - 1])
        if differential_mode:
            ys.append(eb[delay - 1:T - 1] - eb[delay:T])      # previous minus current (cn:153)
        else:
            ys.append(eb[delay:T])
    return np.concatenate(xs), np.concatenate(ys), np.concatenate(zs)


def rdi(expr_runs, delays, uniformization=False, differential_mode=False, k=5):
    """expr_runs: list over genes of list over runs of 1-D float64 arrays.
    Returns {delay: GxG array (rows=source, cols=target, diag NaN)} plus "MAX" (cn:112-187)."""
    G = len(expr_runs)
    est = cumi if uniformization else cmi
    out = {}
    for d in delays:
        M = np.full((G, G), np.nan)
        for a in range(G):
            for b in range(G):
                if a == b:
                    continue
                x, y, z = lagged_columns(expr_runs[a], expr_runs[b], d, differential_mode)
                M[a, b] = est(x, y, z, normalization=differential_mode, k=k)
        out[d] = M
    out["MAX"] = max_over_delays(out, delays)[0]
    return out


def max_over_delays(res, delays):
    """(cn:295-311) strict '>' so the first delay in list order wins ties; diagonal stays -inf."""
    G = res[delays[0]].shape[0]
    val = np.full((G, G), -np.inf)
    arg = np.full((G, G), -1, dtype=np.int64)
    for d in delays:
        M = res[d]
        for a in range(G):
            for b in range(G):
                if a != b and M[a, b] > val[a, b]:
                    val[a, b] = M[a, b]
                    arg[a, b] = d
    return val, arg


def top_incoming(max_val, max_delay, L):
    """(cn:316-333) per target the L+1 strongest sources by streaming replace-first-minimum."""
    G = max_val.shape[0]
    nodes = [[None] * (L + 1) for _ in range(G)]
    dls = [[np.nan] * (L + 1) for _ in range(G)]
    vals = [[-np.inf] * (L + 1) for _ in range(G)]
    for b in range(G):
        for a in range(G):
            if a == b:
                continue
            lo = min(vals[b])
            if max_val[a, b] > lo:
                s = vals[b].index(lo)
                nodes[b][s] = a; dls[b][s] = max_delay[a, b]; vals[b][s] = max_val[a, b]
    return nodes, dls, vals


def crdi_jobs(max_val, max_delay, L):
    """Conditioning set per ordered pair (cn:223-235): {(a,b): (delay, [cond genes], [cond delays])}."""
    nodes, dls, vals = top_incoming(max_val, max_delay, L)
    G = max_val.shape[0]
    jobs = {}
    for a in range(G):
        for b in range(G):
            if a == b:
                continue
            cn, cd = list(nodes[b]), list(dls[b])
            drop = nodes[b].index(a) if a in nodes[b] else vals[b].index(min(vals[b]))
            del cn[drop]; del cd[drop]
            jobs[(a, b)] = (int(max_delay[a, b]), cn, [int(v) for v in cd])
    return jobs


def crdi_columns(expr_runs, a, b, delay, cond, cond_delays, differential=False):
    """x, y, z of one cRDI job following the multi-run branch (cn:257-283); differential: y_t = e_b[tau-1+t] - e_b[tau+t] (cn:271)."""
    L = len(cond)
    tau = max(list(cond_delays) + [delay])
    xs, ys, zs = [], [], []
    for r in range(len(expr_runs[b])):
        eb = expr_runs[b][r]; ea = expr_runs[a][r]
        n = len(eb) - tau
        xs.append(ea[tau - delay:tau - delay + n])
        ys.append(eb[tau - 1:tau - 1 + n] - eb[tau:tau + n] if differential else eb[tau:tau + n])
        zc = [eb[tau - 1:tau - 1 + n]]
        for i in range(L):                                  # each new column is PREPENDED (cn:277)
            e3 = expr_runs[cond[i]][r]
            zc.insert(0, e3[tau - cond_delays[i]:tau - cond_delays[i] + n])
        zs.append(np.stack(zc, axis=1))
    return np.concatenate(xs), np.concatenate(ys), np.concatenate(zs, axis=0)


def crdi(expr_runs, rdi_res, delays, L=1, uniformization=False, k=5, differential_mode=False):
    """differential_mode: y differenced and normalization=True, i.e. every role divided by np.std of its whole array (cn:219-221)."""
    G = len(expr_runs)
    val, arg = max_over_delays(rdi_res, delays)
    jobs = crdi_jobs(val, arg, L)
    est = cumi if uniformization else cmi
    M = np.full((G, G), np.nan)
    for (a, b), (d, cn, cd) in jobs.items():
        x, y, z = crdi_columns(expr_runs, a, b, d, cn, cd, differential_mode)
        M[a, b] = est(x, y, z, k=k, normalization=True) if differential_mode else est(x, y, z, k=k)
    return M


def dynamics_coupling(spliced, velocity, tf_idx=None, target_idx=None, y_is_sum=True, normalize=True,
                      drop_zero_cells=True, k=5, return_counts=False):
    """sc:104-137 on plain (cells x genes) float64 arrays. Returns (len(tf) x len(target)) matrix,
    skipped a==b entries 0."""
    s = np.array(spliced, dtype=np.float64)
    v = np.array(velocity, dtype=np.float64)
    v[np.isnan(v)] = 0.0
    if normalize:
        s = (s - s.mean(axis=0)) / (s.max(axis=0) - s.min(axis=0))
        v = (v - v.mean(axis=0)) / (v.max(axis=0) - v.min(axis=0))
    G = s.shape[1]
    tf_idx = list(range(G)) if tf_idx is None else list(tf_idx)
    target_idx = list(range(G)) if target_idx is None else list(target_idx)
    M = np.zeros((len(tf_idx), len(target_idx)))
    neff = np.zeros((len(tf_idx), len(target_idx)), dtype=np.int64)
    for ia, a in enumerate(tf_idx):
        for ib, b in enumerate(target_idx):
            if a == b:
                continue
            x = s[:, a]; z = s[:, b]
            y = (s[:, b] + v[:, b]) if y_is_sum else v[:, b]
            if drop_zero_cells:
                keep = ((x + y) + z) > 0
                x, y, z = x[keep], y[keep], z[keep]
            neff[ia, ib] = len(x)
            M[ia, ib] = cmi(x, y, z, k=k)
    return (M, neff) if return_counts else M
