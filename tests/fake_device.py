"""Test double for scribe_py_b200._lib.Handle: evaluates submitted jobs with the CPU oracle so that the HOST
logic of the drop-in (job tables, sharding, frames, MAX / top-incoming / conditioning sets, coupling prep)
can be checked without a GPU.  Lives under tests/ -- the product never imports it."""
import numpy as np

from oracle import ksg_oracle as o
from scribe_py_b200 import _lib


class FakeHandle:
    def __init__(self):
        self.calls = []

    resident_token = 0

    def upload_matrix(self, E, run_offsets=None):
        self.E = np.asarray(E, dtype=float)
        self.ro = np.array([0, self.E.shape[1]]) if run_offsets is None else np.asarray(run_offsets)
        self.uploads = getattr(self, "uploads", 0) + 1
        self.resident_token += 1
        return self.resident_token

    def read_matrix(self, gene0=0, n_genes=None):
        return self.E[gene0:None if n_genes is None else gene0 + n_genes].copy()

    def _runs(self, g):
        return [self.E[g, self.ro[r]:self.ro[r + 1]] for r in range(len(self.ro) - 1)]

    def lagged_jobs(self, jobs, opts, differential=False, scales=None, out_dev_ptr=None):
        est = o.cumi if opts.estimator == _lib.EST_CUMI else o.cmi
        out = np.empty(len(jobs))
        self.calls.append(len(jobs))
        expr = [self._runs(g) for g in range(self.E.shape[0])]
        for j, jb in enumerate(jobs):
            L = int(jb["n_cond"])
            if L == 0:
                x, y, z = o.lagged_columns(expr[jb["src"]], expr[jb["dst"]], int(jb["delay"]), differential)
            else:
                x, y, z = o.crdi_columns(expr, int(jb["src"]), int(jb["dst"]), int(jb["delay"]),
                                         [int(v) for v in jb["cond_gene"][:L]], [int(v) for v in jb["cond_delay"][:L]], differential)
            if scales is not None:
                x, y, z = x / scales[j, 0], y / scales[j, 1], z / scales[j, 2]
            out[j] = est(x, y, z)
        return out

    def mi_jobs(self, src, dst, opts):
        return np.array([o.mi(self.E[a], self.E[b]) for a, b in zip(src, dst)])

    def upload_coupling(self, S, Y):
        self.S, self.Y = np.asarray(S, dtype=float), np.asarray(Y, dtype=float)

    def coupling_jobs(self, src, dst, opts, drop_zero_cells=True, out_dev_ptr=None, want_n_eff=False):
        out = np.empty(len(src))
        for j, (a, b) in enumerate(zip(src, dst)):
            x, y, z = self.S[a], self.Y[b], self.S[b]
            if drop_zero_cells:
                keep = ((x + y) + z) > 0
                x, y, z = x[keep], y[keep], z[keep]
            out[j] = o.cmi(x, y, z)
        return out

    # ---- consumers of the causal matrix: numpy restatements of Scribe.py:142-176, plot/networks.py:34-47, causal_network.py:337-389
    def clr(self, M, both_dims):
        mat = np.asarray(M, dtype=float)
        sq = mat.shape[0] == mat.shape[1]

        def z(axis):
            v = np.round(mat, 10)
            mu, sd = np.mean(mat, axis=axis), np.std(mat, axis=axis, ddof=1)
            if not sq:
                mu = mu[:, None] if axis == 1 else mu[None, :]
                sd = sd[:, None] if axis == 1 else sd[None, :]
            v = (v - mu) / sd
            v[v < 0] = 0
            return v

        zr = z(1)
        return np.sqrt((zr ** 2 + z(0) ** 2) / 2) if both_dims else np.sqrt(zr ** 2)

    def top_edges(self, M, n_top):
        M = np.asarray(M, dtype=float).copy()
        M[(M - M.T) < 0] = np.nan
        i, j = np.nonzero(~np.isnan(M))
        order = np.argsort(-M[i, j], kind="stable")[:n_top]
        return i[order].astype(np.int32), j[order].astype(np.int32), M[i, j][order]

    def roc(self, scores, truth):
        n = scores.shape[0]
        i, j = np.nonzero(~np.eye(n, dtype=bool))
        order = np.argsort(-scores[i, j], kind="stable")
        hit = truth[i, j][order].astype(bool)
        tp = np.concatenate(([0], np.cumsum(hit))); fp = np.concatenate(([0], np.cumsum(~hit)))
        if tp[-1] == 0 or fp[-1] == 0:
            raise ZeroDivisionError("float division by zero")
        x = fp / float(fp[-1]); y = tp / float(tp[-1])
        a = 0
        for t in range(1, len(x)):
            a += abs((x[t] - x[t - 1]) * (y[t] + y[t - 1]) / 2)
        return a, x, y
