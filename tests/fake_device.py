"""Test double for scribe_py_b200._lib.Handle: evaluates submitted jobs with the CPU oracle so that the HOST
logic of the drop-in (job tables, sharding, frames, MAX / top-incoming / conditioning sets, coupling prep)
can be checked without a GPU.  Lives under tests/ -- the product never imports it."""
import numpy as np

from oracle import ksg_oracle as o
from scribe_py_b200 import _lib


class FakeHandle:
    def __init__(self):
        self.calls = []

    def upload_matrix(self, E, run_offsets=None):
        self.E = np.asarray(E, dtype=float)
        self.ro = np.array([0, self.E.shape[1]]) if run_offsets is None else np.asarray(run_offsets)

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
                                         [int(v) for v in jb["cond_gene"][:L]], [int(v) for v in jb["cond_delay"][:L]])
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
