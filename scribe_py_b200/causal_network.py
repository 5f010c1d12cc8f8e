"""Drop-in for Scribe/causal_network.py: the `causal_model` class.

`rdi` and `crdi` keep the reference's signatures and return types (causal_network.py:112-187, :190-289) but
instead of calling cmi/cumi once per ordered pair and delay they pack every (source, target, delay) job of
the call into one batched submission to the CUDA library (one upload of the expression matrix, lagged
columns materialised once per (gene, shift) and shared by all pairs that read them), sharded across the
GPUs of the box when torch.distributed is initialised.

Behaviours of the reference that crash on a current numpy/pandas stack (np.int at :299, the single-run
cRDI concatenation at :246, the Pool branch at :158/:181) are made to work here with the semantics of the
branches that do run (multi-run cRDI, :257-283).
"""
from __future__ import annotations

import warnings
from copy import deepcopy

import numpy as np
import pandas as pd

from . import _lib
from . import distributed as _dist

__all__ = ["causal_model"]


class _PairDelayJobs:
    """Job table of rdi(): every ordered pair x delay, delay-major.  Only the slices a rank asks for are materialised
    (`jobs[lo:hi]` -> JOB_DTYPE array), so the per-rank host work of a sharded call is proportional to its slab."""

    def __init__(self, src, tgt, delays):
        self.src, self.tgt, self.delays = src, tgt, np.asarray(list(delays), dtype=np.int32)
        self.per = len(src)
        self._last = None                                # (lo, hi, table): a repeated request (same slab every step) is free

    def __len__(self):
        return self.per * len(self.delays)

    def __getitem__(self, key):
        if not isinstance(key, slice):
            raise TypeError("job tables are sliced, not indexed")
        lo, hi, step = key.indices(len(self))
        assert step == 1
        if self._last is not None and self._last[0] == lo and self._last[1] == hi:
            return self._last[2]
        jobs = np.zeros(max(hi - lo, 0), dtype=_lib.JOB_DTYPE)
        per = max(self.per, 1)
        for i in range(lo // per, len(self.delays)):                 # one contiguous copy per delay the slab touches
            a, b = max(lo, i * per), min(hi, (i + 1) * per)
            if a >= b:
                break
            o = slice(a - lo, b - lo)
            jobs["src"][o] = self.src[a - i * per:b - i * per]; jobs["dst"][o] = self.tgt[a - i * per:b - i * per]
            jobs["delay"][o] = self.delays[i]
        self._last = (lo, hi, jobs)
        return jobs


class causal_model(object):
    """Holds gene expression (genes x cells, single-run index GENE_ID or MultiIndex (GENE_ID, RUN_ID)) and
    computes pairwise RDI / cRDI scores on the GPU."""

    # ------------------------------------------------------------------ expression: host frame <-> matrix resident on the GPU
    # `expression` is the reference's public attribute (causal_network.py:15-23).  Here it is a property over two copies of the
    # same data: the host DataFrame and the dense matrix resident on the device.  rdi()/crdi() upload the matrix once and reuse it
    # while the frame is unchanged; normalize()/smooth() transform the resident matrix in place and the frame is rebuilt from the
    # device only when somebody reads it.  Assigning a frame invalidates the device copy.  (Mutating the assigned frame in place
    # is not seen: call invalidate_device() after doing that.)
    @property
    def expression(self):
        if getattr(self, "_expr_stale", False):
            self._materialise_expression()
        try:
            return self._expr
        except AttributeError:
            raise AttributeError("'causal_model' object has no attribute 'expression'") from None

    @expression.setter
    def expression(self, frame):
        self._expr = frame
        self._expr_stale = False
        self._dev = None

    def invalidate_device(self):
        """Forget the matrix resident on the GPU (after an in-place edit of model.expression)."""
        if getattr(self, "_expr_stale", False):
            self._materialise_expression()
        self._dev = None

    def __init__(self, expression=None):
        self._dev = None
        self._expr_stale = False
        if expression is None:
            warnings.warn(" WARNING: No expression data argument given. if you intend to load the data from a file, "
                          "call the method 'read_expression_file'.")
        else:
            self.expression = deepcopy(expression)
            self.expression_raw = expression
            if isinstance(expression, pd.DataFrame):
                if expression.index.nlevels == 1:
                    self.node_ids = expression.index
                else:
                    self.node_ids = expression.index.levels[0]
                    self.run_ids = expression.index.levels[1]
        self.rdi_results = None
        self.crdi_results = None

    # ------------------------------------------------------------------ I/O and preprocessing (pass-through)
    def read_expression_file(self, expression_path, data_mode, verbose=False, genotype_path=None, phenotype_path=None):
        """Tab-separated expression table -> DataFrame (causal_network.py:28-65)."""
        if verbose:
            print("\nLoading data... ",)
        if data_mode == "single_run":
            self.expression = pd.read_table(expression_path, delimiter="\t", header=0, index_col="GENE_ID")
            self.expression_raw = deepcopy(self.expression)
            self.node_ids = self.expression.index
            self.expression_concatenated = self.expression
        elif data_mode == "multi_run":
            self.expression = pd.read_table(expression_path, delimiter="\t", header=0, index_col=["GENE_ID", "RUN_ID"])
            self.expression_raw = deepcopy(self.expression)
            self.node_ids = self.expression.index.levels[0]
            self.run_ids = self.expression.index.levels[1]
            self.expression_concatenated = pd.concat(
                [pd.DataFrame(self.expression.loc[g].values.reshape(1, -1), index=[g]) for g in self.node_ids])
        else:
            raise ValueError("Data mode invalid")
        if verbose:
            print("DONE.")
            if data_mode == "single_run":
                print("A total number of", self.expression.shape[0], "genes, each containing",
                      self.expression.shape[1], "pseudo-times were read.\n")
            else:
                print("A total number of", len(self.expression.index.levels[0]), "genes, each gene containing",
                      len(self.expression.index.levels[1]), "runs and each run containing at most",
                      self.expression.shape[1], "pseudo-times were read.\n")

    def restore_the_raw_expression_data(self):
        self.expression = deepcopy(self.expression_raw)

    def subset_genes(self, subset):
        if self.expression.index.nlevels == 1:
            self.expression = self.expression.loc[subset, :]
        else:
            keep = [ix for ix in self.expression.index if ix[0] in subset]
            self.expression = self.expression.loc[keep, :]
        self.node_ids = subset

    def subset_runs(self, subset):
        if self.expression.index.nlevels == 1:
            warnings.warn("WARNING: Cannot perform run-subsetting on a single-run mode data.")
        else:
            keep = [ix for ix in self.expression.index if ix[1] in subset]
            self.expression = self.expression.loc[keep, :]
            self.expression_concatenated = pd.concat(
                [pd.DataFrame(self.expression.loc[g].values.reshape(1, -1), index=[g]) for g in self.node_ids])
            self.run_ids = subset

    def smooth(self, window):
        """Rolling mean of expression_raw along the cell axis, all-NaN columns dropped (causal_network.py:97-101; the reference's
        axis= kwarg no longer exists in pandas, `.T.rolling(window).mean().T` is the same computation).  Runs on the GPU when
        the raw matrix has no NaN padding and the device result reproduces pandas on probe rows; otherwise in pandas."""
        if self._smooth_on_device(window):
            return
        sm = self.expression_raw.T.rolling(window=window).mean().T
        self.expression = sm.dropna(axis=1, how="all")

    def normalize(self):
        """z-score every row, NaN -> 0 (causal_network.py:104-109).  Runs in place on the matrix resident on the GPU (uploading
        it first if necessary) when that reproduces pandas bit for bit on probe rows; otherwise in pandas."""
        if self._normalize_on_device():
            return
        e = self.expression
        e = e.sub(e.mean(axis=1), axis=0)
        e = e.div(e.std(axis=1), axis=0)                          # std of the centred frame, as the reference's two statements do
        self.expression = e.fillna(0)

    # ------------------------------------------------------------------ device-side preprocessing
    @staticmethod
    def _ordered_sum(v, order):
        if order == 0:
            return float(np.add.reduce(np.ascontiguousarray(v)))      # numpy pairwise
        s = 0.0
        for x in v:
            s += x
        return s

    def _probe_rows(self, n_rows):
        return sorted(set([0, n_rows // 2, n_rows - 1]))

    def _normalize_on_device(self):
        h = _lib.default_handle()
        if not hasattr(h, "normalize_resident"):
            return False
        try:
            node_ids, run_off, _ = self._ensure_resident()
        except AssertionError:
            return False
        lay = self._dev
        if lay is None or lay["padded"]:
            return False                                           # NaN padding: pandas sums the padded rows, not reproduced here
        # which summation orders does this pandas use?  (3 probe rows, host; the stale frame is never needed: probe from the device)
        n_runs = len(run_off) - 1
        probe = self._probe_rows(len(node_ids))
        rows = np.stack([h.read_matrix(g, 1)[0] for g in probe])
        want = []
        for r in range(n_runs):
            # pandas on a frame of the probe rows (several rows on purpose: the summation order of DataFrame.mean(axis=1)
            # depends on the frame's memory layout, and a one-row frame is laid out differently from a many-row one)
            f = pd.DataFrame(rows[:, run_off[r]:run_off[r + 1]].copy())
            c = f.sub(f.mean(axis=1), axis=0)
            z = c.div(c.std(axis=1), axis=0).fillna(0).to_numpy()
            for i in range(len(probe)):
                want.append((rows[i, run_off[r]:run_off[r + 1]], z[i]))
        pick = None
        for mo in (1, 0):
            for ao in (0, 1):
                ok = True
                for seg, z in want:
                    n = len(seg)
                    c = seg - self._ordered_sum(seg, mo) / n
                    avg = self._ordered_sum(c, ao) / n
                    with np.errstate(invalid="ignore", divide="ignore"):
                        sd = np.sqrt(np.add.reduce((avg - c) * (avg - c)) / (n - 1))
                        got = c / sd
                    got[np.isnan(got)] = 0
                    if not np.array_equal(got, z):
                        ok = False
                        break
                if ok:
                    pick = (mo, ao)
                    break
            if pick:
                break
        if pick is None:
            return False
        had_host_copy = not self._expr_stale
        h.normalize_resident(*pick)
        after = np.stack([h.read_matrix(g, 1)[0] for g in probe])      # the device must reproduce pandas on the probe rows
        for r in range(n_runs):
            for i in range(len(probe)):
                if not np.array_equal(after[i, run_off[r]:run_off[r + 1]], want[r * len(probe) + i][1]):
                    self._dev = None                               # the resident copy is no longer this model's frame
                    if had_host_copy:
                        return False                              # pandas redoes it on the host frame
                    raise RuntimeError("device normalisation differs from pandas on a probe row and no host copy is left; "
                                       "assign model.expression again")
        self._expr_stale = True
        return True

    def _smooth_on_device(self, window):
        h = _lib.default_handle()
        if not hasattr(h, "smooth_resident"):
            return False
        raw = self.expression_raw
        if not isinstance(raw, pd.DataFrame) or window < 1:
            return False
        if getattr(self, "_expr_stale", False):
            self._materialise_expression()                         # the current frame must survive a declined attempt
        keep_expr, keep_dev = getattr(self, "_expr", None), self._dev
        ok = False
        try:
            self._expr, self._expr_stale, self._dev = raw, False, None     # make the RAW matrix the resident one
            node_ids, run_off, _ = self._ensure_resident()
            lay = self._dev
            ok = lay is not None and not lay["padded"] and min(np.diff(run_off)) >= window
        except AssertionError:
            ok = False
        finally:
            if not ok:                                             # declined (or failed): the model is as it was, minus the device copy
                self._expr, self._expr_stale, self._dev = keep_expr, False, None
        if not ok:
            return False
        n_runs = len(run_off) - 1
        probe = self._probe_rows(len(node_ids))
        before = {g: h.read_matrix(g, 1)[0] for g in probe}
        new_off = h.smooth_resident(window, n_runs)
        for g in probe:
            v = h.read_matrix(g, 1)[0]
            for r in range(n_runs):
                seg = before[g][run_off[r]:run_off[r + 1]]
                want = pd.Series(seg).rolling(window=window).mean().to_numpy()[window - 1:]
                if not np.array_equal(v[new_off[r]:new_off[r + 1]], want, equal_nan=True):
                    self._expr, self._expr_stale, self._dev = keep_expr, False, None
                    return False
        lay = dict(lay)
        lay["run_off"] = np.asarray(new_off, dtype=np.int64)
        lay["columns"] = [c[window - 1:] for c in lay["columns"]]
        lay["frame_columns"] = lay["frame_columns"][window - 1:]
        lay["n_cells"] = int(new_off[-1])
        self._dev = lay
        self._expr_stale = True
        return True

    def _materialise_expression(self):
        """Rebuild the host frame from the matrix resident on the device (after normalize()/smooth() ran there)."""
        lay = self._dev
        h = _lib.default_handle()
        if lay is None or getattr(h, "resident_token", None) != lay["token"]:
            raise RuntimeError("the device copy of model.expression was replaced before the frame was read back")
        E = h.read_matrix()
        ro = lay["run_off"]
        if lay["runs"] == [None]:
            frame = pd.DataFrame(E, index=lay["index"], columns=lay["columns"][0])
        else:
            width = max(len(c) for c in lay["columns"])
            rows = np.full((len(lay["index"]), width), np.nan)
            pos = {key: i for i, key in enumerate(lay["index"])}
            for g_i, g in enumerate(lay["node_ids"]):
                for r_i, run in enumerate(lay["runs"]):
                    rows[pos[(g, run)], :ro[r_i + 1] - ro[r_i]] = E[g_i, ro[r_i]:ro[r_i + 1]]
            labels = lay["frame_columns"]
            frame = pd.DataFrame(rows, index=lay["index_obj"], columns=labels[:width] if len(labels) >= width else None)
        self._expr = frame
        self._expr_stale = False

    @staticmethod
    def _fingerprint(frame):
        v = frame.to_numpy()
        flat = v.reshape(-1) if v.flags["C_CONTIGUOUS"] else v.T.reshape(-1)
        step = max(1, len(flat) // 4096)
        return (id(frame), v.shape, hash(flat[::step].tobytes()))

    def _ensure_resident(self):
        """(node_ids, run_offsets, E or None): makes sure the dense matrix of the current expression is resident on this
        rank's GPU.  E is None when the device already holds it (nothing was uploaded)."""
        h = _lib.default_handle()
        lay = self._dev
        if lay is not None and getattr(h, "resident_token", None) == lay["token"] and \
                (self._expr_stale or lay["fingerprint"] == self._fingerprint(self._expr)):
            return lay["node_ids"], lay["run_off"], None
        node_ids, E, run_off = self._dense()
        token = h.upload_matrix(E, run_off)
        expr = self._expr
        runs = [None] if expr.index.nlevels == 1 else list(self.run_ids)
        padded = bool(np.isnan(expr.to_numpy(dtype=np.float64)).any())
        if runs == [None]:
            columns = [list(expr.columns)]
        else:
            columns = [list(range(int(run_off[r + 1] - run_off[r]))) for r in range(len(runs))]
        self._dev = None if token is None else {
            "token": token, "fingerprint": self._fingerprint(expr), "node_ids": node_ids, "run_off": np.asarray(run_off, dtype=np.int64),
            "runs": runs, "padded": padded, "index": expr.index if runs == [None] else list(expr.index), "index_obj": expr.index,
            "frame_columns": list(expr.columns),
            "columns": columns,
            "n_cells": int(E.shape[1]), "E_host": None}
        if runs == [None] and not (len(expr.index) == len(node_ids) and list(expr.index) == node_ids):
            self._dev = None                                   # reordered / subset rows: no write-back layout, upload every time
        return node_ids, run_off, E

    # ------------------------------------------------------------------ dense layout for the device
    def _dense(self):
        """(node_ids list, E genes x cells float64 with NaN dropped, run_offsets int64).

        Mirrors `.loc[gene(, run)].dropna()` (causal_network.py:147,:167): NaNs are removed per gene and run;
        all genes must then agree on every run length, otherwise the reference's cmi would fail its
        "Lists should have same length" assertion."""
        node_ids = list(self.node_ids)
        expr = self.expression
        if expr.index.nlevels == 1:
            runs = [None]
        else:
            runs = list(self.run_ids)
        blocks, lengths = [], []
        for run in runs:
            if run is None:
                aligned = len(expr.index) == len(node_ids) and list(expr.index) == node_ids
                vals = (expr if aligned else expr.loc[node_ids]).to_numpy(dtype=np.float64)
            else:
                vals = expr.loc[[(g, run) for g in node_ids]].to_numpy(dtype=np.float64)
            bad = np.isnan(vals)
            if not bad.any():                                    # common case: no padding at all
                blocks.append(vals); lengths.append(vals.shape[1])
                continue
            ok = ~bad
            cnt = ok.sum(axis=1)
            assert (cnt == cnt[0]).all(), "Lists should have same length"
            T = int(cnt[0])
            if ok.all():
                blk = vals
            elif (ok == ok[0]).all():
                blk = vals[:, ok[0]]
            else:
                blk = np.stack([vals[g][ok[g]] for g in range(len(node_ids))])
            blocks.append(blk)
            lengths.append(T)
        E = np.ascontiguousarray(blocks[0] if len(blocks) == 1 else np.concatenate(blocks, axis=1))
        run_off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
        return node_ids, E, run_off

    @staticmethod
    def _lag_std(E, run_off, delays):
        """np.std of the x / y(differential) / z columns per (gene, delay) for normalization=True
        (information_estimators.py:150-153 applied to the slices of causal_network.py:147-153)."""
        G = E.shape[0]
        out = {}
        for d in delays:
            sx = np.empty(G); sy = np.empty(G); sz = np.empty(G)
            for g in range(G):
                xs, ys, zs = [], [], []
                for r in range(len(run_off) - 1):
                    e = E[g, run_off[r]:run_off[r + 1]]
                    T = len(e)
                    if T <= d:
                        continue
                    xs.append(e[:T - d]); zs.append(e[d - 1:T - 1]); ys.append(e[d - 1:T - 1] - e[d:T])
                sx[g] = np.std(np.concatenate(xs)[:, None]); sy[g] = np.std(np.concatenate(ys)[:, None])
                sz[g] = np.std(np.concatenate(zs)[:, None])
            out[d] = (sx, sy, sz)
        return out

    @staticmethod
    def _crdi_scales(E, run_off, jobs, L):
        """Per-job np.std of the x, y and z arrays of differential-mode cRDI (cmi/cumi(..., normalization=True),
        causal_network.py:219-221 -> information_estimators.py:150-153): x = e_a[tau-d : tau-d+n], y = e_b[tau-1+t] - e_b[tau+t]
        (:271), z = the whole (N, L+1) block [cond[L-1] | ... | cond[0] | e_b[tau-1 : tau-1+n]] (:267-278), runs concatenated.
        The slices of a job depend on few keys ((gene, offset, tau) / (target, tau) / (target, tau, conditioning set)), so the
        standard deviations are cached by key: every source of a target that is outside its top set shares them."""
        nr = len(run_off) - 1
        runs = [(int(run_off[r]), int(run_off[r + 1])) for r in range(nr)]

        def col(g, shift, tau):
            return np.concatenate([E[g, a + shift:b - tau + shift] for a, b in runs if b - a > tau])

        cx, cy, cz = {}, {}, {}
        out = np.empty((len(jobs), 3))
        for j in range(len(jobs)):
            jb = jobs[j]
            a, b, d = int(jb["src"]), int(jb["dst"]), int(jb["delay"])
            cg = [int(v) for v in jb["cond_gene"][:L]]; cd = [int(v) for v in jb["cond_delay"][:L]]
            tau = max([d] + cd)
            kx, ky, kz = (a, tau - d, tau), (b, tau), (b, tau, tuple(cg), tuple(cd))
            if kx not in cx:
                cx[kx] = np.std(col(a, tau - d, tau)[:, None])
            if ky not in cy:
                cy[ky] = np.std((col(b, tau - 1, tau) - col(b, tau, tau))[:, None])
            if kz not in cz:
                cols = [col(cg[i], tau - cd[i], tau) for i in range(L - 1, -1, -1)] + [col(b, tau - 1, tau)]
                cz[kz] = np.std(np.stack(cols, axis=1))
            out[j, 0], out[j, 1], out[j, 2] = cx[kx], cy[ky], cz[kz]
        return out

    def _frame(self, M, node_ids):
        ix = getattr(self, "_ix_cache", None)
        if ix is None or len(ix[0]) != len(node_ids) or ix[0] != node_ids:      # build the label Index once per gene list
            ix = (list(node_ids), pd.Index(node_ids))
            self._ix_cache = ix
        return pd.DataFrame(M, index=ix[1], columns=ix[1], copy=False)

    def _rdi_plan(self, delays, differential_mode=False):
        """Host-side job table of one rdi() call: every ordered pair x delay, ordered delay-major, then
        target, then source, so that a rank's contiguous slab shares its targets' y/z columns."""
        delays = list(delays)
        node_ids, run_off, E = self._ensure_resident()              # uploads the matrix unless the device already holds it
        if E is None and differential_mode:
            E = _lib.default_handle().read_matrix()
        G = len(node_ids)
        tgt, src = np.meshgrid(np.arange(G, dtype=np.int32), np.arange(G, dtype=np.int32), indexing="ij")
        off = tgt != src
        tgt, src = tgt[off], src[off]
        per = len(tgt)
        jobs = _PairDelayJobs(src, tgt, delays)
        scales = None
        if differential_mode:
            st = self._lag_std(E, run_off, delays)
            scales = np.empty((len(jobs), 3))
            for i, d in enumerate(delays):
                sl = slice(i * per, (i + 1) * per)
                sx, sy, sz = st[d]
                scales[sl, 0] = sx[src]; scales[sl, 1] = sy[tgt]; scales[sl, 2] = sz[tgt]
        return {"delays": delays, "node_ids": node_ids, "E": E, "run_off": run_off, "jobs": jobs, "scales": scales,
                "src": src, "tgt": tgt, "per": per}

    @staticmethod
    def _rdi_execute(plan, uniformization=False, differential_mode=False):
        """Evaluate a plan against the matrix resident on this rank's GPU; all ranks get all results."""
        h = _lib.default_handle()
        jobs, scales = plan["jobs"], plan["scales"]
        opts = _lib.make_opts(_lib.EST_CUMI if uniformization else _lib.EST_CMI, 5)

        def evaluate(lo, hi, out_dev_ptr=None):
            return h.lagged_jobs(jobs[lo:hi], opts, differential_mode, None if scales is None else scales[lo:hi],
                                 out_dev_ptr=out_dev_ptr)

        return _dist.sharded_evaluate(len(jobs), evaluate, h)

    # ------------------------------------------------------------------ RDI
    def rdi(self, delays, number_of_processes=1, uniformization=False, differential_mode=False):
        """Pairwise Restricted Directed Information for every delay in `delays` (causal_network.py:112-187).

        Returns (and stores in self.rdi_results) {delay: G x G DataFrame (rows = source, columns = target,
        NaN diagonal)} plus "MAX", the element-wise maximum over delays (-inf diagonal).
        `number_of_processes` is accepted and ignored: parallelism comes from the GPU(s)."""
        plan = self._rdi_plan(delays, differential_mode)
        vals = self._rdi_execute(plan, uniformization, differential_mode)
        delays, node_ids, src, tgt, per = plan["delays"], plan["node_ids"], plan["src"], plan["tgt"], plan["per"]
        G = len(node_ids)
        self.rdi_results = {}
        off = ~np.eye(G, dtype=bool)                      # jobs are target-major, sources ascending: the off-diagonal of M.T in C order
        for i, d in enumerate(delays):
            Mt = np.full((G, G), np.nan)
            Mt[off] = vals[i * per:(i + 1) * per]
            self.rdi_results[d] = self._frame(Mt.T, node_ids)
        self.rdi_results["MAX"] = self.__extract_max_rdi_value_delay(want_delays=False)[0]
        return self.rdi_results

    # ------------------------------------------------------------------ cRDI
    def crdi(self, L=1, number_of_processes=1, uniformization=False, differential_mode=False):
        """Conditional RDI given the top-L other incoming regulators of each target (causal_network.py:190-289)."""
        if self.rdi_results is None:
            warnings.warn("WARNING: The method first needs RDI to be run, RUNNING RDI NOW... delay set to [1]")
            self.rdi(delays=[1], number_of_processes=number_of_processes, uniformization=uniformization,
                     differential_mode=differential_mode)
        if L > _lib.MAX_COND:
            raise ValueError("crdi supports at most %d conditioning genes" % _lib.MAX_COND)
        node_ids, run_off, E = self._ensure_resident()
        G = len(node_ids)
        pos = {g: i for i, g in enumerate(node_ids)}
        max_val, max_delay = self.__extract_max_rdi_value_delay()
        top_nodes, top_delays, top_values = self.__extract_top_incoming_nodes_delays(max_val, max_delay, k_nodes=L)

        mv = max_delay.to_numpy()
        # Job table, target-major (a rank's slab then shares its targets' columns).  For a target b every source uses the
        # same conditioning set -- top[b] minus its first minimal slot (causal_network.py:233-235) -- except the L+1
        # sources that are themselves in top[b], which drop their own slot instead (:230-232).
        jobs = np.zeros(G * (G - 1), dtype=_lib.JOB_DTYPE)
        j = 0
        for b_i, b in enumerate(node_ids):
            nodes, dls, vals = top_nodes[b], top_delays[b], top_values[b]
            if any(nd is None for nd in nodes):
                raise ValueError("crdi needs at least L+1 scored incoming regulators per target")
            slot_gene = np.array([pos[nd] for nd in nodes], dtype=np.int32)
            slot_delay = np.array([int(d) for d in dls], dtype=np.int32)
            src = np.concatenate((np.arange(b_i, dtype=np.int32), np.arange(b_i + 1, G, dtype=np.int32)))
            sl = slice(j, j + G - 1)
            jobs["src"][sl] = src; jobs["dst"][sl] = b_i; jobs["n_cond"][sl] = L
            jobs["delay"][sl] = mv[src, b_i].astype(np.int32)
            keep = [s_ for s_ in range(L + 1) if s_ != vals.index(min(vals))]
            jobs["cond_gene"][sl, :L] = slot_gene[keep]; jobs["cond_delay"][sl, :L] = slot_delay[keep]
            for s_own, g_own in enumerate(slot_gene):                       # sources inside the top set
                row = j + (g_own if g_own < b_i else g_own - 1)
                keep_own = [s_ for s_ in range(L + 1) if s_ != s_own]
                jobs["cond_gene"][row, :L] = slot_gene[keep_own]; jobs["cond_delay"][row, :L] = slot_delay[keep_own]
            j += G - 1
        h = _lib.default_handle()
        scales = None
        if differential_mode:
            scales = self._crdi_scales(h.read_matrix() if E is None else E, run_off, jobs, L)
        opts = _lib.make_opts(_lib.EST_CUMI if uniformization else _lib.EST_CMI, 5)

        def evaluate(lo, hi, out_dev_ptr=None):
            return h.lagged_jobs(jobs[lo:hi], opts, differential_mode, None if scales is None else scales[lo:hi],
                                 out_dev_ptr=out_dev_ptr)

        vals = _dist.sharded_evaluate(len(jobs), evaluate, h)
        M = np.full((G, G), np.nan)
        M[jobs["src"], jobs["dst"]] = vals
        self.crdi_results = self._frame(M, node_ids)
        return self.crdi_results

    # ------------------------------------------------------------------ reductions over the RDI matrices
    def __square(self, frame, node_ids, dtype=np.float64):
        """frame.loc[node_ids, node_ids] as an array, without the label lookup when the frame is already in node order."""
        ix = getattr(self, "_ix_cache", None)
        if ix is not None and frame.index is ix[1] and frame.columns is ix[1] and ix[0] == node_ids:   # built by _frame()
            return frame.to_numpy(dtype=dtype)
        if frame.shape == (len(node_ids), len(node_ids)) and list(frame.index) == node_ids and list(frame.columns) == node_ids:
            return frame.to_numpy(dtype=dtype)
        return frame.loc[node_ids, node_ids].to_numpy(dtype=dtype)

    def __extract_max_rdi_value_delay(self, want_delays=True):
        """Max over delays and the delay attaining it; strict '>' so the first delay wins ties, the diagonal
        keeps -inf / NaN (causal_network.py:295-311, without the removed np.int).  want_delays=False skips the
        (object-typed) delay matrix, which rdi() itself never uses."""
        node_ids = list(self.node_ids)
        G = len(node_ids)
        val = np.full((G, G), -np.inf)
        arg = np.full((G, G), np.nan, dtype=object) if want_delays else None
        for delay, frame in self.rdi_results.items():
            if isinstance(delay, str) and delay == "MAX":
                continue
            M = self.__square(frame, node_ids)
            better = M > val
            np.fill_diagonal(better, False)
            np.copyto(val, M, where=better)
            if want_delays:
                arg[better] = delay
        if not want_delays:
            return self._frame(val, node_ids), None
        return self._frame(val, node_ids), pd.DataFrame(arg, index=node_ids, columns=node_ids, dtype=object)

    def __extract_top_incoming_nodes_delays(self, max_rdi_values, max_rdi_delays, k_nodes):
        """Per target the k_nodes+1 strongest sources by streaming replace-first-minimum
        (causal_network.py:316-333).  The reference's double loop (targets x sources) is run here as one pass over
        the sources that updates all targets at once: identical slot contents and slot order, G numpy steps
        instead of G^2 Python iterations."""
        node_ids = list(self.node_ids)
        G, K = len(node_ids), k_nodes + 1
        V = self.__square(max_rdi_values, node_ids)
        Dl = self.__square(max_rdi_delays, node_ids, dtype=object)
        vals = np.full((G, K), -np.inf)                    # [target, slot]
        who = np.full((G, K), -1, dtype=np.int64)
        tgt = np.arange(G)
        for a_i in range(G):
            lo = vals.min(axis=1)
            slot = vals.argmin(axis=1)                     # FIRST minimal slot, like list.index(min(...))
            upd = V[a_i, :] > lo                           # strict: an equal value never replaces; NaN never enters
            upd[a_i] = False                               # a source is not its own regulator
            t = tgt[upd]
            vals[t, slot[upd]] = V[a_i, upd]
            who[t, slot[upd]] = a_i
        top_nodes, top_delays, top_values = {}, {}, {}
        for b_i, b in enumerate(node_ids):
            w = who[b_i]
            top_nodes[b] = [node_ids[i] if i >= 0 else None for i in w]
            top_delays[b] = [Dl[i, b_i] if i >= 0 else np.nan for i in w]
            top_values[b] = [float(v) for v in vals[b_i]]
        return top_nodes, top_delays, top_values

    # ------------------------------------------------------------------ evaluation (pass-through)
    def roc(self, results, true_graph_path):
        """ROC curve and AUROC of a score matrix against an edge-list file (causal_network.py:337-389): every ordered pair of
        node_ids, ranked by descending score (stable: ties keep source-major order), true / false positive counts along the
        ranking normalised by their totals, trapezoid area.  The ranking, the cumulative counts and the area run on the GPU
        (scribe_roc); the file parsing stays on the host.  Raises ZeroDivisionError, like the reference, when the ranking holds
        no true or no false edge."""
        with open(true_graph_path) as fh:
            true_edges = set(line.strip().lower() for line in fh)
        ids = list(self.node_ids)
        low = [str(a).lower() for a in ids]
        scores = self.__square(results, ids)
        G = len(ids)
        truth = np.zeros((G, G), dtype=np.uint8)
        for i, a in enumerate(low):
            for j, b in enumerate(low):
                if i != j and (a + "\t" + b) in true_edges:
                    truth[i, j] = 1
        auroc, x, y = _lib.default_handle().roc(scores, truth)
        return auroc, list(x), list(y)
