"""ctypes binding of libscribe_b200.so (the C-ABI in include/scribe_b200.h).

There is deliberately no CPU fallback here: if the shared library is missing or no CUDA device is
usable, every compute call raises.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCRIBE_B200_LIB") or os.path.join(_HERE, "lib", "libscribe_b200.so")   # override: tuning builds

SCRIBE_OK, E_ARG, E_LENGTH, E_CUDA, E_DOMAIN, E_STATE = 0, 1, 2, 3, 4, 5
EST_MI, EST_CMI, EST_UMI, EST_CUMI = 0, 1, 2, 3
DENSITY_KDE, DENSITY_KNN = 0, 1
PATH_AUTO, PATH_GENERIC, PATH_DENSE, PATH_SWEEP, PATH_STRIP = 0, 1, 2, 3, 4
MAX_COND = 4


class ScribeCudaError(RuntimeError):
    """CUDA runtime failure, missing device or missing native library."""


class Opts(C.Structure):
    _fields_ = [("estimator", C.c_int32), ("k", C.c_int32), ("density", C.c_int32), ("k_density", C.c_int32),
                ("path", C.c_int32), ("reserved", C.c_int32), ("bw", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("estimator_launches", C.c_int64), ("estimator_ms", C.c_double),
                ("total_ms", C.c_double), ("point_pairs", C.c_int64), ("jobs", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("visited_knn", C.c_int64), ("visited_count", C.c_int64), ("visited_kde", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


JOB_DTYPE = np.dtype([("src", np.int32), ("dst", np.int32), ("delay", np.int32), ("n_cond", np.int32),
                      ("cond_gene", np.int32, (MAX_COND,)), ("cond_delay", np.int32, (MAX_COND,))])
assert JOB_DTYPE.itemsize == 48

_lib = None
_lib_lock = threading.Lock()

# every symbol include/scribe_b200.h declares
EXPORTS = [
    "scribe_abi_version", "scribe_last_error", "scribe_create", "scribe_destroy", "scribe_device_info",
    "scribe_get_stats", "scribe_device_index", "scribe_estimate", "scribe_debug_counts", "scribe_upload_matrix", "scribe_normalize_resident", "scribe_smooth_resident", "scribe_read_matrix", "scribe_lagged_jobs",
    "scribe_lagged_jobs_dev", "scribe_mi_jobs", "scribe_upload_coupling", "scribe_upload_coupling_raw", "scribe_read_coupling_row", "scribe_coupling_jobs", "scribe_coupling_jobs_dev",
    "scribe_clr", "scribe_top_edges", "scribe_roc", "scribe_knn_regress",
]


def load_library():
    """dlopen the native library and declare prototypes.  Raises ScribeCudaError when it is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ScribeCudaError(
                f"{LIB_PATH} not found: build it with ./build.sh (nvcc, sm_100a). scribe_py_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        P, D, I32, I64 = C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_int64
        lib.scribe_abi_version.restype = C.c_int
        lib.scribe_last_error.restype = C.c_char_p
        lib.scribe_create.argtypes = [C.c_int, C.POINTER(P)]
        lib.scribe_destroy.argtypes = [P]
        lib.scribe_destroy.restype = None
        lib.scribe_device_info.argtypes = [P, C.POINTER(I32), C.POINTER(I32), C.POINTER(I64), C.c_char_p, I32]
        lib.scribe_get_stats.argtypes = [P, C.POINTER(Stats)]
        lib.scribe_device_index.argtypes = [P]
        lib.scribe_estimate.argtypes = [P, P, P, P, I64, I32, I32, I32, C.POINTER(Opts), D]
        lib.scribe_debug_counts.argtypes = [P, P, P, P, I64, I32, I32, I32, C.POINTER(Opts), P, P, P, P]
        lib.scribe_upload_matrix.argtypes = [P, P, I64, I64, P, I32]
        lib.scribe_normalize_resident.argtypes = [P, I32, I32]
        lib.scribe_smooth_resident.argtypes = [P, I32, C.POINTER(I64), P]
        lib.scribe_read_matrix.argtypes = [P, I64, I64, P]
        lib.scribe_lagged_jobs.argtypes = [P, P, I64, C.POINTER(Opts), I32, P, P]
        lib.scribe_lagged_jobs_dev.argtypes = [P, P, I64, C.POINTER(Opts), I32, P, P]
        lib.scribe_mi_jobs.argtypes = [P, P, P, I64, C.POINTER(Opts), P]
        lib.scribe_upload_coupling.argtypes = [P, P, P, I64, I64]
        lib.scribe_upload_coupling_raw.argtypes = [P, P, P, I64, I64, I32, I32, C.POINTER(I32)]
        lib.scribe_read_coupling_row.argtypes = [P, I32, I64, P]
        lib.scribe_coupling_jobs.argtypes = [P, P, P, I64, I32, C.POINTER(Opts), P, P]
        lib.scribe_coupling_jobs_dev.argtypes = [P, P, P, I64, I32, C.POINTER(Opts), P, P]
        lib.scribe_clr.argtypes = [P, P, I64, I64, I32, P]
        lib.scribe_top_edges.argtypes = [P, P, I64, I64, P, P, P, C.POINTER(I64)]
        lib.scribe_roc.argtypes = [P, P, P, I64, D, P, P]
        lib.scribe_knn_regress.argtypes = [P, P, P, P, I64, P, P, I64, I32, P]
        _lib = lib
        return lib


def _raise(code):
    msg = load_library().scribe_last_error().decode("utf-8", "replace")
    if code == E_LENGTH:
        raise AssertionError(msg)
    if code == E_DOMAIN:
        raise ValueError(msg)                 # math.log(0) in the reference (information_estimators.py:274)
    if code == E_ARG:
        raise ValueError(msg)
    raise ScribeCudaError(msg)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_PATH_NAMES = {"auto": PATH_AUTO, "generic": PATH_GENERIC, "dense": PATH_DENSE, "sweep": PATH_SWEEP, "strip": PATH_STRIP}


def default_path():
    """Kernel path used by the drivers: SCRIBE_B200_PATH=auto|dense|sweep|strip|generic (default auto)."""
    return _PATH_NAMES[os.environ.get("SCRIBE_B200_PATH", "auto").lower()]


def make_opts(estimator, k=5, density="kde", k_density=5, bw=0.01, path=None):
    if path is None:
        path = default_path()
    m = str(density).lower()
    if m == "kde":
        d = DENSITY_KDE
    elif m == "knn":
        d = DENSITY_KNN
    else:
        raise ValueError("The density estimation method is not recognized")   # information_estimators.py:259,:394
    return Opts(int(estimator), int(k), d, int(k_density), int(path), 0, float(bw))


class Handle:
    """One device context (stream + workspaces + resident matrices).  Not re-entrant."""

    def __init__(self, device=-1):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.scribe_create(int(device), C.byref(h))
        if rc != SCRIBE_OK:
            _raise(rc)
        self._h = h
        sm, clk, mem = C.c_int32(), C.c_int32(), C.c_int64()
        name = C.create_string_buffer(128)
        self._lib.scribe_device_info(h, C.byref(sm), C.byref(clk), C.byref(mem), name, 128)
        self.sm_count, self.clock_khz, self.mem_bytes = sm.value, clk.value, mem.value
        self.device_name = name.value.decode()
        self.device = int(self._lib.scribe_device_index(h))          # CUDA ordinal: collectives must run on the same GPU
        self.resident_token = 0                                      # bumped whenever the resident expression matrix changes owner
        self.resident_shape = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.scribe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        s = Stats()
        self._lib.scribe_get_stats(self._h, C.byref(s))
        return s.as_dict()

    # ---- estimator level ---------------------------------------------------------------------------
    @staticmethod
    def _soa(a):
        """(N, d) -> contiguous (d, N) float64."""
        return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)

    def estimate(self, x, y, z, opts):
        xs, ys = self._soa(x), self._soa(y)
        zs = self._soa(z) if z is not None else None
        n = xs.shape[1]
        out = C.c_double()
        rc = self._lib.scribe_estimate(self._h, _ptr(xs), _ptr(ys), _ptr(zs), n, xs.shape[0], ys.shape[0],
                                       0 if zs is None else zs.shape[0], C.byref(opts), C.byref(out))
        if rc != SCRIBE_OK:
            _raise(rc)
        return out.value

    def debug_counts(self, x, y, z, opts):
        xs, ys = self._soa(x), self._soa(y)
        zs = self._soa(z) if z is not None else None
        n = xs.shape[1]
        eps = np.empty(n)
        cnt = np.empty((n, 4), dtype=np.int32)
        weighted = opts.estimator in (EST_UMI, EST_CUMI)
        wsum = np.empty((n, 2)) if weighted else None
        w = np.empty(n) if weighted else None
        rc = self._lib.scribe_debug_counts(self._h, _ptr(xs), _ptr(ys), _ptr(zs), n, xs.shape[0], ys.shape[0],
                                           0 if zs is None else zs.shape[0], C.byref(opts),
                                           _ptr(eps), _ptr(cnt), _ptr(wsum), _ptr(w))
        if rc != SCRIBE_OK:
            _raise(rc)
        return {"eps": eps, "counts": cnt.T.copy(), "wsum": None if wsum is None else wsum.T.copy(), "weights": w}

    # ---- lagged jobs (rdi / crdi) ---------------------------------------------------------------------
    def upload_matrix(self, E, run_offsets=None):
        E = np.ascontiguousarray(E, dtype=np.float64)
        ro = None if run_offsets is None else np.ascontiguousarray(run_offsets, dtype=np.int64)
        rc = self._lib.scribe_upload_matrix(self._h, _ptr(E), E.shape[0], E.shape[1], _ptr(ro),
                                            0 if ro is None else len(ro) - 1)
        if rc != SCRIBE_OK:
            _raise(rc)
        self.resident_token += 1
        self.resident_shape = E.shape
        return self.resident_token

    def normalize_resident(self, mean_order, avg_order):
        rc = self._lib.scribe_normalize_resident(self._h, int(mean_order), int(avg_order))
        if rc != SCRIBE_OK:
            _raise(rc)

    def smooth_resident(self, window, n_runs):
        n = C.c_int64(0)
        off = np.zeros(n_runs + 1, dtype=np.int64)
        rc = self._lib.scribe_smooth_resident(self._h, int(window), C.byref(n), _ptr(off))
        if rc != SCRIBE_OK:
            _raise(rc)
        self.resident_shape = (self.resident_shape[0], int(n.value))
        return off

    def read_matrix(self, gene0=0, n_genes=None):
        G, N = self.resident_shape
        n_genes = G - gene0 if n_genes is None else n_genes
        out = np.empty((n_genes, N))
        rc = self._lib.scribe_read_matrix(self._h, int(gene0), int(n_genes), _ptr(out))
        if rc != SCRIBE_OK:
            _raise(rc)
        return out

    def lagged_jobs(self, jobs, opts, differential=False, scales=None, out_dev_ptr=None):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        sc = None if scales is None else np.ascontiguousarray(scales, dtype=np.float64)
        if out_dev_ptr is not None:
            rc = self._lib.scribe_lagged_jobs_dev(self._h, _ptr(jobs), len(jobs), C.byref(opts), int(bool(differential)),
                                                  _ptr(sc), C.c_void_p(out_dev_ptr))
            if rc != SCRIBE_OK:
                _raise(rc)
            return None
        out = np.empty(len(jobs))
        rc = self._lib.scribe_lagged_jobs(self._h, _ptr(jobs), len(jobs), C.byref(opts), int(bool(differential)),
                                          _ptr(sc), _ptr(out))
        if rc != SCRIBE_OK:
            _raise(rc)
        return out

    def mi_jobs(self, src, dst, opts):
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        out = np.empty(len(src))
        rc = self._lib.scribe_mi_jobs(self._h, _ptr(src), _ptr(dst), len(src), C.byref(opts), _ptr(out))
        if rc != SCRIBE_OK:
            _raise(rc)
        return out

    # ---- dynamics coupling --------------------------------------------------------------------------
    def upload_coupling(self, S, Y):
        S = np.ascontiguousarray(S, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        assert S.shape == Y.shape
        rc = self._lib.scribe_upload_coupling(self._h, _ptr(S), _ptr(Y), S.shape[0], S.shape[1])
        if rc != SCRIBE_OK:
            _raise(rc)

    def upload_coupling_raw(self, t0_layer, t1_layer, normalize=True, y_is_sum=True):
        """Raw cells x genes layers -> resident normalised S / Y (device normalisation).  Returns False when the t0 layer
        holds NaN (pandas would skip them in the mean): the caller must then take the host path."""
        a = np.ascontiguousarray(t0_layer, dtype=np.float64)
        b = np.ascontiguousarray(t1_layer, dtype=np.float64)
        assert a.shape == b.shape and a.ndim == 2
        flag = C.c_int32(0)
        rc = self._lib.scribe_upload_coupling_raw(self._h, _ptr(a), _ptr(b), a.shape[0], a.shape[1], int(bool(normalize)),
                                                  int(bool(y_is_sum)), C.byref(flag))
        if rc != SCRIBE_OK:
            _raise(rc)
        return flag.value == 0

    def read_coupling_row(self, which, gene, n_cells):
        out = np.empty(n_cells)
        rc = self._lib.scribe_read_coupling_row(self._h, int(which), int(gene), _ptr(out))
        if rc != SCRIBE_OK:
            _raise(rc)
        return out

    def coupling_jobs(self, src, dst, opts, drop_zero_cells=True, out_dev_ptr=None, want_n_eff=False):
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        neff = np.empty(len(src), dtype=np.int32) if want_n_eff else None
        if out_dev_ptr is not None:
            rc = self._lib.scribe_coupling_jobs_dev(self._h, _ptr(src), _ptr(dst), len(src), int(bool(drop_zero_cells)),
                                                    C.byref(opts), C.c_void_p(out_dev_ptr), _ptr(neff))
            if rc != SCRIBE_OK:
                _raise(rc)
            return (None, neff) if want_n_eff else None
        out = np.empty(len(src))
        rc = self._lib.scribe_coupling_jobs(self._h, _ptr(src), _ptr(dst), len(src), int(bool(drop_zero_cells)),
                                            C.byref(opts), _ptr(out), _ptr(neff))
        if rc != SCRIBE_OK:
            _raise(rc)
        return (out, neff) if want_n_eff else out


    # ---- consumers of the causal matrix ----------------------------------------------------------------
    def clr(self, M, both_dims):
        M = np.ascontiguousarray(M, dtype=np.float64)
        out = np.empty_like(M)
        rc = self._lib.scribe_clr(self._h, _ptr(M), M.shape[0], M.shape[1], int(bool(both_dims)), _ptr(out))
        if rc != SCRIBE_OK:
            _raise(rc)
        return out

    def top_edges(self, M, n_top):
        M = np.ascontiguousarray(M, dtype=np.float64)
        assert M.ndim == 2 and M.shape[0] == M.shape[1]
        n_top = int(min(n_top, M.size))
        src = np.empty(n_top, dtype=np.int32); dst = np.empty(n_top, dtype=np.int32); w = np.empty(n_top)
        found = C.c_int64(0)
        rc = self._lib.scribe_top_edges(self._h, _ptr(M), M.shape[0], n_top, _ptr(src), _ptr(dst), _ptr(w), C.byref(found))
        if rc != SCRIBE_OK:
            _raise(rc)
        k = found.value
        return src[:k], dst[:k], w[:k]

    def roc(self, scores, truth):
        scores = np.ascontiguousarray(scores, dtype=np.float64)
        truth = np.ascontiguousarray(truth, dtype=np.uint8)
        n = scores.shape[0]
        m = n * (n - 1) - int(np.isnan(scores[~np.eye(n, dtype=bool)]).sum())
        x = np.empty(m + 1); y = np.empty(m + 1)
        a = C.c_double()
        rc = self._lib.scribe_roc(self._h, _ptr(scores), _ptr(truth), n, C.byref(a), _ptr(x), _ptr(y))
        if rc == E_DOMAIN:
            raise ZeroDivisionError(load_library().scribe_last_error().decode())
        if rc != SCRIBE_OK:
            _raise(rc)
        return a.value, x, y


    def knn_regress(self, px, py, value, qx, qy, k=5):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (px, py, value, qx, qy)]
        assert len(a[0]) == len(a[1]) == len(a[2]) and len(a[3]) == len(a[4])
        out = np.empty(len(a[3]))
        rc = self._lib.scribe_knn_regress(self._h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), len(a[0]), _ptr(a[3]), _ptr(a[4]), len(a[3]),
                                          int(k), _ptr(out))
        if rc != SCRIBE_OK:
            _raise(rc)
        return out


_default = {}
_default_lock = threading.Lock()


def default_handle(device=None):
    """Process-wide handle for `device` (default: LOCAL_RANK if set, else the current CUDA device)."""
    if device is None:
        device = int(os.environ.get("SCRIBE_B200_DEVICE", os.environ.get("LOCAL_RANK", "-1")))
    with _default_lock:
        h = _default.get(device)
        if h is None:
            h = Handle(device)
            _default[device] = h
        return h
