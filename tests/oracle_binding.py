"""ctypes binding of oracle/libarrowspace_oracle.so -- the CHECKER.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

TAU_FIXED, TAU_MEDIAN, TAU_MEAN, TAU_PERCENTILE = 0, 1, 2, 3
_P, _I64, _D = C.c_void_p, C.c_int64, C.c_double


class LapParams(C.Structure):
    _fields_ = [("eps", _D), ("k", _I64), ("topk", _I64), ("p", _D), ("has_sigma", C.c_int), ("sigma", _D),
                ("normalise", C.c_int), ("sparsity_check", C.c_int), ("self_included", C.c_int),
                ("rectified", C.c_int)]


class OracleError(RuntimeError):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def _p(a):
    return None if a is None else a.ctypes.data


class Oracle:
    def __init__(self):
        import arrowspace_b200 as asb
        path = asb._build.build_oracle()
        lib = C.CDLL(str(path))
        lib.aso_select_tau.restype = _D
        lib.aso_select_tau.argtypes = [_P, C.c_size_t, C.c_int, _D]
        lib.aso_synthetic_lambda.restype = _D
        lib.aso_synthetic_lambda.argtypes = [_P, _I64, _P, _P, _P, _D]
        lib.aso_compute_taumode.argtypes = [_P, _I64, _I64, _P, _P, _P, C.c_int, _D, _P, C.c_int]
        lib.aso_prepare_query_item.argtypes = [_P, _I64, _P, _P, _P, C.c_int, _D, C.POINTER(_D)]
        lib.aso_nearest_centroid.restype = _I64
        lib.aso_nearest_centroid.argtypes = [_P, _P, _I64, _I64, C.POINTER(_D)]
        lib.aso_cluster_incremental.argtypes = [_P, _I64, _I64, _I64, _D, _P, _P, _P, C.POINTER(_I64)]
        lib.aso_twonn_distances.argtypes = [_P, _I64, _I64, _P, _I64, _P, _P, C.c_int]
        lib.aso_intrinsic_dim.restype = _I64
        lib.aso_intrinsic_dim.argtypes = [_I64, _I64, _P, _P, _I64]
        lib.aso_step1_bounds.restype = None
        lib.aso_step1_bounds.argtypes = [_I64, _I64, _I64, C.POINTER(_I64), C.POINTER(_I64)]
        lib.aso_feature_laplacian.argtypes = [_P, _I64, _I64, C.POINTER(LapParams), _P, _P, _P, C.POINTER(_I64)]
        lib.aso_spectral_signals.argtypes = [_P, _P, _P, _I64, C.POINTER(LapParams), _P, _P, _P, C.POINTER(_I64)]
        lib.aso_search_lambda_aware.argtypes = [_P, _P, _I64, _I64, _P, _D, _I64, _D, _P, _P, C.POINTER(_I64)]
        lib.aso_search_lambda_aware_batch.argtypes = [_P, _P, _I64, _I64, _P, _P, _I64, _I64, _D, _P, _P, _P, C.c_int]
        lib.aso_search_lambda_aware_hybrid.argtypes = [_P, _P, _I64, _I64, _P, _D, _I64, _D, _P, _P, C.POINTER(_I64)]
        lib.aso_range_search.argtypes = [_P, _I64, _D, _D, _P, _P, C.POINTER(_I64)]
        lib.aso_jl_dimension.restype = _I64
        lib.aso_jl_dimension.argtypes = [_I64, _D]
        lib.aso_project_matrix.argtypes = [_P, _I64, _I64, _P, _I64, _P]
        lib.aso_search_energy.argtypes = [_P, _P, _I64, _I64, _P, _D, _I64, _D, _D, _P, _P, C.POINTER(_I64)]
        lib.aso_search_energy_ex.argtypes = [_P, _P, _I64, _I64, _P, _D, _I64, _D, _D, _P, _I64, _P, _P, _P, _I64, _P, _P,
                                             C.POINTER(_I64)]
        lib.aso_num_threads.restype = C.c_int
        lib.aso_synthetic_lambda_g.restype = _D
        lib.aso_synthetic_lambda_g.argtypes = [_P, _I64, _I64, _P, _P, _P, _D]
        lib.aso_set_num_threads.restype = C.c_int
        lib.aso_set_num_threads.argtypes = [C.c_int]
        self.lib = lib

    @staticmethod
    def _chk(rc):
        if rc != 0:
            raise OracleError(rc)

    def num_threads(self) -> int:
        return int(self.lib.aso_num_threads())

    def set_num_threads(self, threads: int) -> int:
        """OpenMP team size of every later call (overrides an inherited OMP_NUM_THREADS); returns the size in effect."""
        return int(self.lib.aso_set_num_threads(int(threads)))

    def select_tau(self, x, mode, value=0.0) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        return float(self.lib.aso_select_tau(_p(x), len(x), mode, value))

    def synthetic_lambda(self, x, csr, tau) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        ip, ii, dd = csr
        return float(self.lib.aso_synthetic_lambda(_p(x), len(x), _p(ip), _p(ii), _p(dd), tau))

    def synthetic_lambda_prefix(self, x, csr, tau, rg) -> float:
        """lambda of an F-long item against a graph of rg <= F nodes (JL-projected build): the Rayleigh sums read
        x[0 .. rg), the denominator the whole item."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        ip, ii, dd = csr
        return float(self.lib.aso_synthetic_lambda_g(_p(x), len(x), int(rg), _p(ip), _p(ii), _p(dd), tau))

    def compute_taumode(self, items, csr, mode, value=0.0, threads=0) -> np.ndarray:
        items = np.ascontiguousarray(items, dtype=np.float64)
        n, f = items.shape
        ip, ii, dd = csr
        out = np.empty(n, dtype=np.float64)
        self._chk(self.lib.aso_compute_taumode(_p(items), n, f, _p(ip), _p(ii), _p(dd), mode, value, _p(out), threads))
        return out

    def prepare_query_item(self, q, csr, mode, value=0.0) -> float:
        q = np.ascontiguousarray(q, dtype=np.float64)
        ip, ii, dd = csr
        out = _D(0)
        self._chk(self.lib.aso_prepare_query_item(_p(q), len(q), _p(ip), _p(ii), _p(dd), mode, value, C.byref(out)))
        return out.value

    def nearest_centroid(self, row, centroids):
        row = np.ascontiguousarray(row, dtype=np.float64)
        c = np.ascontiguousarray(centroids, dtype=np.float64)
        d2 = _D(0)
        i = self.lib.aso_nearest_centroid(_p(row), _p(c), c.shape[0], c.shape[1], C.byref(d2))
        return int(i), d2.value

    def cluster_incremental(self, rows, max_clusters, radius):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        n, f = rows.shape
        cent = np.zeros((max_clusters, f), dtype=np.float64)
        asg = np.empty(n, dtype=np.int64)
        sizes = np.zeros(max_clusters, dtype=np.uint64)
        x = _I64(0)
        self._chk(self.lib.aso_cluster_incremental(_p(rows), n, f, max_clusters, radius, _p(cent), _p(asg), _p(sizes),
                                                   C.byref(x)))
        return cent[: x.value].copy(), asg, sizes[: x.value].copy()

    def twonn_distances(self, rows, sample_idx, threads=0):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        si = np.ascontiguousarray(sample_idx, dtype=np.int64)
        d1 = np.empty(len(si))
        d2 = np.empty(len(si))
        self._chk(self.lib.aso_twonn_distances(_p(rows), rows.shape[0], rows.shape[1], _p(si), len(si), _p(d1), _p(d2),
                                               threads))
        return d1, d2

    def intrinsic_dim(self, n, f, d1, d2) -> int:
        d1 = np.ascontiguousarray(d1, dtype=np.float64)
        d2 = np.ascontiguousarray(d2, dtype=np.float64)
        return int(self.lib.aso_intrinsic_dim(n, f, _p(d1), _p(d2), len(d1)))

    def step1_bounds(self, n, f, id_est):
        a, b = _I64(0), _I64(0)
        self.lib.aso_step1_bounds(n, f, id_est, C.byref(a), C.byref(b))
        return a.value, b.value

    def feature_laplacian(self, centroids, eps, k, topk, p, sigma, normalise=False, sparsity_check=False,
                          self_included=False, rectified=False):
        c = np.ascontiguousarray(centroids, dtype=np.float64)
        x, f = c.shape
        cap = max(f * (1 + 2 * (topk + 1)), f) + 8
        ip = np.zeros(f + 1, dtype=np.int64)
        ii = np.zeros(cap, dtype=np.int64)
        dd = np.zeros(cap, dtype=np.float64)
        nnz = _I64(0)
        P = LapParams(eps, k, topk, p, 1 if sigma is not None else 0, sigma if sigma is not None else 0.0,
                      int(normalise), int(sparsity_check), int(self_included), int(rectified))
        self._chk(self.lib.aso_feature_laplacian(_p(c), x, f, C.byref(P), _p(ip), _p(ii), _p(dd), C.byref(nnz)))
        return ip, ii[: nnz.value].copy(), dd[: nnz.value].copy()

    def spectral_signals(self, csr, eps, k, topk, p, sigma, normalise=False, sparsity_check=False,
                         self_included=False, rectified=False):
        """aspace.signals from the feature Laplacian ``csr`` (src/graph.rs:211-231)."""
        lp = np.ascontiguousarray(csr[0], dtype=np.int64)
        li = np.ascontiguousarray(csr[1], dtype=np.int64)
        ld = np.ascontiguousarray(csr[2], dtype=np.float64)
        f = len(lp) - 1
        cap = max(f * (1 + 2 * (topk + 1)), f) + 8
        ip = np.zeros(f + 1, dtype=np.int64)
        ii = np.zeros(cap, dtype=np.int64)
        dd = np.zeros(cap, dtype=np.float64)
        nnz = _I64(0)
        P = LapParams(eps, k, topk, p, 1 if sigma is not None else 0, sigma if sigma is not None else 0.0,
                      int(normalise), int(sparsity_check), int(self_included), int(rectified))
        self._chk(self.lib.aso_spectral_signals(_p(lp), _p(li), _p(ld), f, C.byref(P), _p(ip), _p(ii), _p(dd),
                                                C.byref(nnz)))
        return ip, ii[: nnz.value].copy(), dd[: nnz.value].copy()

    def jl_dimension(self, n_points, eps):
        return int(self.lib.aso_jl_dimension(int(n_points), float(eps)))

    def project_matrix(self, rows, projection):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        projection = np.ascontiguousarray(projection, dtype=np.float64)
        n, f = rows.shape
        r = projection.shape[1]
        out = np.empty((n, r), dtype=np.float64)
        self._chk(self.lib.aso_project_matrix(_p(rows), n, f, _p(projection), r, _p(out)))
        return out

    def search_energy(self, items, lambdas, q, lambda_q, k, w_lambda, w_dirichlet):
        items = np.ascontiguousarray(items, dtype=np.float64)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        n, f = items.shape
        idx = np.full(max(k, 1), -1, dtype=np.int64)
        sc = np.zeros(max(k, 1), dtype=np.float64)
        cnt = _I64(0)
        self._chk(self.lib.aso_search_energy(_p(items), _p(lambdas), n, f, _p(q), lambda_q, k, w_lambda, w_dirichlet,
                                             _p(idx), _p(sc), C.byref(cnt)))
        return [(int(idx[r]), float(sc[r])) for r in range(cnt.value)]

    def search_energy_ex(self, items, lambdas, q, lambda_q, k, w_lambda, w_dirichlet, projection=None, signals=None):
        """Every branch of ProjectedEnergy::score: ``projection`` F x r (or None), ``signals`` CSR triple (or None)."""
        items = np.ascontiguousarray(items, dtype=np.float64)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        n, f = items.shape
        idx = np.full(max(k, 1), -1, dtype=np.int64)
        sc = np.zeros(max(k, 1), dtype=np.float64)
        cnt = _I64(0)
        pj = np.ascontiguousarray(projection, dtype=np.float64) if projection is not None else None
        r = pj.shape[1] if pj is not None else 0
        if signals is not None:
            sp, si, sd = (np.ascontiguousarray(signals[0], dtype=np.int64), np.ascontiguousarray(signals[1], dtype=np.int64),
                          np.ascontiguousarray(signals[2], dtype=np.float64))
            sn = len(sp) - 1
        else:
            sp = si = sd = None
            sn = 0
        self._chk(self.lib.aso_search_energy_ex(_p(items), _p(lambdas), n, f, _p(q), lambda_q, k, w_lambda, w_dirichlet,
                                                _p(pj) if pj is not None else None, r,
                                                _p(sp) if sp is not None else None, _p(si) if si is not None else None,
                                                _p(sd) if sd is not None else None, sn, _p(idx), _p(sc), C.byref(cnt)))
        return [(int(idx[r_]), float(sc[r_])) for r_ in range(cnt.value)]

    def search_lambda_aware(self, items, lambdas, q, lambda_q, k, alpha):
        items = np.ascontiguousarray(items, dtype=np.float64)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        n, f = items.shape
        idx = np.full(max(k, 1), -1, dtype=np.int64)
        sc = np.zeros(max(k, 1), dtype=np.float64)
        cnt = _I64(0)
        self._chk(self.lib.aso_search_lambda_aware(_p(items), _p(lambdas), n, f, _p(q), lambda_q, k, alpha, _p(idx),
                                                   _p(sc), C.byref(cnt)))
        return [(int(idx[r]), float(sc[r])) for r in range(cnt.value)]

    def search_lambda_aware_batch(self, items, lambdas, queries, lambda_q, k, alpha, threads=0):
        items = np.ascontiguousarray(items, dtype=np.float64)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        queries = np.ascontiguousarray(queries, dtype=np.float64)
        lambda_q = np.ascontiguousarray(lambda_q, dtype=np.float64)
        n, f = items.shape
        nq = queries.shape[0]
        idx = np.full((nq, max(k, 1)), -1, dtype=np.int64)
        sc = np.zeros((nq, max(k, 1)), dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int64)
        self._chk(self.lib.aso_search_lambda_aware_batch(_p(items), _p(lambdas), n, f, _p(queries), _p(lambda_q), nq, k,
                                                         alpha, _p(idx), _p(sc), _p(cnt), threads))
        return idx, sc, cnt

    def search_lambda_aware_hybrid(self, items, lambdas, q, lambda_q, k, alpha):
        items = np.ascontiguousarray(items, dtype=np.float64)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        n, f = items.shape
        idx = np.full(max(k, 1), -1, dtype=np.int64)
        sc = np.zeros(max(k, 1), dtype=np.float64)
        cnt = _I64(0)
        self._chk(self.lib.aso_search_lambda_aware_hybrid(_p(items), _p(lambdas), n, f, _p(q), lambda_q, k, alpha,
                                                          _p(idx), _p(sc), C.byref(cnt)))
        return [(int(idx[r]), float(sc[r])) for r in range(cnt.value)]

    def range_search(self, lambdas, lambda_q, eps):
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        n = len(lambdas)
        idx = np.empty(max(n, 1), dtype=np.int64)
        dist = np.empty(max(n, 1), dtype=np.float64)
        cnt = _I64(0)
        self._chk(self.lib.aso_range_search(_p(lambdas), n, lambda_q, eps, _p(idx), _p(dist), C.byref(cnt)))
        return idx[: cnt.value].copy(), dist[: cnt.value].copy()
