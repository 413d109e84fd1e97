"""Host-side mirror of the reference's API for the lambda-tau build + lambda-aware search path.

Every class / method here has the name, argument meaning and error behaviour of its Rust
counterpart (cited per method, paths relative to the arrowspace-rs repository) and does
nothing but marshal buffers into the C ABI of ``libarrowspace_b200.so``
(``include/arrowspace_b200.h``).  There is no CPU arithmetic path: if the CUDA library or a
B200 is missing, construction of :class:`Context` raises.

Buffers may be numpy arrays (host) or torch CUDA tensors (device, zero copy).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _build

# ----------------------------------------------------------------------------- status codes
ASB_OK = 0
ASB_ERR_INVALID = 1
ASB_ERR_CUDA = 2
ASB_ERR_NCCL = 3
ASB_ERR_NONFINITE_QUERY = 4
ASB_ERR_ZERO_LAMBDA = 5
ASB_ERR_SHAPE = 6
ASB_ERR_TOO_SPARSE = 7
ASB_ERR_NO_CLUSTERS = 8
ASB_ERR_NAN_SCORE = 9
ASB_ERR_ZERO_NORM = 10
ASB_ERR_EMPTY = 11
ASB_ERR_DIM = 12
ASB_ERR_CAPACITY = 13
ASB_ERR_UNSUPPORTED = 14

TAU_FIXED, TAU_MEDIAN, TAU_MEAN, TAU_PERCENTILE = 0, 1, 2, 3


class ArrowSpaceError(RuntimeError):
    """Raised where the reference panics (the C ABI returns a status instead of aborting)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[asb status {status}] {message}")
        self.status = status
        self.message = message


class GraphParamsC(C.Structure):
    _fields_ = [
        ("eps", C.c_double), ("k", C.c_int64), ("topk", C.c_int64), ("p", C.c_double),
        ("has_sigma", C.c_int32), ("sigma", C.c_double), ("normalise", C.c_int32),
        ("sparsity_check", C.c_int32), ("self_included", C.c_int32), ("rectified", C.c_int32),
    ]


class BuildParamsC(C.Structure):
    _fields_ = [
        ("graph", GraphParamsC), ("tau_mode", C.c_int32), ("tau_value", C.c_double),
        ("max_clusters", C.c_int64), ("radius", C.c_double), ("apply_define_result_k", C.c_int32),
        ("spectral", C.c_int32), ("projection", C.c_void_p), ("reduced_dim", C.c_int64),
    ]


class IndexInfoC(C.Structure):
    _fields_ = [
        ("n_items", C.c_int64), ("n_features", C.c_int64), ("n_clusters", C.c_int64), ("nnz", C.c_int64),
        ("lambda_min", C.c_double), ("lambda_max", C.c_double), ("lambda_sum", C.c_double),
        ("radius", C.c_double), ("max_clusters", C.c_int64),
        ("ms_cluster", C.c_double), ("ms_laplacian", C.c_double), ("ms_taumode", C.c_double),
        ("ms_total", C.c_double), ("nnz_signals", C.c_int64),
    ]


# every symbol include/arrowspace_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_I64 = C.c_int64
_D = C.c_double
ABI_SYMBOLS = {
    "asb_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "asb_ctx_destroy": (None, [_P]),
    "asb_last_error": (C.c_char_p, [_P]),
    "asb_status_string": (C.c_char_p, [C.c_int]),
    "asb_version": (C.c_char_p, []),
    "asb_kernel_launches": (_I64, [_P]),
    "asb_last_kernel_ms": (_D, [_P, C.c_char_p]),
    "asb_ctx_set_option": (C.c_int, [_P, C.c_char_p, _D]),
    "asb_twonn_distances": (C.c_int, [_P, _P, _I64, _I64, _P, _I64, _P, _P]),
    "asb_cluster_incremental": (C.c_int, [_P, _P, _I64, _I64, _I64, _D, _P, _P, _P, C.POINTER(_I64)]),
    "asb_cluster_incremental_resume": (C.c_int, [_P, _P, _I64, _I64, _I64, _D, _P, _P, _P, C.POINTER(_I64)]),
    "asb_laplacian_max_nnz": (_I64, [_I64, _I64]),
    "asb_build_feature_laplacian": (C.c_int, [_P, _P, _I64, _I64, C.POINTER(GraphParamsC), _P, _P, _P, _I64,
                                              C.POINTER(_I64)]),
    "asb_compute_taumode": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P, C.c_int32, _D, _P, _P, _P]),
    "asb_prepare_query_lambdas": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P, C.c_int32, _D, _P]),
    "asb_search_lambda_aware_batch": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, _I64, _I64, _D, _I64, _P, _P,
                                                _P]),
    "asb_project_matrix": (C.c_int, [_P, _P, _I64, _I64, _P, _I64, _P]),
    "asb_jl_dimension": (C.c_int64, [_I64, _D]),
    "asb_search_slab_plan": (C.c_int, [C.c_int, _I64, _I64, _I64, C.POINTER(_I64), C.POINTER(_I64)]),
    "asb_search_energy_batch": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, _I64, _I64, _D, _D, _I64, _P, _P, _P]),
    "asb_search_lambda_aware_hybrid_batch": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, _I64, _I64, _D, _P, _P, _P]),
    "asb_range_search": (C.c_int, [_P, _P, _I64, _D, _D, _I64, _P, _P, _I64, C.POINTER(_I64)]),
    "asb_topk_merge": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _P, _P]),
    "asb_index_build": (C.c_int, [_P, _P, _I64, _I64, C.POINTER(BuildParamsC), C.POINTER(_P)]),
    "asb_index_destroy": (None, [_P]),
    "asb_index_info_get": (C.c_int, [_P, C.POINTER(IndexInfoC)]),
    "asb_index_lambdas": (C.c_int, [_P, _P, _P]),
    "asb_index_centroids": (C.c_int, [_P, _P, _P]),
    "asb_index_assignments": (C.c_int, [_P, _P, _P]),
    "asb_index_cluster_sizes": (C.c_int, [_P, _P, _P]),
    "asb_index_laplacian": (C.c_int, [_P, _P, _P, _P, _P]),
    "asb_index_signals": (C.c_int, [_P, _P, _P, _P, _P]),
    "asb_index_search": (C.c_int, [_P, _P, _P, _I64, _I64, _D, _P, _P, _P, _P]),
    "asb_index_search_lambda_aware": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _D, _P, _P, _P]),
    "asb_index_prepare_query": (C.c_int, [_P, _P, _P, _I64, _P]),
    "asb_index_search_energy": (C.c_int, [_P, _P, _P, _I64, _I64, _D, _D, _P, _P, _P]),
    "asb_comm_unique_id": (C.c_int, [_P, _P]),
    "asb_comm_init_rank": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "asb_comm_from_nccl": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "asb_comm_destroy": (None, [_P]),
    "asb_comm_rank": (C.c_int, [_P]),
    "asb_comm_size": (C.c_int, [_P]),
    "asb_twonn_distances_sharded": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _I64, _P, _P]),
    "asb_index_build_sharded": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, C.POINTER(BuildParamsC), C.POINTER(_P)]),
    "asb_index_shard_offset": (_I64, [_P]),
    "asb_index_search_sharded": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _D, _P, _P, _P, _P]),
}

_lib = None


def load_library(build: bool = True) -> C.CDLL:
    """dlopen the in-tree CUDA library (building it first when stale) and type every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if build:
        _build.build_cuda()
    if not _build.LIB_PATH.exists():
        raise ArrowSpaceError(ASB_ERR_CUDA, f"{_build.LIB_PATH} missing: the CUDA extension is the only "
                                            "implementation (no CPU fallback)")
    lib = C.CDLL(str(_build.LIB_PATH))
    for name, (res, args) in ABI_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def search_slab_plan(sm_count: int, nq: int, n: int, max_slabs: int = 4096) -> Tuple[int, int]:
    """(nslabs, tiles_per_slab) the search kernels use for nq queries x n items (host arithmetic only)."""
    ns, tps = _I64(0), _I64(0)
    rc = load_library().asb_search_slab_plan(int(sm_count), int(nq), int(n), int(max_slabs), C.byref(ns), C.byref(tps))
    if rc != 0:
        raise ArrowSpaceError(rc, "search_slab_plan: bad sizes")
    return int(ns.value), int(tps.value)


def _ptr(a) -> int:
    """Address of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if getattr(a, "is_cuda", False):
            # The library runs on its context's stream, which has no ordering against torch's streams: whatever torch
            # still has in flight for this tensor must be complete before the pointer is handed over (header contract:
            # device buffers are ready when the call is made).  Idle streams make this a no-op.
            import torch
            torch.cuda.current_stream(a.device).synchronize()
        return a.data_ptr()
    raise TypeError(f"unsupported buffer type {type(a)}")


def _check_tensor(t, what: str = "buffer"):
    """A torch tensor crosses the C ABI as a raw pointer: it must already be what the kernels expect -- float64 (or the
    integer type the argument names), contiguous, and when it lives on a GPU that GPU must be usable by the caller's
    context (the library classifies pointers, it cannot see dtypes or strides)."""
    import torch
    if not t.is_contiguous():
        raise ArrowSpaceError(ASB_ERR_INVALID, f"{what}: torch tensor must be contiguous (row-major)")
    if t.dtype not in (torch.float64, torch.int64, torch.uint64):
        raise ArrowSpaceError(ASB_ERR_INVALID, f"{what}: torch tensor must be float64 (got {t.dtype})")
    return t


def _as_f64_matrix(rows) -> np.ndarray:
    """Vec<Vec<f64>> -> contiguous row-major f64 (the one repack the shim does)."""
    if hasattr(rows, "data_ptr"):
        import torch
        if rows.dtype != torch.float64:
            raise ArrowSpaceError(ASB_ERR_INVALID, f"rows: torch tensor must be float64 (got {rows.dtype})")
        if rows.dim() != 2:
            raise ArrowSpaceError(ASB_ERR_INVALID, "rows must be a 2-D matrix")
        return _check_tensor(rows, "rows")
    a = np.ascontiguousarray(rows, dtype=np.float64)
    if a.ndim != 2:
        raise ArrowSpaceError(ASB_ERR_INVALID, "rows must be a 2-D matrix")
    return a


def _shape2(a) -> Tuple[int, int]:
    return int(a.shape[0]), int(a.shape[1])


def _is_device(a) -> bool:
    return hasattr(a, "is_cuda") and bool(a.is_cuda)


class Context:
    """One CUDA stream + workspace (SURVEY 8b "Threading"): one context per host thread."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = load_library()
        h = _P()
        rc = self.lib.asb_ctx_create(int(device), _P(stream) if stream else None, C.byref(h))
        if rc != ASB_OK:
            raise ArrowSpaceError(rc, "asb_ctx_create failed: no usable B200 (sm_100) device / CUDA runtime; "
                                      "this library has no CPU fallback")
        self.handle = h
        self._options = {}
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.asb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != ASB_OK:
            msg = self.lib.asb_last_error(self.handle)
            raise ArrowSpaceError(rc, msg.decode() if msg else self.lib.asb_status_string(rc).decode())

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.asb_kernel_launches(self.handle))

    def kernel_ms(self, which: str) -> float:
        return float(self.lib.asb_last_kernel_ms(self.handle, which.encode()))

    def set_option(self, key: str, value: float) -> None:
        self.check(self.lib.asb_ctx_set_option(self.handle, key.encode(), float(value)))
        self._options[key] = float(value)

    def get_option(self, key: str, default: float = 0.0) -> float:
        """The value last set through :meth:`set_option` (the library's default is the caller's to state)."""
        return self._options.get(key, float(default))

    # ---- thin wrappers over the stage entry points -------------------------------------------
    def twonn_distances(self, rows, sample_idx) -> Tuple[np.ndarray, np.ndarray]:
        rows = _as_f64_matrix(rows)
        n, f = _shape2(rows)
        si = np.ascontiguousarray(sample_idx, dtype=np.int64)
        d1 = np.empty(len(si), dtype=np.float64)
        d2 = np.empty(len(si), dtype=np.float64)
        self.check(self.lib.asb_twonn_distances(self.handle, _ptr(rows), n, f, _ptr(si), len(si), _ptr(d1), _ptr(d2)))
        return d1, d2

    def twonn_distances_sharded(self, comm: "Comm", rows_local, shard_offset: int, sample_idx):
        """``asb_twonn_distances_sharded``: ``sample_idx`` are GLOBAL row indices, the same on every rank."""
        rows_local = _as_f64_matrix(rows_local)
        n, f = _shape2(rows_local)
        si = np.ascontiguousarray(sample_idx, dtype=np.int64)
        d1 = np.empty(len(si), dtype=np.float64)
        d2 = np.empty(len(si), dtype=np.float64)
        self.check(self.lib.asb_twonn_distances_sharded(self.handle, comm.handle, _ptr(rows_local), n, f, int(shard_offset),
                                                        _ptr(si), len(si), _ptr(d1), _ptr(d2)))
        return d1, d2

    def cluster_incremental(self, rows, max_clusters: int, radius: float):
        rows = _as_f64_matrix(rows)
        n, f = _shape2(rows)
        cent = np.zeros((max_clusters, f), dtype=np.float64)
        asg = np.empty(n, dtype=np.int64)
        sizes = np.zeros(max_clusters, dtype=np.uint64)
        x = _I64(0)
        self.check(self.lib.asb_cluster_incremental(self.handle, _ptr(rows), n, f, int(max_clusters), float(radius),
                                                    _ptr(cent), _ptr(asg), _ptr(sizes), C.byref(x)))
        return cent[: x.value].copy(), asg, sizes[: x.value].copy()

    def cluster_incremental_resume(self, rows, max_clusters: int, radius: float, centroids, sizes, x: int,
                                   assignments=None):
        """Continue the walk from (centroids[:x], sizes[:x]); buffers are updated in place.
        Returns (x_new, assignments)."""
        rows = _as_f64_matrix(rows)
        n, f = _shape2(rows)
        if assignments is None:
            if _is_device(rows):
                import torch
                assignments = torch.empty(n, dtype=torch.int64, device=rows.device)
            else:
                assignments = np.empty(n, dtype=np.int64)
        xx = _I64(int(x))
        self.check(self.lib.asb_cluster_incremental_resume(self.handle, _ptr(rows), n, f, int(max_clusters),
                                                           float(radius), _ptr(centroids), _ptr(assignments),
                                                           _ptr(sizes), C.byref(xx)))
        return xx.value, assignments

    def build_feature_laplacian(self, centroids, gp: "GraphParams"):
        centroids = _as_f64_matrix(centroids)
        x, f = _shape2(centroids)
        cap = max(int(self.lib.asb_laplacian_max_nnz(f, gp.topk)), f)
        indptr = np.zeros(f + 1, dtype=np.int64)
        indices = np.zeros(cap, dtype=np.int64)
        data = np.zeros(cap, dtype=np.float64)
        nnz = _I64(0)
        gpc = gp.to_c()
        self.check(self.lib.asb_build_feature_laplacian(self.handle, _ptr(centroids), x, f, C.byref(gpc), _ptr(indptr),
                                                        _ptr(indices), _ptr(data), cap, C.byref(nnz)))
        return indptr, indices[: nnz.value].copy(), data[: nnz.value].copy()

    def compute_taumode(self, items, csr, taumode: "TauMode", want_norms: bool = False, out=None):
        items = _as_f64_matrix(items)
        n, f = _shape2(items)
        indptr, indices, data = csr
        if out is None:
            if _is_device(items):
                import torch
                lam = torch.empty(n, dtype=torch.float64, device=items.device)
                n2 = torch.empty(n, dtype=torch.float64, device=items.device) if want_norms else None
            else:
                lam = np.empty(n, dtype=np.float64)
                n2 = np.empty(n, dtype=np.float64) if want_norms else None
        else:
            lam, n2 = out
        stats = np.zeros(3, dtype=np.float64)
        self.check(self.lib.asb_compute_taumode(self.handle, _ptr(items), n, f, _ptr(indptr), _ptr(indices), _ptr(data),
                                                taumode.mode, taumode.value, _ptr(lam), _ptr(n2), _ptr(stats)))
        return lam, n2, stats

    def prepare_query_lambdas(self, queries, csr, taumode: "TauMode"):
        queries = _as_f64_matrix(queries)
        nq, f = _shape2(queries)
        indptr, indices, data = csr
        if _is_device(queries):
            import torch
            lq = torch.empty(nq, dtype=torch.float64, device=queries.device)
        else:
            lq = np.empty(nq, dtype=np.float64)
        self.check(self.lib.asb_prepare_query_lambdas(self.handle, _ptr(queries), nq, f, _ptr(indptr), _ptr(indices),
                                                      _ptr(data), taumode.mode, taumode.value, _ptr(lq)))
        return lq

    def search_lambda_aware_batch(self, items, lambdas, queries, lambda_q, k: int, alpha: float, norms2=None,
                                  index_offset: int = 0, out=None):
        items = _as_f64_matrix(items)
        queries = _as_f64_matrix(queries)
        n, f = _shape2(items)
        nq, fq = _shape2(queries)
        if fq != f:
            raise ArrowSpaceError(ASB_ERR_DIM, f"Query dimension {fq} doesn't match index original dimension {f}")
        if out is None:
            if _is_device(queries):
                import torch
                idx = torch.empty((nq, max(k, 1)), dtype=torch.int64, device=queries.device)
                score = torch.empty((nq, max(k, 1)), dtype=torch.float64, device=queries.device)
                count = torch.empty(nq, dtype=torch.int64, device=queries.device)
            else:
                idx = np.full((nq, max(k, 1)), -1, dtype=np.int64)
                score = np.zeros((nq, max(k, 1)), dtype=np.float64)
                count = np.zeros(nq, dtype=np.int64)
        else:
            idx, score, count = out
        self.check(self.lib.asb_search_lambda_aware_batch(self.handle, _ptr(items), _ptr(lambdas), _ptr(norms2), n, f,
                                                          _ptr(queries), _ptr(lambda_q), nq, int(k), float(alpha),
                                                          int(index_offset), _ptr(idx), _ptr(score), _ptr(count)))
        return idx, score, count

    def project_matrix(self, rows, projection):
        """``project_matrix`` (src/reduction.rs:143-166) with the Gaussian matrix materialised by the caller:
        ``projection`` is F x r, drawn in the reference's order (feature-major).  Returns n x r (bit-identical to
        the reference's accumulation order)."""
        rows = rows if _is_device(rows) else _as_f64_matrix(rows)
        projection = projection if _is_device(projection) else _as_f64_matrix(projection)
        n, f = _shape2(rows)
        f2, r = _shape2(projection)
        if f2 != f:
            raise ArrowSpaceError(ASB_ERR_DIM, f"projection has {f2} rows, items have {f} features")
        if _is_device(rows):
            import torch
            out = torch.empty((n, r), dtype=torch.float64, device=rows.device)
        else:
            out = np.empty((n, r), dtype=np.float64)
        self.check(self.lib.asb_project_matrix(self.handle, _ptr(rows), n, f, _ptr(projection), r, _ptr(out)))
        return out

    def search_energy_batch(self, items, lambdas, queries, lambda_q, k: int, w_lambda: float, w_dirichlet: float,
                            norms2=None, index_offset: int = 0):
        """``EnergyMaps::search_energy`` (src/energymaps.rs:368-407) for a batch of queries: (index, -energy),
        best first.  Returns (idx[nq,k], score[nq,k], count[nq])."""
        n, f = _shape2(items)
        nq = _shape2(queries)[0]
        idx = np.full((nq, max(k, 1)), -1, dtype=np.int64)
        score = np.zeros((nq, max(k, 1)), dtype=np.float64)
        count = np.zeros(nq, dtype=np.int64)
        keep = [_as_f64_matrix(items) if not _is_device(items) else items,
                np.ascontiguousarray(lambdas, dtype=np.float64) if not _is_device(lambdas) else lambdas,
                _as_f64_matrix(queries) if not _is_device(queries) else queries,
                np.ascontiguousarray(lambda_q, dtype=np.float64) if not _is_device(lambda_q) else lambda_q]
        self.check(self.lib.asb_search_energy_batch(self.handle, _ptr(keep[0]), _ptr(keep[1]),
                                                    _ptr(norms2) if norms2 is not None else None, n, f, _ptr(keep[2]),
                                                    _ptr(keep[3]), nq, int(k), float(w_lambda), float(w_dirichlet),
                                                    int(index_offset), _ptr(idx), _ptr(score), _ptr(count)))
        return idx[:, :k], score[:, :k], count

    def search_lambda_aware_hybrid_batch(self, items, lambdas, queries, lambda_q, k: int, alpha: float, norms2=None):
        items = _as_f64_matrix(items)
        queries = _as_f64_matrix(queries)
        n, f = _shape2(items)
        nq, fq = _shape2(queries)
        if fq != f:
            raise ArrowSpaceError(ASB_ERR_DIM, f"Query dimension {fq} doesn't match index original dimension {f}")
        idx = np.full((nq, max(k, 1)), -1, dtype=np.int64)
        score = np.zeros((nq, max(k, 1)), dtype=np.float64)
        count = np.zeros(nq, dtype=np.int64)
        self.check(self.lib.asb_search_lambda_aware_hybrid_batch(
            self.handle, _ptr(items), _ptr(lambdas), _ptr(norms2), n, f, _ptr(queries), _ptr(lambda_q), nq, int(k),
            float(alpha), _ptr(idx), _ptr(score), _ptr(count)))
        return idx, score, count

    def range_search(self, lambdas, lambda_q: float, eps: float, index_offset: int = 0):
        n = int(lambdas.shape[0])
        idx = np.empty(n, dtype=np.int64)
        dist = np.empty(n, dtype=np.float64)
        cnt = _I64(0)
        self.check(self.lib.asb_range_search(self.handle, _ptr(lambdas), n, float(lambda_q), float(eps),
                                             int(index_offset), _ptr(idx), _ptr(dist), n, C.byref(cnt)))
        return idx[: cnt.value].copy(), dist[: cnt.value].copy()

    def topk_merge(self, in_score, in_idx, parts: int, nq: int, k: int):
        if _is_device(in_score):
            import torch
            os_ = torch.empty((nq, k), dtype=torch.float64, device=in_score.device)
            oi = torch.empty((nq, k), dtype=torch.int64, device=in_score.device)
            oc = torch.empty(nq, dtype=torch.int64, device=in_score.device)
        else:
            in_score = np.ascontiguousarray(in_score, dtype=np.float64)
            in_idx = np.ascontiguousarray(in_idx, dtype=np.int64)
            os_ = np.empty((nq, k), dtype=np.float64)
            oi = np.empty((nq, k), dtype=np.int64)
            oc = np.empty(nq, dtype=np.int64)
        self.check(self.lib.asb_topk_merge(self.handle, _ptr(in_score), _ptr(in_idx), parts, nq, k, _ptr(os_), _ptr(oi),
                                           _ptr(oc)))
        return os_, oi, oc


_default_ctx: Optional[Context] = None


class Comm:
    """One rank of a row-sharded job (``asb_comm``; one process per GPU, NCCL over NVLink inside the library).

    ``Comm.make_unique_id(ctx)`` on rank 0 returns the 128 bytes of an ``ncclUniqueId``; ship them to the other ranks by
    any means (``torch.distributed.broadcast_object_list``, a file, MPI ...) and construct ``Comm(ctx, nranks, rank, id)``
    everywhere.  Every ``*_sharded`` call is collective."""

    def __init__(self, ctx: "Context", nranks: int, rank: int, unique_id: bytes):
        self.ctx = ctx
        h = _P()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        ctx.check(ctx.lib.asb_comm_init_rank(ctx.handle, buf, int(nranks), int(rank), C.byref(h)))
        self.handle = h
        self.rank, self.size = int(rank), int(nranks)

    @staticmethod
    def make_unique_id(ctx: "Context") -> bytes:
        buf = C.create_string_buffer(128)
        ctx.check(ctx.lib.asb_comm_unique_id(ctx.handle, buf))
        return bytes(buf.raw)

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.asb_comm_destroy(self.handle)
            self.handle = None


class ShardedIndex:
    """``asb_index_build_sharded``: stages 1-3 of ``ArrowSpaceBuilder::build`` (src/builder.rs:249-455) over a row-sharded
    dataset, every intermediate resident in HBM.  ``rows_local``: this rank's rows (torch CUDA tensor or host array),
    global rows ``[shard_offset, shard_offset + n_local)``.  Results equal a single-GPU build of the concatenated rows:
    centroids bit-identical, assignments / Laplacian identical, lambdas of the local rows."""

    def __init__(self, ctx: "Context", comm: Comm, rows_local, shard_offset: int, n_global: int, params: BuildParamsC):
        self.ctx, self.comm = ctx, comm
        rows_local = _as_f64_matrix(rows_local)
        self.rows = rows_local              # borrowed by the native index when it is a device tensor
        n, f = _shape2(rows_local)
        self.n_local, self.f, self.offset, self.n_global = n, f, int(shard_offset), int(n_global)
        h = _P()
        ctx.check(ctx.lib.asb_index_build_sharded(ctx.handle, comm.handle, _ptr(rows_local), n, f, int(shard_offset),
                                                  int(n_global), C.byref(params), C.byref(h)))
        self.handle = h
        self.params = params

    def info(self) -> IndexInfoC:
        info = IndexInfoC()
        self.ctx.check(self.ctx.lib.asb_index_info_get(self.handle, C.byref(info)))
        return info

    def lambdas(self) -> np.ndarray:
        out = np.empty(self.n_local, dtype=np.float64)
        self.ctx.check(self.ctx.lib.asb_index_lambdas(self.ctx.handle, self.handle, _ptr(out)))
        return out

    def assignments(self) -> np.ndarray:
        out = np.empty(self.n_local, dtype=np.int64)
        self.ctx.check(self.ctx.lib.asb_index_assignments(self.ctx.handle, self.handle, _ptr(out)))
        return out

    def centroids(self) -> np.ndarray:
        x = int(self.info().n_clusters)
        out = np.empty((x, self.f), dtype=np.float64)
        self.ctx.check(self.ctx.lib.asb_index_centroids(self.ctx.handle, self.handle, _ptr(out)))
        return out

    def cluster_sizes(self) -> np.ndarray:
        out = np.empty(int(self.info().n_clusters), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.asb_index_cluster_sizes(self.ctx.handle, self.handle, _ptr(out)))
        return out

    def laplacian(self):
        nnz = int(self.info().nnz)
        ip, ii, dd = np.empty(self.f + 1, dtype=np.int64), np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.float64)
        self.ctx.check(self.ctx.lib.asb_index_laplacian(self.ctx.handle, self.handle, _ptr(ip), _ptr(ii), _ptr(dd)))
        return ip, ii, dd

    def search(self, queries, k: int, alpha: float):
        """``EigenMaps::search`` for a batch over all shards: merged (idx, score, count, lambda_q) on every rank."""
        queries = _as_f64_matrix(queries)
        nq, fq = _shape2(queries)
        if fq != self.f:
            raise ArrowSpaceError(ASB_ERR_DIM, f"Query dimension {fq} doesn't match index original dimension {self.f}")
        if _is_device(queries):
            import torch
            idx = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
            score = torch.empty((nq, k), dtype=torch.float64, device=queries.device)
            count = torch.empty(nq, dtype=torch.int64, device=queries.device)
            lq = torch.empty(nq, dtype=torch.float64, device=queries.device)
        else:
            idx, score = np.empty((nq, k), dtype=np.int64), np.empty((nq, k), dtype=np.float64)
            count, lq = np.empty(nq, dtype=np.int64), np.empty(nq, dtype=np.float64)
        self.ctx.check(self.ctx.lib.asb_index_search_sharded(self.ctx.handle, self.comm.handle, self.handle, _ptr(queries), nq,
                                                             int(k), float(alpha), _ptr(idx), _ptr(score), _ptr(count),
                                                             _ptr(lq)))
        return idx, score, count, lq

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.asb_index_destroy(self.handle)
            self.handle = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ------------------------------------------------------------------------------- TauMode
@dataclass(frozen=True)
class TauMode:
    """``TauMode`` (src/taumode.rs:75-82): Fixed(f64) | Median (default) | Mean | Percentile(f64)."""
    mode: int = TAU_MEDIAN
    value: float = 0.0

    @staticmethod
    def Fixed(t: float) -> "TauMode":
        return TauMode(TAU_FIXED, float(t))

    @staticmethod
    def Percentile(p: float) -> "TauMode":
        return TauMode(TAU_PERCENTILE, float(p))

    def __str__(self):  # src/taumode.rs:663-672
        return {TAU_FIXED: f"Fixed({self.value})", TAU_MEDIAN: "Median", TAU_MEAN: "Mean",
                TAU_PERCENTILE: f"Percentile({self.value})"}[self.mode]


TauMode.Median = TauMode(TAU_MEDIAN, 0.0)
TauMode.Mean = TauMode(TAU_MEAN, 0.0)
TAUDEFAULT = TauMode.Median  # src/core.rs:387
TAU_FLOOR = 1e-10            # src/taumode.rs:84


# --------------------------------------------------------------------------- graph types
@dataclass
class GraphParams:
    """``GraphParams`` (src/graph.rs:94-102)."""
    eps: float
    k: int
    topk: int
    p: float
    sigma: Optional[float]
    normalise: bool = False
    sparsity_check: bool = False
    # smartcore CosinePair semantics that the reference's tests do not pin
    self_included: bool = False
    rectified: bool = False

    def to_c(self) -> GraphParamsC:
        return GraphParamsC(self.eps, self.k, self.topk, self.p, 1 if self.sigma is not None else 0,
                            self.sigma if self.sigma is not None else 0.0, int(self.normalise),
                            int(self.sparsity_check), int(self.self_included), int(self.rectified))


@dataclass
class GraphLaplacian:
    """``GraphLaplacian`` (src/graph.rs:127-135): ``matrix`` is the F x F CSR, ``nnodes`` = N."""
    indptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray
    nnodes: int
    graph_params: GraphParams
    init_data: Optional[np.ndarray] = None  # X x F centroids (the reference stores the transpose)

    @property
    def csr(self):
        return self.indptr, self.indices, self.data

    def shape(self) -> Tuple[int, int]:
        f = len(self.indptr) - 1
        return f, f

    def nnz(self) -> int:
        return int(self.indptr[-1])

    @staticmethod
    def sparsity(gl: "GraphLaplacian") -> float:  # src/graph.rs:572-578
        r, c = gl.shape()
        return 1.0 - gl.nnz() / float(r * c)

    def to_dense(self) -> np.ndarray:
        f = len(self.indptr) - 1
        m = np.zeros((f, f))
        for i in range(f):
            for e in range(self.indptr[i], self.indptr[i + 1]):
                m[i, self.indices[e]] = self.data[e]
        return m


class ImplicitProjection:
    """``ImplicitProjection`` (src/reduction.rs:168-199) with the matrix materialised: the reference keeps only a
    seed and redraws the F x r Gaussians (ChaCha8 + StandardNormal) on every call; those generators are third-party,
    so the caller supplies the matrix, drawn feature-major like the reference does."""

    def __init__(self, matrix, ctx: Optional["Context"] = None):
        self.matrix = _as_f64_matrix(matrix)
        self.original_dim, self.reduced_dim = self.matrix.shape
        self.ctx = ctx

    def project(self, query) -> np.ndarray:  # :180-199
        q = np.ascontiguousarray(query, dtype=np.float64).reshape(1, -1)[:, : self.original_dim]
        return np.asarray((self.ctx or default_context()).project_matrix(q, self.matrix))[0]


def project_matrix(data, projection: ImplicitProjection, ctx: Optional["Context"] = None):
    """``project_matrix`` (src/reduction.rs:143-166)."""
    return (ctx or projection.ctx or default_context()).project_matrix(data, projection.matrix)


def compute_jl_dimension(n_points: int, epsilon: float) -> int:
    """``compute_jl_dimension`` (src/reduction.rs:127-141)."""
    return int(load_library().asb_jl_dimension(int(n_points), float(epsilon)))


@dataclass
class ArrowItem:
    """``ArrowItem`` (src/core.rs:84-87)."""
    item: np.ndarray
    lambda_: float = 0.0

    @staticmethod
    def new(item, lambda_: float = 0.0) -> "ArrowItem":
        return ArrowItem(np.ascontiguousarray(item, dtype=np.float64), float(lambda_))


@dataclass
class ClusteredOutput:
    """``ClusteredOutput`` (src/eigenmaps.rs, returned by start_clustering)."""
    aspace: "ArrowSpace"
    centroids: np.ndarray
    reduced_dim: int
    n_items: int
    n_features: int


# ---------------------------------------------------------------------------- ArrowSpace
class ArrowSpace:
    """``ArrowSpace`` (src/core.rs:366-385) with the ``EigenMaps`` stages (src/eigenmaps.rs:174-456).

    ``data`` stays wherever the caller put it (numpy host array or torch CUDA tensor); after
    ``ArrowSpaceBuilder.build`` the items, lambdas and graph also live in HBM inside a native
    index handle so that searches do not re-transfer them.
    """

    def __init__(self, items, taumode: TauMode = TAUDEFAULT, ctx: Optional[Context] = None):
        items = _as_f64_matrix(items)
        n, f = _shape2(items) if len(items.shape) == 2 else (0, 0)
        if n == 0:
            raise ArrowSpaceError(ASB_ERR_EMPTY, "items cannot be empty")  # core.rs:416
        if n <= 1:
            raise ArrowSpaceError(ASB_ERR_EMPTY, "cannot create a arrowspace of one arrow only")  # :417-420
        self.ctx = ctx or default_context()
        self.nitems, self.nfeatures = n, f
        self.data = items
        self.lambdas = np.zeros(n, dtype=np.float64)
        self.norms2 = None
        self.taumode = taumode
        self.n_clusters = 0
        self.cluster_assignments: Optional[np.ndarray] = None
        self.cluster_sizes: Optional[np.ndarray] = None
        self.cluster_radius = 0.0
        self.projection_matrix = None
        self.reduced_dim = None
        self.signals = None  # (indptr, indices, data) of aspace.signals, src/core.rs:370
        self._index = None  # native asb_index*
        self._device_cache = None

    # -- index handle ---------------------------------------------------------------------
    def _release(self):
        if getattr(self, "_index", None) is not None:
            lib = getattr(self.ctx, "lib", None)
            if lib is not None:
                lib.asb_index_destroy(self._index)
            self._index = None

    def __del__(self):  # pragma: no cover
        try:
            self._release()
        except Exception:
            pass

    def index_info(self) -> IndexInfoC:
        info = IndexInfoC()
        if self._index is None:
            raise ArrowSpaceError(ASB_ERR_INVALID, "no native index: call ArrowSpaceBuilder.build first")
        self.ctx.check(self.ctx.lib.asb_index_info_get(self._index, C.byref(info)))
        return info

    # -- EigenMaps stages --------------------------------------------------------------------
    @staticmethod
    def start_clustering(builder: "ArrowSpaceBuilder", rows) -> ClusteredOutput:
        """``EigenMaps::start_clustering`` (src/eigenmaps.rs:175-290)."""
        rows = _as_f64_matrix(rows)
        aspace = ArrowSpace(rows, builder.synthesis, builder.ctx)
        if builder.sampling is not None:
            raise ArrowSpaceError(ASB_ERR_UNSUPPORTED,
                                  "inline sampling uses an OS-seeded RNG in the reference (src/sampling.rs:123,184); "
                                  "use with_inline_sampling(None)")
        if builder.use_dims_reduction:
            raise ArrowSpaceError(ASB_ERR_UNSUPPORTED, "JL projection is a 'next' row (SURVEY 8f); use "
                                                       "with_dims_reduction(False, None)")
        k_opt, radius = builder.resolve_cluster_params(rows)
        builder.cluster_max_clusters = k_opt   # eigenmaps.rs:216
        builder.cluster_radius = radius        # :217
        cent, asg, sizes = builder.ctx.cluster_incremental(rows, k_opt, radius)
        aspace.n_clusters = cent.shape[0]      # :242-245
        aspace.is_standin = bool(getattr(builder, "is_standin", False))
        aspace.cluster_assignments = asg
        aspace.cluster_sizes = sizes
        aspace.cluster_radius = radius
        return ClusteredOutput(aspace, cent, aspace.nfeatures, aspace.nitems, aspace.nfeatures)

    def eigenmaps(self, builder: "ArrowSpaceBuilder", centroids, n_items: int) -> GraphLaplacian:
        """``EigenMaps::eigenmaps`` (src/eigenmaps.rs:292-356)."""
        centroids = _as_f64_matrix(centroids)
        if centroids.shape[0] > n_items:  # graph.rs:168 assert
            raise ArrowSpaceError(ASB_ERR_INVALID, "clustered.shape().0 <= n_items violated")
        gp = builder.graph_params()
        indptr, indices, data = self.ctx.build_feature_laplacian(centroids, gp)
        gl = GraphLaplacian(indptr, indices, data, n_items, gp, np.asarray(centroids))
        if builder.prebuilt_spectral:  # src/eigenmaps.rs:325-345 -> GraphFactory::build_spectral_laplacian
            self.signals = self.build_spectral_laplacian(gl)
        return gl

    def build_spectral_laplacian(self, gl: GraphLaplacian):
        """``GraphFactory::build_spectral_laplacian`` (src/graph.rs:211-231): the Laplacian construction run on
        ``dense(L)^T`` -- the F rows of L are the "items", its F columns the nodes.  Returns the signals CSR."""
        return self.ctx.build_feature_laplacian(np.ascontiguousarray(gl.to_dense()), gl.graph_params)

    def compute_taumode(self, gl: GraphLaplacian) -> None:
        """``EigenMaps::compute_taumode`` (src/eigenmaps.rs:358-383); reads ``signals`` instead of the feature
        Laplacian when they exist (src/taumode.rs:195-200)."""
        graph = self.signals if getattr(self, "signals", None) is not None else gl.csr
        lam, n2, stats = self.ctx.compute_taumode(self.data, graph, self.taumode, want_norms=True)
        self.lambdas = lam
        self.norms2 = n2
        self.lambda_stats = (stats[0], stats[1], stats[2] / self.nitems)

    def prepare_query_item(self, item, gl: GraphLaplacian) -> float:
        """``ArrowSpace::prepare_query_item`` (src/core.rs:533-549)."""
        q = np.ascontiguousarray(item, dtype=np.float64).reshape(1, -1)
        if q.shape[1] != self.nfeatures:
            raise ArrowSpaceError(ASB_ERR_DIM, f"Query dimension {q.shape[1]} doesn't match index original dimension "
                                               f"{self.nfeatures}")
        if self.projection_matrix is not None:      # :540-545: the query is projected first
            if not np.all(np.isfinite(q)):           # :534-537 looks at the raw query
                raise ArrowSpaceError(ASB_ERR_NONFINITE_QUERY, "Query item contains invalid values (NaN or infinity). "
                                                               "All values must be finite.")
            q = np.ascontiguousarray(self.projection_matrix.project(q[0]).reshape(1, -1))
        return float(self.ctx.prepare_query_lambdas(q, gl.csr, self.taumode)[0])

    def prepare_query_items(self, queries, gl: GraphLaplacian):
        """Batched ``prepare_query_item``."""
        return self.ctx.prepare_query_lambdas(queries, gl.csr, self.taumode)

    def _device_items(self):
        """(items, lambdas, norms2) already resident in HBM when torch CUDA tensors were given."""
        return self.data, self.lambdas, self.norms2

    def search_lambda_aware(self, query: ArrowItem, k: int, alpha: float) -> List[Tuple[int, float]]:
        """``ArrowSpace::search_lambda_aware`` (src/core.rs:760-798)."""
        q = np.ascontiguousarray(query.item, dtype=np.float64).reshape(1, -1)
        idx, score, count = self.search_lambda_aware_batch(q, np.array([query.lambda_], dtype=np.float64), k, alpha)
        c = int(np.asarray(count)[0])
        return [(int(idx[0][r]), float(score[0][r])) for r in range(c)]

    def search_lambda_aware_batch(self, queries, lambda_q, k: int, alpha: float):
        """The reference's "batched" search is a loop over queries
        (benches/index_compute_bench.rs:250-262); here one fused launch."""
        if not _is_device(queries):
            lambda_q = np.ascontiguousarray(lambda_q, dtype=np.float64)
        if self._index is not None:  # items + lambdas already resident in HBM
            queries = _as_f64_matrix(queries)
            nq, fq = _shape2(queries)
            if fq != self.nfeatures:
                raise ArrowSpaceError(ASB_ERR_DIM, f"Query dimension {fq} doesn't match index original dimension "
                                                   f"{self.nfeatures}")
            if _is_device(queries):
                import torch
                idx = torch.empty((nq, max(k, 1)), dtype=torch.int64, device=queries.device)
                score = torch.empty((nq, max(k, 1)), dtype=torch.float64, device=queries.device)
                count = torch.empty(nq, dtype=torch.int64, device=queries.device)
            else:
                idx = np.full((nq, max(k, 1)), -1, dtype=np.int64)
                score = np.zeros((nq, max(k, 1)), dtype=np.float64)
                count = np.zeros(nq, dtype=np.int64)
            self.ctx.check(self.ctx.lib.asb_index_search_lambda_aware(
                self.ctx.handle, self._index, _ptr(queries), _ptr(lambda_q), nq, int(k), float(alpha), _ptr(idx),
                _ptr(score), _ptr(count)))
            return idx, score, count
        items, lambdas, norms2 = self._device_items()
        return self.ctx.search_lambda_aware_batch(items, lambdas, queries, lambda_q, k, alpha, norms2=norms2)

    def search_lambda_aware_hybrid(self, query: ArrowItem, k: int, alpha: float) -> List[Tuple[int, float]]:
        """``ArrowSpace::search_lambda_aware_hybrid`` (src/core.rs:802-928)."""
        q = np.ascontiguousarray(query.item, dtype=np.float64).reshape(1, -1)
        items, lambdas, norms2 = self._device_items()
        idx, score, count = self.ctx.search_lambda_aware_hybrid_batch(
            items, lambdas, q, np.array([query.lambda_], dtype=np.float64), k, alpha, norms2=norms2)
        return [(int(idx[0][r]), float(score[0][r])) for r in range(int(count[0]))]

    def search_energy(self, query, gl_energy: GraphLaplacian, k: int, w_lambda: float,
                      w_dirichlet: float) -> List[Tuple[int, float]]:
        """``EnergyMaps::search_energy`` (src/energymaps.rs:368-407): lambda_q from ``prepare_query_item`` on
        ``gl_energy`` (:885), then (index, -energy) for the k items of least projected energy.  Spaces with a JL
        projection or spectral signals take other branches of the reference's score (:858-882) -- unsupported."""
        q = np.ascontiguousarray(query.item if isinstance(query, ArrowItem) else query, dtype=np.float64)
        if self.projection_matrix is not None or self.signals is not None:
            # project_vec through the projection, projected_dirichlet through the signals (:858-882): both live in the
            # native index (the space must come from ArrowSpaceBuilder.build)
            if self._index is None:
                raise ArrowSpaceError(ASB_ERR_UNSUPPORTED, "search_energy with a projection / signals needs the native "
                                                           "index of ArrowSpaceBuilder.build")
            idx, score, count = self.search_energy_batch(q.reshape(1, -1), k, w_lambda, w_dirichlet)
            return [(int(idx[0, r]), float(score[0, r])) for r in range(int(count[0]))]
        lq = self.prepare_query_item(q, gl_energy)
        idx, score, count = self.ctx.search_energy_batch(self.data, self.lambdas, q.reshape(1, -1), np.array([lq]), k,
                                                         w_lambda, w_dirichlet)
        return [(int(idx[0, r]), float(score[0, r])) for r in range(int(count[0]))]

    def search_energy_batch(self, queries, k: int, w_lambda: float, w_dirichlet: float):
        """Batched ``search_energy`` against the native index, every branch of ``ProjectedEnergy::score``."""
        if self._index is None:
            raise ArrowSpaceError(ASB_ERR_INVALID, "no native index: call ArrowSpaceBuilder.build first")
        queries = _as_f64_matrix(queries)
        nq, fq = _shape2(queries)
        if fq != self.nfeatures:
            raise ArrowSpaceError(ASB_ERR_DIM, f"Query dimension {fq} doesn't match index original dimension "
                                               f"{self.nfeatures}")
        idx = np.full((nq, k), -1, dtype=np.int64)
        score = np.zeros((nq, k), dtype=np.float64)
        count = np.zeros(nq, dtype=np.int64)
        self.ctx.check(self.ctx.lib.asb_index_search_energy(self.ctx.handle, self._index, _ptr(queries), nq, int(k),
                                                            float(w_lambda), float(w_dirichlet), _ptr(idx), _ptr(score),
                                                            _ptr(count)))
        return idx, score, count

    def prepare_query_items_index(self, queries):
        """``prepare_query_item`` for a batch against the native index (queries are projected first when the build
        used a JL projection, src/core.rs:540-545)."""
        if self._index is None:
            raise ArrowSpaceError(ASB_ERR_INVALID, "no native index: call ArrowSpaceBuilder.build first")
        queries = _as_f64_matrix(queries)
        nq = _shape2(queries)[0]
        lq = np.empty(nq, dtype=np.float64)
        self.ctx.check(self.ctx.lib.asb_index_prepare_query(self.ctx.handle, self._index, _ptr(queries), nq, _ptr(lq)))
        return lq

    def range_search(self, query: ArrowItem, gl: GraphLaplacian, eps: float) -> List[Tuple[int, float]]:
        """``ArrowSpace::range_search`` (src/core.rs:944-976): the query lambda is re-prepared when it is
        (relatively) zero, then every item with ``lambda_q - lambda_i <= eps`` is returned in index order."""
        lam_q = query.lambda_
        if abs(lam_q) <= 1e-9:  # relative_eq!(query.lambda, 0.0, epsilon = 1e-9), :953
            lam_q = self.prepare_query_item(query.item, gl)
        idx, dist = self.ctx.range_search(self.lambdas, lam_q, eps)
        return [(int(i), float(d)) for i, d in zip(idx, dist)]

    def search(self, item, gl: GraphLaplacian, k: int, alpha: float) -> List[Tuple[int, float]]:
        """``EigenMaps::search`` (src/eigenmaps.rs:410-455) = prepare_query_item + search_lambda_aware."""
        lam = self.prepare_query_item(item, gl)
        return self.search_lambda_aware(ArrowItem.new(item, lam), k, alpha)

    def search_batch(self, queries, k: int, alpha: float):
        """Batched ``EigenMaps::search`` against the native (HBM-resident) index."""
        if self._index is None:
            raise ArrowSpaceError(ASB_ERR_INVALID, "no native index: call ArrowSpaceBuilder.build first")
        queries = _as_f64_matrix(queries)
        nq, fq = _shape2(queries)
        if fq != self.nfeatures:
            raise ArrowSpaceError(ASB_ERR_DIM, f"Query dimension {fq} doesn't match index original dimension "
                                               f"{self.nfeatures}")
        if _is_device(queries):
            import torch
            idx = torch.empty((nq, max(k, 1)), dtype=torch.int64, device=queries.device)
            score = torch.empty((nq, max(k, 1)), dtype=torch.float64, device=queries.device)
            count = torch.empty(nq, dtype=torch.int64, device=queries.device)
            lq = torch.empty(nq, dtype=torch.float64, device=queries.device)
        else:
            idx = np.full((nq, max(k, 1)), -1, dtype=np.int64)
            score = np.zeros((nq, max(k, 1)), dtype=np.float64)
            count = np.zeros(nq, dtype=np.int64)
            lq = np.zeros(nq, dtype=np.float64)
        self.ctx.check(self.ctx.lib.asb_index_search(self.ctx.handle, self._index, _ptr(queries), nq, int(k),
                                                     float(alpha), _ptr(idx), _ptr(score), _ptr(count), _ptr(lq)))
        return idx, score, count, lq

    def lambdas_host(self) -> np.ndarray:
        lam = self.lambdas
        return lam.cpu().numpy() if hasattr(lam, "cpu") else np.asarray(lam)


# --------------------------------------------------------------------- ArrowSpaceBuilder
class ArrowSpaceBuilder:
    """``ArrowSpaceBuilder`` (src/builder.rs:20-57), defaults from ``Default`` (:59-91)."""

    def __init__(self, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.prebuilt_spectral = False
        self.lambda_eps = 1e-3
        self.lambda_k = 6
        self.lambda_topk = 3
        self.lambda_p = 2.0
        self.lambda_sigma: Optional[float] = None
        self.normalise = False
        self.sparsity_check = False
        self.sampling = "Simple(0.6)"  # SamplerType::Simple(0.6); only None is supported here
        self.synthesis = TAUDEFAULT
        self.cluster_max_clusters: Optional[int] = None
        self.cluster_radius = 1.0
        self.clustering_seed: Optional[int] = None
        self.deterministic_clustering = False
        self.use_dims_reduction = False
        self.rp_eps = 0.3
        self.projection = None
        self._explicit_cluster_params = False
        self.self_included = False
        self.rectified = False

    @staticmethod
    def new(ctx: Optional[Context] = None) -> "ArrowSpaceBuilder":
        return ArrowSpaceBuilder(ctx)

    def with_lambda_graph(self, eps: float, k: int, topk: int, p: float, sigma_override: Optional[float]):
        """src/builder.rs:109-137 (five arguments)."""
        self.lambda_eps, self.lambda_k, self.lambda_topk = float(eps), int(k), int(topk)
        self.lambda_p, self.lambda_sigma = float(p), sigma_override
        return self

    def with_synthesis(self, tau_mode: TauMode):  # :142-146
        self.synthesis = tau_mode
        return self

    def with_normalisation(self, normalise: bool):  # :148-152
        self.normalise = bool(normalise)
        return self

    def with_spectral(self, compute_spectral: bool):  # :157-162
        self.prebuilt_spectral = bool(compute_spectral)
        return self

    def with_sparsity_check(self, sparsity_check: bool):  # :164-168
        self.sparsity_check = bool(sparsity_check)
        return self

    def with_inline_sampling(self, sampling):  # :170-179
        self.sampling = sampling
        return self

    def with_dims_reduction(self, enable: bool, eps: Optional[float]):  # :181-185
        self.use_dims_reduction = bool(enable)
        self.rp_eps = 0.5 if eps is None else float(eps)
        return self

    def with_projection(self, projection):
        """The materialised Gaussian matrix of the build's ``ImplicitProjection`` (F x r, feature-major: row j holds the
        r samples drawn for feature j).  The reference keeps an 8-byte seed and redraws the matrix with ChaCha8 +
        StandardNormal (src/reduction.rs:176-199) -- third-party generators -- so a Rust host draws it once and hands it
        over; without Rust any matrix gives a self-consistent JL build.  Either an :class:`ImplicitProjection`, an array,
        or a callable ``(F, r) -> array`` evaluated once r is known (r = min(compute_jl_dimension(n_clusters, eps), F / 2),
        src/eigenmaps.rs:249-250)."""
        self.projection = projection
        return self

    def with_seed(self, seed: int):  # :190-195 -> deterministic (sequential) clustering
        self.clustering_seed = int(seed)
        self.deterministic_clustering = True
        return self

    def with_cluster_params(self, max_clusters: int, radius: float):
        """Hand over the output of the HOST heuristic ``compute_optimal_k``
        (src/clustering.rs:36-72: smartcore k-means + StdRng, stays in Rust): ``(k_opt, radius)``."""
        self.cluster_max_clusters = int(max_clusters)
        self.cluster_radius = float(radius)
        self._explicit_cluster_params = True
        return self

    def with_knn_semantics(self, self_included: bool = False, rectified: bool = False):
        """Switches for the un-vendored smartcore CosinePair behaviour ("unpinned")."""
        self.self_included, self.rectified = bool(self_included), bool(rectified)
        return self

    def define_result_k(self):  # :225-233
        if self.lambda_k <= 5:
            self.lambda_topk = 3
        elif self.lambda_k < 10:
            self.lambda_topk = 4

    def graph_params(self) -> GraphParams:
        return GraphParams(self.lambda_eps, self.lambda_k, self.lambda_topk, self.lambda_p, self.lambda_sigma,
                           self.normalise, self.sparsity_check, self.self_included, self.rectified)

    def resolve_cluster_params(self, rows) -> Tuple[int, float]:
        """(k_opt, radius).  The reference derives them with ``compute_optimal_k`` (src/clustering.rs:36-72: a
        Calinski-Harabasz search over smartcore KMeans seeded through rand::StdRng) -- third-party code that cannot be
        reproduced outside Rust, so a Rust host keeps calling the reference's own function and passes the pair in
        (``with_cluster_params``).  Without them this Python mirror falls back to the documented STAND-IN
        (heuristics.compute_optimal_k_standin): the build is then self-consistent but NOT the reference's build of the
        same rows; ``is_standin`` is set on the returned space and a warning is emitted once per builder."""
        if self._explicit_cluster_params:
            self.is_standin = False
            return int(self.cluster_max_clusters), float(self.cluster_radius)
        import warnings
        from .heuristics import compute_optimal_k_standin
        k_opt, radius, _ = compute_optimal_k_standin(self.ctx, rows, self.clustering_seed)
        if not getattr(self, "is_standin", False):
            warnings.warn("ArrowSpaceBuilder: (max_clusters, radius) come from the stand-in heuristic, not from the "
                          "reference's compute_optimal_k (smartcore KMeans + StdRng): results differ from arrowspace-rs on "
                          "the same rows; pass with_cluster_params(k_opt, radius) for a drop-in build", RuntimeWarning,
                          stacklevel=3)
        self.is_standin = True
        return k_opt, radius

    def build(self, rows) -> Tuple[ArrowSpace, GraphLaplacian]:
        """``ArrowSpaceBuilder::build`` (src/builder.rs:249-455): stages 1-3 in one native call,
        every intermediate resident in HBM."""
        rows = _as_f64_matrix(rows)
        if len(rows.shape) != 2 or rows.shape[0] == 0:
            raise ArrowSpaceError(ASB_ERR_EMPTY, "items cannot be empty")
        self.define_result_k()  # :255
        if self.sampling is not None:
            raise ArrowSpaceError(ASB_ERR_UNSUPPORTED,
                                  "inline sampling uses an OS-seeded RNG in the reference (src/sampling.rs:123,184); "
                                  "use with_inline_sampling(None)")
        n, f = _shape2(rows)
        aspace = ArrowSpace(rows, self.synthesis, self.ctx)
        k_opt, radius = self.resolve_cluster_params(rows)
        self.cluster_max_clusters, self.cluster_radius = k_opt, radius
        gp = self.graph_params()
        proj_mat, r = None, 0
        if self.use_dims_reduction and f > 64:                      # src/eigenmaps.rs:248
            if not self._explicit_cluster_params:
                raise ArrowSpaceError(ASB_ERR_UNSUPPORTED, "with_dims_reduction needs with_cluster_params: the target "
                                                           "dimension follows the cluster count (src/eigenmaps.rs:249)")
            if self.projection is None:
                raise ArrowSpaceError(ASB_ERR_UNSUPPORTED,
                                      "with_dims_reduction(true, eps) needs the materialised projection matrix "
                                      "(with_projection): the reference's seed-driven generators are third-party")
            # the reference sizes the projection by the number of clusters actually formed; the walk is deterministic,
            # so one clustering call tells it (the build below repeats the walk: same bits)
            cent0, _, _ = self.ctx.cluster_incremental(rows, int(k_opt), float(radius))
            target = min(compute_jl_dimension(int(cent0.shape[0]), self.rp_eps), f // 2)   # :249-250
            if target < f:                                           # :252
                pm = self.projection
                if callable(pm):
                    pm = pm(f, target)
                if isinstance(pm, ImplicitProjection):
                    pm = pm.matrix
                proj_mat = np.ascontiguousarray(pm, dtype=np.float64)
                if proj_mat.shape != (f, target):
                    raise ArrowSpaceError(ASB_ERR_DIM, f"projection matrix must be {f} x {target}, got {proj_mat.shape}")
                r = target
        bp = BuildParamsC(gp.to_c(), self.synthesis.mode, self.synthesis.value, int(k_opt), float(radius), 0,
                          1 if self.prebuilt_spectral else 0, _ptr(proj_mat) if proj_mat is not None else None, int(r))
        h = _P()
        lib = self.ctx.lib
        self.ctx.check(lib.asb_index_build(self.ctx.handle, _ptr(rows), n, f, C.byref(bp), C.byref(h)))
        aspace._index = h
        if r:
            aspace.projection_matrix = ImplicitProjection(proj_mat, self.ctx)   # eigenmaps.rs:260-261
            aspace.reduced_dim = r
        info = aspace.index_info()
        x, nnz = int(info.n_clusters), int(info.nnz)
        fg = r if r else f                                           # nodes of the feature graph
        lam = np.empty(n, dtype=np.float64)
        cent = np.empty((x, f), dtype=np.float64)
        asg = np.empty(n, dtype=np.int64)
        sizes = np.empty(x, dtype=np.uint64)
        indptr = np.empty(fg + 1, dtype=np.int64)
        indices = np.empty(nnz, dtype=np.int64)
        data = np.empty(nnz, dtype=np.float64)
        self.ctx.check(lib.asb_index_lambdas(self.ctx.handle, h, _ptr(lam)))
        self.ctx.check(lib.asb_index_centroids(self.ctx.handle, h, _ptr(cent)))
        self.ctx.check(lib.asb_index_assignments(self.ctx.handle, h, _ptr(asg)))
        self.ctx.check(lib.asb_index_cluster_sizes(self.ctx.handle, h, _ptr(sizes)))
        self.ctx.check(lib.asb_index_laplacian(self.ctx.handle, h, _ptr(indptr), _ptr(indices), _ptr(data)))
        aspace.lambdas = lam
        aspace.n_clusters = x
        aspace.cluster_assignments = asg
        aspace.cluster_sizes = sizes
        aspace.cluster_radius = radius
        aspace.is_standin = bool(getattr(self, "is_standin", False))   # cluster parameters from the stand-in heuristic
        aspace.lambda_stats = (info.lambda_min, info.lambda_max, info.lambda_sum / n)
        gl = GraphLaplacian(indptr, indices, data, n, gp, cent)
        if self.prebuilt_spectral:
            snnz = int(info.nnz_signals)
            sp, si, sd = np.empty(fg + 1, dtype=np.int64), np.empty(snnz, dtype=np.int64), np.empty(snnz, dtype=np.float64)
            self.ctx.check(lib.asb_index_signals(self.ctx.handle, h, _ptr(sp), _ptr(si), _ptr(sd)))
            aspace.signals = (sp, si, sd)
        return aspace, gl

    def __str__(self):  # src/builder.rs:459-524 Display: key=value, ...
        return (f"prebuilt_spectral={self.prebuilt_spectral}, lambda_eps={self.lambda_eps}, lambda_k={self.lambda_k}, "
                f"lambda_topk={self.lambda_topk}, lambda_p={self.lambda_p}, lambda_sigma={self.lambda_sigma}, "
                f"normalise={self.normalise}, sparsity_check={self.sparsity_check}, sampling={self.sampling}, "
                f"synthesis={self.synthesis}, cluster_max_clusters={self.cluster_max_clusters}, "
                f"cluster_radius={self.cluster_radius}, clustering_seed={self.clustering_seed}, "
                f"deterministic_clustering={self.deterministic_clustering}, "
                f"use_dims_reduction={self.use_dims_reduction}, rp_eps={self.rp_eps}")
