//! Raw bindings to the C ABI of `include/arrowspace_b200.h` -- one `extern "C"` item per exported symbol, same
//! names, same argument order.  Every data pointer may be a host or a device pointer (the library classifies it);
//! every function returning `c_int` returns `ASB_OK` or one of the `ASB_ERR_*` codes and leaves a message in
//! `asb_last_error(ctx)`.  The safe shim that implements arrowspace-rs' `EigenMaps` trait on top of these is shown in
//! INTEGRATION.md section 2.
//!
//! This crate is written against the header by hand (the authoring environment has no Rust toolchain and no
//! bindgen); `tests/test_abi.py::test_rust_sys_crate_covers_header` keeps the symbol list in sync with the header.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const ASB_OK: c_int = 0;
pub const ASB_ERR_INVALID: c_int = 1;
pub const ASB_ERR_CUDA: c_int = 2;
pub const ASB_ERR_NCCL: c_int = 3;
pub const ASB_ERR_NONFINITE_QUERY: c_int = 4; // src/core.rs:534-537
pub const ASB_ERR_ZERO_LAMBDA: c_int = 5; // src/core.rs:773-776
pub const ASB_ERR_SHAPE: c_int = 6; // src/laplacian.rs:129-134
pub const ASB_ERR_TOO_SPARSE: c_int = 7; // src/graph.rs:185-193
pub const ASB_ERR_NO_CLUSTERS: c_int = 8; // src/clustering.rs:869-874
pub const ASB_ERR_NAN_SCORE: c_int = 9; // src/core.rs:785
pub const ASB_ERR_ZERO_NORM: c_int = 10;
pub const ASB_ERR_EMPTY: c_int = 11; // src/core.rs:416-420
pub const ASB_ERR_DIM: c_int = 12; // src/core.rs:510-516
pub const ASB_ERR_CAPACITY: c_int = 13;
pub const ASB_ERR_UNSUPPORTED: c_int = 14;

pub const ASB_TAU_FIXED: i32 = 0; // TauMode, src/taumode.rs:75-82
pub const ASB_TAU_MEDIAN: i32 = 1;
pub const ASB_TAU_MEAN: i32 = 2;
pub const ASB_TAU_PERCENTILE: i32 = 3;

#[repr(C)]
pub struct asb_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct asb_comm {
    _private: [u8; 0],
}
#[repr(C)]
pub struct asb_index {
    _private: [u8; 0],
}

/// GraphParams, src/graph.rs:94-102, as passed by with_lambda_graph (src/builder.rs:109-137)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct asb_graph_params {
    pub eps: f64,
    pub k: i64,
    pub topk: i64,
    pub p: f64,
    pub has_sigma: i32,
    pub sigma: f64,
    pub normalise: i32,
    pub sparsity_check: i32,
    pub self_included: i32,
    pub rectified: i32,
}

/// Everything ArrowSpaceBuilder::build needs once (max_clusters, radius) are known (src/builder.rs:20-57)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct asb_build_params {
    pub graph: asb_graph_params,
    pub tau_mode: i32,
    pub tau_value: f64,
    pub max_clusters: i64,
    pub radius: f64,
    pub apply_define_result_k: i32,
    pub spectral: i32,
    pub projection: *const f64,
    pub reduced_dim: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct asb_index_info {
    pub n_items: i64,
    pub n_features: i64,
    pub n_clusters: i64,
    pub nnz: i64,
    pub lambda_min: f64,
    pub lambda_max: f64,
    pub lambda_sum: f64,
    pub radius: f64,
    pub max_clusters: i64,
    pub ms_cluster: f64,
    pub ms_laplacian: f64,
    pub ms_taumode: f64,
    pub ms_total: f64,
    pub nnz_signals: i64,
}

extern "C" {
    // ---- context ---------------------------------------------------------------------------------------------
    pub fn asb_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut asb_ctx) -> c_int;
    pub fn asb_ctx_destroy(ctx: *mut asb_ctx);
    pub fn asb_last_error(ctx: *mut asb_ctx) -> *const c_char;
    pub fn asb_status_string(status: c_int) -> *const c_char;
    pub fn asb_version() -> *const c_char;
    pub fn asb_kernel_launches(ctx: *mut asb_ctx) -> i64;
    pub fn asb_last_kernel_ms(ctx: *mut asb_ctx, which: *const c_char) -> f64;
    pub fn asb_ctx_set_option(ctx: *mut asb_ctx, key: *const c_char, value: f64) -> c_int;
    pub fn asb_search_slab_plan(sm_count: c_int, nq: i64, n: i64, max_slabs: i64, nslabs: *mut i64,
        tiles_per_slab: *mut i64) -> c_int;

    // ---- stage 1: clustering (src/clustering.rs) ---------------------------------------------------------------
    /// distance pass of estimate_intrinsic_dimension, src/clustering.rs:118-145
    pub fn asb_twonn_distances(ctx: *mut asb_ctx, rows: *const f64, n: i64, f: i64, sample_idx: *const i64, s: i64,
        d1: *mut f64, d2: *mut f64) -> c_int;
    /// run_incremental_clustering_with_sampling + nearest_centroid, src/clustering.rs:547-928
    pub fn asb_cluster_incremental(ctx: *mut asb_ctx, rows: *const f64, n: i64, f: i64, max_clusters: i64,
        radius: f64, centroids: *mut f64, assignments: *mut i64, sizes: *mut u64, x_out: *mut i64) -> c_int;
    pub fn asb_cluster_incremental_resume(ctx: *mut asb_ctx, rows: *const f64, n: i64, f: i64, max_clusters: i64,
        radius: f64, centroids: *mut f64, assignments: *mut i64, sizes: *mut u64, x_inout: *mut i64) -> c_int;

    // ---- stage 2: feature Laplacian (src/graph.rs:149-204, src/laplacian.rs:122-417) -----------------------------
    pub fn asb_laplacian_max_nnz(f: i64, topk: i64) -> i64;
    pub fn asb_build_feature_laplacian(ctx: *mut asb_ctx, centroids: *const f64, x: i64, f: i64,
        params: *const asb_graph_params, indptr: *mut i64, indices: *mut i64, data: *mut f64, capacity: i64,
        nnz_out: *mut i64) -> c_int;

    // ---- stage 3: taumode (src/taumode.rs:87-127,174-312,552-660; src/core.rs:533-549) ----------------------------
    pub fn asb_compute_taumode(ctx: *mut asb_ctx, items: *const f64, n: i64, f: i64, indptr: *const i64,
        indices: *const i64, data: *const f64, tau_mode: i32, tau_value: f64, lambdas: *mut f64, norms2: *mut f64,
        stats: *mut f64) -> c_int;
    pub fn asb_prepare_query_lambdas(ctx: *mut asb_ctx, queries: *const f64, nq: i64, f: i64, indptr: *const i64,
        indices: *const i64, data: *const f64, tau_mode: i32, tau_value: f64, lambda_q: *mut f64) -> c_int;

    // ---- search (src/core.rs:135-239,760-798,802-976; src/energymaps.rs:368-407,838-895) ---------------------------
    pub fn asb_search_lambda_aware_batch(ctx: *mut asb_ctx, items: *const f64, lambdas: *const f64,
        norms2: *const f64, n: i64, f: i64, queries: *const f64, lambda_q: *const f64, nq: i64, k: i64, alpha: f64,
        index_offset: i64, idx: *mut i64, score: *mut f64, count: *mut i64) -> c_int;
    pub fn asb_search_lambda_aware_hybrid_batch(ctx: *mut asb_ctx, items: *const f64, lambdas: *const f64,
        norms2: *const f64, n: i64, f: i64, queries: *const f64, lambda_q: *const f64, nq: i64, k: i64, alpha: f64,
        idx: *mut i64, score: *mut f64, count: *mut i64) -> c_int;
    pub fn asb_search_energy_batch(ctx: *mut asb_ctx, items: *const f64, lambdas: *const f64, norms2: *const f64,
        n: i64, f: i64, queries: *const f64, lambda_q: *const f64, nq: i64, k: i64, w_lambda: f64, w_dirichlet: f64,
        index_offset: i64, idx: *mut i64, score: *mut f64, count: *mut i64) -> c_int;
    pub fn asb_range_search(ctx: *mut asb_ctx, lambdas: *const f64, n: i64, lambda_q: f64, eps: f64,
        index_offset: i64, idx: *mut i64, dist: *mut f64, capacity: i64, count_out: *mut i64) -> c_int;
    pub fn asb_topk_merge(ctx: *mut asb_ctx, in_score: *const f64, in_idx: *const i64, parts: i64, nq: i64, k: i64,
        out_score: *mut f64, out_idx: *mut i64, out_count: *mut i64) -> c_int;

    // ---- JL projection with a materialised matrix (src/reduction.rs:127-199) --------------------------------------
    pub fn asb_project_matrix(ctx: *mut asb_ctx, rows: *const f64, n: i64, f: i64, projection: *const f64, r: i64,
        out: *mut f64) -> c_int;
    pub fn asb_jl_dimension(n_points: i64, epsilon: f64) -> i64;

    // ---- whole build with every intermediate resident in HBM (ArrowSpaceBuilder::build, src/builder.rs:249-455) ---
    pub fn asb_index_build(ctx: *mut asb_ctx, rows: *const f64, n: i64, f: i64, params: *const asb_build_params,
        out: *mut *mut asb_index) -> c_int;
    pub fn asb_index_destroy(index: *mut asb_index);
    pub fn asb_index_info_get(index: *const asb_index, info: *mut asb_index_info) -> c_int;
    pub fn asb_index_lambdas(ctx: *mut asb_ctx, index: *const asb_index, dst: *mut f64) -> c_int;
    pub fn asb_index_centroids(ctx: *mut asb_ctx, index: *const asb_index, dst: *mut f64) -> c_int;
    pub fn asb_index_assignments(ctx: *mut asb_ctx, index: *const asb_index, dst: *mut i64) -> c_int;
    pub fn asb_index_cluster_sizes(ctx: *mut asb_ctx, index: *const asb_index, dst: *mut u64) -> c_int;
    pub fn asb_index_laplacian(ctx: *mut asb_ctx, index: *const asb_index, indptr: *mut i64, indices: *mut i64,
        data: *mut f64) -> c_int;
    pub fn asb_index_signals(ctx: *mut asb_ctx, idx: *const asb_index, indptr: *mut i64, indices: *mut i64,
        data: *mut f64) -> c_int;
    pub fn asb_index_search(ctx: *mut asb_ctx, index: *const asb_index, queries: *const f64, nq: i64, k: i64,
        alpha: f64, idx: *mut i64, score: *mut f64, count: *mut i64, lambda_q_out: *mut f64) -> c_int;
    pub fn asb_index_search_lambda_aware(ctx: *mut asb_ctx, index: *const asb_index, queries: *const f64,
        lambda_q: *const f64, nq: i64, k: i64, alpha: f64, idx: *mut i64, score: *mut f64, count: *mut i64) -> c_int;

    pub fn asb_index_prepare_query(ctx: *mut asb_ctx, index: *const asb_index, queries: *const f64, nq: i64,
        lambda_q: *mut f64) -> c_int;
    pub fn asb_index_search_energy(ctx: *mut asb_ctx, index: *mut asb_index, queries: *const f64, nq: i64, k: i64,
        w_lambda: f64, w_dirichlet: f64, idx: *mut i64, score: *mut f64, count: *mut i64) -> c_int;

    // ---- row-sharded multi-GPU variants (one process per GPU, NCCL) ----
    pub fn asb_comm_unique_id(ctx: *mut asb_ctx, id_out_128_bytes: *mut c_void) -> c_int;
    pub fn asb_comm_init_rank(ctx: *mut asb_ctx, unique_id_128_bytes: *const c_void, nranks: c_int, rank: c_int,
        out: *mut *mut asb_comm) -> c_int;
    pub fn asb_comm_from_nccl(ctx: *mut asb_ctx, nccl_comm: *mut c_void, out: *mut *mut asb_comm) -> c_int;
    pub fn asb_comm_destroy(comm: *mut asb_comm);
    pub fn asb_comm_rank(comm: *const asb_comm) -> c_int;
    pub fn asb_comm_size(comm: *const asb_comm) -> c_int;
    pub fn asb_twonn_distances_sharded(ctx: *mut asb_ctx, comm: *mut asb_comm, rows_local: *const f64, n_local: i64,
        f: i64, shard_offset: i64, sample_idx: *const i64, s: i64, d1: *mut f64, d2: *mut f64) -> c_int;
    pub fn asb_index_build_sharded(ctx: *mut asb_ctx, comm: *mut asb_comm, rows_local: *const f64, n_local: i64, f: i64,
        shard_offset: i64, n_global: i64, params: *const asb_build_params, out: *mut *mut asb_index) -> c_int;
    pub fn asb_index_shard_offset(index: *const asb_index) -> i64;
    pub fn asb_index_search_sharded(ctx: *mut asb_ctx, comm: *mut asb_comm, index: *const asb_index,
        queries: *const f64, nq: i64, k: i64, alpha: f64, idx: *mut i64, score: *mut f64, count: *mut i64,
        lambda_q_out: *mut f64) -> c_int;
}

/// `Err(status)` unless `rc == ASB_OK` -- the shim turns it into the reference's panic.
#[inline]
pub fn check(rc: c_int) -> Result<(), c_int> {
    if rc == ASB_OK {
        Ok(())
    } else {
        Err(rc)
    }
}
