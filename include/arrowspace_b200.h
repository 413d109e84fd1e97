/*
 * arrowspace_b200.h -- C ABI of the B200-native (sm_100a) lambda-tau build + lambda-aware
 * search path of ArrowSpace.
 *
 * The reference (arrowspace-rs v0.18.1, pure Rust, CPU only) has no FFI; its seam is the
 * public `EigenMaps` trait (src/eigenmaps.rs:93-172) plus `ArrowSpace::prepare_query_item`
 * / `search_lambda_aware` (src/core.rs:533,760) and `ArrowSpaceBuilder::build`
 * (src/builder.rs:249).  Each entry point below replaces the body of one of those
 * methods; a Rust `-sys` crate binds exactly these symbols.
 *
 * Conventions
 *  - every matrix is contiguous ROW-MAJOR f64; CSR uses int64 indptr/indices (Rust usize).
 *  - every data pointer may be a HOST pointer or a DEVICE pointer on the context's GPU;
 *    the library classifies it (cudaPointerGetAttributes).  Host inputs are copied to the
 *    device inside the call, host outputs are copied back before the call returns; device
 *    pointers are used in place (zero copy).
 *  - calls are synchronous on return; one context per host thread; no global state.
 *  - stream contract: all work runs on the context's stream (the one given to asb_ctx_create, or a private non-blocking
 *    stream).  DEVICE buffers passed in must be complete when the call is made -- or produced on that same stream; the
 *    library inserts no cross-stream dependencies.  Outputs are complete when the call returns.
 *  - return value: 0 = ASB_OK, otherwise one of the ASB_ERR_* codes; asb_last_error(ctx)
 *    gives the text.  The library never aborts and has NO CPU fallback.
 */
#ifndef ARROWSPACE_B200_H
#define ARROWSPACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASB_OK 0
#define ASB_ERR_INVALID 1         /* bad argument                                         */
#define ASB_ERR_CUDA 2            /* CUDA runtime error / no device                        */
#define ASB_ERR_NCCL 3            /* NCCL error / libnccl.so.2 not loadable (row-sharded entry points only) */
#define ASB_ERR_NONFINITE_QUERY 4 /* src/core.rs:534-537 assert                            */
#define ASB_ERR_ZERO_LAMBDA 5     /* src/core.rs:773-776 assert_ne!(query.lambda, 0.0)     */
#define ASB_ERR_SHAPE 6           /* src/laplacian.rs:129-134 (needs >= 2x2)               */
#define ASB_ERR_TOO_SPARSE 7      /* src/graph.rs:185-193 (sparsity_check)                 */
#define ASB_ERR_NO_CLUSTERS 8     /* src/clustering.rs:869-874                             */
#define ASB_ERR_NAN_SCORE 9       /* src/core.rs:785 partial_cmp().unwrap() on NaN         */
#define ASB_ERR_ZERO_NORM 10      /* zero-magnitude feature column in the cosine kNN       */
#define ASB_ERR_EMPTY 11          /* src/core.rs:416-420 (items empty / single row)        */
#define ASB_ERR_DIM 12            /* src/core.rs:510-516 query dimension mismatch          */
#define ASB_ERR_CAPACITY 13       /* caller buffer too small                               */
#define ASB_ERR_UNSUPPORTED 14    /* valid in the reference, not built here (documented)   */

/* TauMode, src/taumode.rs:75-82 */
#define ASB_TAU_FIXED 0
#define ASB_TAU_MEDIAN 1
#define ASB_TAU_MEAN 2
#define ASB_TAU_PERCENTILE 3

typedef struct asb_ctx asb_ctx;
typedef struct asb_index asb_index;
typedef struct asb_comm asb_comm; /* one rank of a row-sharded job (wraps an ncclComm_t) */

/* GraphParams, src/graph.rs:94-102, as passed by with_lambda_graph (src/builder.rs:109-137) */
typedef struct {
    double eps;
    int64_t k;
    int64_t topk;
    double p;
    int32_t has_sigma; /* sigma: Option<f64>; None -> 1.0 (src/laplacian.rs:254) */
    double sigma;
    int32_t normalise;      /* src/laplacian.rs:146-151 StandardScaler over the F x X matrix: 0 off, 1 population variance
                             * (sum x^2 / F - mean^2), 2 sample variance; smartcore's own convention is unpinned */
    int32_t sparsity_check; /* src/graph.rs:185-193 */
    int32_t self_included;  /* smartcore kNN switch, default 0 (see DESIGN.md "unpinned") */
    int32_t rectified;      /* 0: 1-cos (default) ; 1: 1-max(0,cos) */
} asb_graph_params;

/* Everything ArrowSpaceBuilder::build needs once (max_clusters, radius) are known
 * (src/builder.rs:20-57; compute_optimal_k stays on the host, src/clustering.rs:36-72). */
typedef struct {
    asb_graph_params graph;
    int32_t tau_mode; /* with_synthesis, src/builder.rs:142-146 */
    double tau_value;
    int64_t max_clusters; /* builder.cluster_max_clusters, src/eigenmaps.rs:216 */
    double radius;        /* builder.cluster_radius (a SQUARED distance), :217 */
    int32_t apply_define_result_k; /* 1: k<=5 -> topk=3, k<10 -> topk=4 (src/builder.rs:225-233) */
    int32_t spectral; /* with_spectral (src/builder.rs:157-162): also build the Laplacian-of-Laplacian "signals"
                       * (src/graph.rs:211-231) and synthesise the ITEM lambdas from it (src/taumode.rs:195-200);
                       * query lambdas keep using the feature Laplacian (src/core.rs:548) */
    /* with_dims_reduction (src/builder.rs:181-185, src/eigenmaps.rs:248-269): the F x r Gaussian matrix of the
     * ImplicitProjection, MATERIALISED by the host (projection[j * r + k] = the sample for feature j, output k; the
     * reference redraws it from an 8-byte seed with ChaCha8 + StandardNormal, third-party generators), host or device
     * pointer; NULL / 0 = no projection.  The host decides r = min(compute_jl_dimension(n_clusters, eps), F / 2) and
     * whether to project at all (F > 64).  Effects, as in the reference: the centroids are projected before the feature
     * Laplacian (r x r graph); item lambdas read item[0 .. r) against that graph with tau and the denominator over all F
     * values; query lambdas and the energy search work on projected vectors; the lambda-aware search
     * of a projected index fails with ASB_ERR_DIM exactly where the reference panics ("items should be of the same
     * length", src/core.rs:157-161). */
    const double *projection;
    int64_t reduced_dim;
} asb_build_params;

typedef struct {
    int64_t n_items, n_features, n_clusters, nnz;
    double lambda_min, lambda_max, lambda_sum; /* src/eigenmaps.rs:372-382 */
    double radius;
    int64_t max_clusters;
    double ms_cluster, ms_laplacian, ms_taumode, ms_total; /* device stage times */
    int64_t nnz_signals; /* stored entries of the spectral signals matrix, 0 without with_spectral */
} asb_index_info;

/* ---- context ------------------------------------------------------------------------ */
/* stream: a cudaStream_t (or NULL: the library creates its own non-blocking stream). */
int asb_ctx_create(int device, void *stream, asb_ctx **out);
void asb_ctx_destroy(asb_ctx *ctx);
const char *asb_last_error(asb_ctx *ctx);
const char *asb_status_string(int status);
const char *asb_version(void);
/* number of kernels THIS library launched on the context since creation (bench evidence) */
int64_t asb_kernel_launches(asb_ctx *ctx);
/* device time of the most recent top-level call's dominant kernel(s), ms, measured with
 * CUDA events on the context's stream (0 if none). */
double asb_last_kernel_ms(asb_ctx *ctx, const char *which);
/* debug / test switches (results are identical whatever they are set to):
 *   "cluster_force_exact" (0|1)   take every clustering decision from the reference-arithmetic path instead of
 *                                 the certified fast path;
 *   "cluster_first_variant" (v)   start the clustering kernel selection at variant v: -2 pipelined tensor-core
 *                                 kernel (default), -1 FP32 blocked, 0 / 1 FP64 blocked (16 / 8 rows), 2 row-wise;
 *                                 a variant that does not fit shared memory falls through to the next one;
 *   "cluster_no_pipeline", "cluster_no_f32", "cluster_rowwise" (0|1)  shorthands for -1, 0 and 2;
 *   "cluster_phase_times" (0|1), "cluster_tick_tid" (t)  per-block cycle probes, read back with
 *                                 asb_last_kernel_ms(ctx, "cluster_phaseN");
 *   "taumode_generic" (0|1)       use the generic CSR kernel even for a symmetric graph;
 *   "taumode_regs" (1|0), "taumode_ipp" (1|2)   symmetric graphs, f <= 1024: keep the item in registers as well as in
 *                                 shared memory; ipp = 2 serves two items (f <= 512) per schedule entry;
 *                                 taumode_regs = 0: the shared-memory-only kernel of round 1;
 *   "search_prefilter" (1|0)      1 (default): asb_search_lambda_aware_batch / asb_index_search rank all pairs with a
 *                                 certified 3xTF32 tensor-core score and compute only the pairs the error bound cannot
 *                                 exclude from the top-k in the reference's FP64 arithmetic (k <= 32, n >= 1024; any
 *                                 input the bound does not cover falls back); 0: the exact FP64 DMMA kernel scores every
 *                                 pair.  Same ids either way; scores agree to the last few bits (both within 1e-12 of
 *                                 the reference, the prefilter path bit-identical to a sequential evaluation).
 *   "cluster_replay" (1|0)        0: the whole clustering walk runs on the sequential kernel.  1 (default): after a
 *                                 sequential prefix ("cluster_replay_prefix" rows,
 *                                 default 2048) chunks of "cluster_replay_chunk" rows (default 1024, doubling after every proven
 *                                 chunk up to "cluster_replay_chunk_max", default 262144) are replayed in
 *                                 parallel -- nearest / runner-up centroid of the chunk-start snapshot for all rows at
 *                                 once, one sequential chain per centroid, every row's decision proven from the
 *                                 centroids' net displacement -- and a chunk with a single unproven row is walked by
 *                                 the sequential kernel instead ("cluster_replay_generic_chain" = 1 selects the chain
 *                                 kernel that keeps the centroid in memory instead of registers).  Same bits either
 *                                 way (csrc/cluster_replay.cu,
 *                                 tests/replay_proto.py).
 *   "twonn_prefilter" (1|0)       1 (default): the Two-NN scan is ranked by the certified split-operand score of the search
 *                                 prefilter (score -|q - x|^2, tcgen05 tile) and the surviving distances are evaluated in
 *                                 the reference's direct form (bit-identical to a sequential evaluation); 0: the FP64
 *                                 tensor kernel (distances to 1e-9).
 *   "cluster_replay_near" (1|0)   1 (default): the replay ranks a chunk's rows against the snapshot on the tcgen05 tile
 *                                 and works with certified distance BOUNDS (a chunk whose bounds are too wide is ranked
 *                                 again by the FP64 kernel); 0: always the FP64 kernel's exact top-2.
 *   "cluster_replay_tf32" (0|1)   the exact top-2 through the certified prefilter + direct-form distances (slower than
 *                                 either of the above; kept for cross-checks).
 *   "search_umma_bf16" (1|0)      operand planes of the tcgen05 tile: BF16x3 on kind::f16 (default) or 3xTF32 on kind::tf32;
 *   "search_umma" (1|0), "search_umma_kc" (16|32), "search_umma_cluster" (2|1|4), "search_umma_slab_mb" (80 | 48)
 *                                 the prefilter tile: tcgen05.mma + TMA + TMEM (1) or mma.sync + cp.async (0); 32-bit words
 *                                 per stage row (32: TF32 planes only); CTAs sharing one multicast item stream; MB of item
 *                                 planes per slab (default 80 for BF16 planes, 48 for TF32 planes).
 *   "cluster_replay_growth" (2)   chunk size factor after a proven chunk; "cluster_chain_probe" (0|1) records per chain
 *                                 block {rows, start, end} of the largest chunk ("cluster_probe_*" diagnostics).
 *   "build_overlap_upload" (1|0)  asb_index_build from HOST rows: upload in 16 MB chunks on a second stream while stage 1
 *                                 already works on the head of the matrix.
 *   "cluster_growth_run" (1|0), "cluster_shard_snapshot_rows", "cluster_shard_piece", "cluster_shard_speculate"
 *                                 the creator run at the start of a walk; rows rank 0 walks before it broadcasts the
 *                                 common snapshot (262144); rows per certified piece of a later shard (262144); 0 turns
 *                                 the speculative ranking off (plain hand-off).
 * Read-only diagnostics through asb_last_kernel_ms: "cluster_replay_chunks", "cluster_replay_chunks_ok",
 * "cluster_replay_rows", "cluster_replay_{seq,top2,chain}_ms", "cluster_wall_{growth,prefix,prepare,run,fallback}_ms",
 * "cluster_chain_{rows_grouped,rows_by_row,exact_steps,checkpoints}", "search_pf_used", "search_pf_umma",
 * "search_umma_bf16", "search_pf_flags", "search_pf_overflow_queries" (queries sent alone to the exact kernel),
 * "search_pf_candidates", "search_pf_rescored", "search_pf_cap", "search_pf_slabs", "search_pf_band",
 * "shard_t_{snapshot,ranked,state_in,walked,done}_ms", "cluster_shard_{speculative,fallback}". */
int asb_ctx_set_option(asb_ctx *ctx, const char *key, double value);
/* The slab split the search kernels use for nq queries x n items on a device with sm_count SMs (pure host
 * arithmetic, no device needed): a (128-query tile, slab) pair is one CTA and CTAs run in waves of sm_count, so the
 * split minimises ceil(units / sm_count) * tiles_per_slab.  max_slabs: 4096 for the exact kernel, 64 for the
 * prefilter.  nslabs * tiles_per_slab covers ceil(n / 128) item tiles. */
int asb_search_slab_plan(int sm_count, int64_t nq, int64_t n, int64_t max_slabs, int64_t *nslabs,
                         int64_t *tiles_per_slab);

/* ---- stage 1: clustering ------------------------------------------------------------ */
/* Two-NN scan: replaces the distance pass of estimate_intrinsic_dimension
 * (src/clustering.rs:118-145).  For each sample row the two smallest Euclidean distances
 * to every OTHER row.  sample_idx (host or device, int64[s]) is produced by the host
 * (StdRng shuffle, :113-116).  d1/d2: f64[s]. */
int asb_twonn_distances(asb_ctx *ctx, const double *rows, int64_t n, int64_t f,
                        const int64_t *sample_idx, int64_t s, double *d1, double *d2);

/* Incremental leader clustering in row order: replaces
 * run_incremental_clustering_with_sampling + nearest_centroid (src/clustering.rs:547-928)
 * for the deterministic configuration (with_seed, with_inline_sampling(None)).
 * centroids: f64[max_clusters*f] (first x rows valid), assignments: int64[n] (-1 = None),
 * sizes: uint64[max_clusters]. */
int asb_cluster_incremental(asb_ctx *ctx, const double *rows, int64_t n, int64_t f,
                            int64_t max_clusters, double radius, double *centroids,
                            int64_t *assignments, uint64_t *sizes, int64_t *x_out);

/* Same walk, continuing from an existing state: centroids/sizes hold *x_inout centroids on
 * entry and the updated state on exit.  Processing shard 0, then shard 1 resumed from shard 0's
 * state, ... equals one call over the concatenated rows -- the order-preserving hand-off used
 * when the items are row-sharded over several GPUs (the centroid state travels, not the rows). */
int asb_cluster_incremental_resume(asb_ctx *ctx, const double *rows, int64_t n, int64_t f,
                                   int64_t max_clusters, double radius, double *centroids,
                                   int64_t *assignments, uint64_t *sizes, int64_t *x_inout);

/* ---- stage 2: feature-graph Laplacian ----------------------------------------------- */
/* upper bound of stored entries: f * (1 + 2*(topk+1)) */
int64_t asb_laplacian_max_nnz(int64_t f, int64_t topk);
/* Replaces GraphFactory::build_laplacian_matrix_from_k_cluster (src/graph.rs:149-204) ->
 * build_laplacian_matrix (src/laplacian.rs:122-417).  centroids: X x F; output CSR F x F:
 * indptr int64[f+1], indices int64[capacity], data f64[capacity]. */
int asb_build_feature_laplacian(asb_ctx *ctx, const double *centroids, int64_t x, int64_t f,
                                const asb_graph_params *params, int64_t *indptr,
                                int64_t *indices, double *data, int64_t capacity,
                                int64_t *nnz_out);

/* ---- stage 3: taumode --------------------------------------------------------------- */
/* Replaces TauMode::compute_taumode_lambdas_parallel (src/taumode.rs:174-312): per item
 * tau = select_tau(item values) (:87-127), lambda = synthetic lambda (:552-660).
 * stats (optional, host f64[3]) = {min, max, sum} of lambdas (src/eigenmaps.rs:372-382).
 * norms2 (optional, f64[n]) receives sum(x^2) per item (reused by search). */
int asb_compute_taumode(asb_ctx *ctx, const double *items, int64_t n, int64_t f,
                        const int64_t *indptr, const int64_t *indices, const double *data,
                        int32_t tau_mode, double tau_value, double *lambdas, double *norms2,
                        double *stats);

/* Replaces ArrowSpace::prepare_query_item (src/core.rs:533-549) for a batch of queries.
 * Returns ASB_ERR_NONFINITE_QUERY if any query value is NaN/Inf. */
int asb_prepare_query_lambdas(asb_ctx *ctx, const double *queries, int64_t nq, int64_t f,
                              const int64_t *indptr, const int64_t *indices,
                              const double *data, int32_t tau_mode, double tau_value,
                              double *lambda_q);

/* ---- stage 5: search ---------------------------------------------------------------- */
/* Replaces ArrowSpace::search_lambda_aware (src/core.rs:760-798) for a batch: score =
 * alpha*cos + (1-alpha)*(1 - min(|lq - li|, 1)); stable descending order (ties -> lower
 * index); count[q] = min(k, n).  idx: int64[nq*k], score: f64[nq*k], count: int64[nq].
 * index_offset is added to every returned index (row-sharded multi-GPU search).
 * norms2 may be NULL (computed internally).  k <= 128. */
int asb_search_lambda_aware_batch(asb_ctx *ctx, const double *items, const double *lambdas,
                                  const double *norms2, int64_t n, int64_t f,
                                  const double *queries, const double *lambda_q, int64_t nq,
                                  int64_t k, double alpha, int64_t index_offset,
                                  int64_t *idx, double *score, int64_t *count);

/* ---- "next" rows ---------------------------------------------------- */
/* JL random projection with a MATERIALISED matrix.  ImplicitProjection::project /
 * project_matrix / project_query (src/reduction.rs:143-199, src/core.rs:509-529) regenerate the F x r Gaussian matrix
 * from a seed on every call (ChaCha8 + StandardNormal, third-party generators); the host draws it once, in the
 * reference's order (projection[j * r + k] = the sample for feature j, output k), and hands it over.
 * out[i * r + k] = sum_j (rows[i,j] * projection[j,k]) * (1 / sqrt(r)), accumulated in the reference's order with
 * separately rounded operations (bit-identical).  rows: n x f, out: n x r (host or device). */
int asb_project_matrix(asb_ctx *ctx, const double *rows, int64_t n, int64_t f, const double *projection, int64_t r,
                       double *out);
/* compute_jl_dimension (src/reduction.rs:127-141): max(ceil(8 ln(n_points) / eps^2), 32) */
int64_t asb_jl_dimension(int64_t n_points, double epsilon);

/* EnergyMaps::search_energy (src/energymaps.rs:368-407) with ProjectedEnergy::score
 * (:838-895) for an index without projection and without spectral signals: per item
 *   energy = w_lambda * |lambda_q - lambda_i| + w_dirichlet * min(d / (1 + d), 1),  d = |q - x_i|_2,
 * results are (index, -energy), best (largest) first, ties -> lower index, truncated to k (<= 60).
 * lambda_q comes from asb_prepare_query_lambdas (the reference recomputes it per item, :884).  A fused pass
 * ranks k+4 candidates by the GEMM-form distance, a second pass rescores them in the reference's direct form. */
int asb_search_energy_batch(asb_ctx *ctx, const double *items, const double *lambdas, const double *norms2, int64_t n,
                            int64_t f, const double *queries, const double *lambda_q, int64_t nq, int64_t k,
                            double w_lambda, double w_dirichlet, int64_t index_offset, int64_t *idx, double *score,
                            int64_t *count);
/* Replaces ArrowSpace::search_lambda_aware_hybrid (src/core.rs:802-928) for a batch: the union of
 * {cosine > 0.9999} (scored by cosine), the lambda-aware top-k (scored alpha*cos+(1-alpha)*lam
 * unless already present) and the semantic top-1, sorted by that score, truncated to k.  Built
 * from two fused searches (blended, and alpha = 1).  Score ties (unspecified in the reference:
 * rayon fold + sort_unstable) resolve to the lower index.  count[q] <= k. */
int asb_search_lambda_aware_hybrid_batch(asb_ctx *ctx, const double *items, const double *lambdas,
                                         const double *norms2, int64_t n, int64_t f,
                                         const double *queries, const double *lambda_q, int64_t nq,
                                         int64_t k, double alpha, int64_t *idx, double *score,
                                         int64_t *count);
/* Replaces the scan of ArrowSpace::range_search (src/core.rs:959-969): every item with
 * lambda_q - lambda_i <= eps (signed difference, as written), in index order.  idx/dist hold
 * `capacity` entries; *count_out receives the number of hits (ASB_ERR_CAPACITY if larger). */
int asb_range_search(asb_ctx *ctx, const double *lambdas, int64_t n, double lambda_q, double eps,
                     int64_t index_offset, int64_t *idx, double *dist, int64_t capacity,
                     int64_t *count_out);

/* k-way merge of per-shard top-k lists (multi-GPU search): in_score/in_idx are
 * [parts][nq][k] (unused tail slots: idx = -1); out [nq][k] ordered by (score desc, idx
 * asc) -- the order a single-process stable sort gives (src/core.rs:785). */
int asb_topk_merge(asb_ctx *ctx, const double *in_score, const int64_t *in_idx, int64_t parts,
                   int64_t nq, int64_t k, double *out_score, int64_t *out_idx,
                   int64_t *out_count);

/* ---- whole build: ArrowSpaceBuilder::build (src/builder.rs:249-455) ----------------- */
/* stages 1-3 with every intermediate resident in HBM.  rows may be host (copied once) or
 * device (borrowed: must outlive the index). */
int asb_index_build(asb_ctx *ctx, const double *rows, int64_t n, int64_t f,
                    const asb_build_params *params, asb_index **out);
void asb_index_destroy(asb_index *index);
int asb_index_info_get(const asb_index *index, asb_index_info *info);
/* copy-out accessors (dst host or device). */
int asb_index_lambdas(asb_ctx *ctx, const asb_index *index, double *dst);       /* f64[n]   */
int asb_index_centroids(asb_ctx *ctx, const asb_index *index, double *dst);     /* f64[x*f] */
int asb_index_assignments(asb_ctx *ctx, const asb_index *index, int64_t *dst);  /* i64[n]   */
int asb_index_cluster_sizes(asb_ctx *ctx, const asb_index *index, uint64_t *dst); /* u64[x] */
int asb_index_laplacian(asb_ctx *ctx, const asb_index *index, int64_t *indptr,
                        int64_t *indices, double *data); /* f+1, nnz, nnz */
/* EigenMaps::search (src/eigenmaps.rs:410-455) for a batch: prepare_query_item +
 * search_lambda_aware.  lambda_q_out optional (f64[nq]). */
/* aspace.signals (src/core.rs:370; F x F CSR, capacity asb_laplacian_max_nnz(F, topk)); ASB_ERR_INVALID when the
 * index was built without `spectral`. */
int asb_index_signals(asb_ctx *ctx, const asb_index *idx, int64_t *indptr, int64_t *indices, double *data);
int asb_index_search(asb_ctx *ctx, const asb_index *index, const double *queries, int64_t nq,
                     int64_t k, double alpha, int64_t *idx, double *score, int64_t *count,
                     double *lambda_q_out);

/* ArrowSpace::prepare_query_item (src/core.rs:533-549) against the resident index for a batch: the queries are projected
 * first when the index carries a projection (:540-545), tau comes from the (projected) query, lambda from gl.matrix. */
int asb_index_prepare_query(asb_ctx *ctx, const asb_index *index, const double *queries, int64_t nq, double *lambda_q);

/* EnergyMaps::search_energy with ProjectedEnergy::score (src/energymaps.rs:368-407,838-895) against the resident index,
 * every branch: project_vec through the index's projection (if any), projected_dirichlet through the spectral signals
 * when they exist and match the (projected) dimension -- y = S (q' - x'), |y| / (1 + |y|) -- else bounded L2.  A fused
 * pass ranks k + 4 candidates per query in the transformed space, a second pass rescoring them from the difference
 * vector.  Results (index, -energy), best first, ties -> lower index, k <= 56. */
int asb_index_search_energy(asb_ctx *ctx, asb_index *index, const double *queries, int64_t nq, int64_t k, double w_lambda,
                            double w_dirichlet, int64_t *idx, double *score, int64_t *count);

/* ArrowSpace::search_lambda_aware (src/core.rs:760-798) against the resident index with
 * caller-prepared query lambdas (the reference's two-step prepare_query_item + search). */
int asb_index_search_lambda_aware(asb_ctx *ctx, const asb_index *index, const double *queries,
                                  const double *lambda_q, int64_t nq, int64_t k, double alpha,
                                  int64_t *idx, double *score, int64_t *count);

/* ---- row-sharded multi-GPU variants: one process per GPU, NCCL over NVLink ---------------------------
 * Items (and their lambdas) are sharded by row in rank order: rank g holds global rows [offset_g, offset_g + n_g);
 * centroids, the feature Laplacian and the queries are replicated.  No collective touches the N x F items.  The library
 * opens libnccl.so.2 at run time, so a process that already loaded NCCL shares that copy.
 *   - either wrap the host's communicator:        asb_comm_from_nccl(ctx, ncclComm_t, &comm)
 *   - or let the library create one: rank 0 calls asb_comm_unique_id (128 bytes = ncclUniqueId), ships the bytes to the
 *     other ranks by any means, every rank calls asb_comm_init_rank.
 * Every sharded call is collective: all ranks must make it, with the same replicated arguments. */
int asb_comm_unique_id(asb_ctx *ctx, void *id_out_128_bytes);
int asb_comm_init_rank(asb_ctx *ctx, const void *unique_id_128_bytes, int nranks, int rank, asb_comm **out);
int asb_comm_from_nccl(asb_ctx *ctx, void *nccl_comm, asb_comm **out); /* borrowed: not destroyed with the handle */
void asb_comm_destroy(asb_comm *comm);
int asb_comm_rank(const asb_comm *comm);
int asb_comm_size(const asb_comm *comm);

/* asb_twonn_distances over the shards (src/clustering.rs:118-145): sample_idx holds GLOBAL row indices, identical on every
 * rank.  Sample rows are assembled with one all-reduce (s x f doubles), each rank scans its own shard, the per-shard two
 * nearest distances are all-gathered (2 s doubles per rank) and merged.  d1 / d2: f64[s] on every rank. */
int asb_twonn_distances_sharded(asb_ctx *ctx, asb_comm *comm, const double *rows_local, int64_t n_local, int64_t f,
                                int64_t shard_offset, const int64_t *sample_idx, int64_t s, double *d1, double *d2);

/* asb_index_build over the shards.  Stage 1 is the order-preserving walk: rank 0 walks the head of its shard and
 * broadcasts that state as the common snapshot; the other ranks rank their rows against it while the ranks before them
 * are still walking; the K x F state (<= 6.1 MB) then travels rank to rank (ncclSend / ncclRecv) and each shard's chains are
 * certified against it exactly as on one GPU (a shard that cannot be certified is re-ranked from the fresh state: same
 * bits).  Stage 2 runs on rank 0, the CSR is broadcast.  Stage 3 is local; lambda min / max / sum are all-reduced
 * (src/eigenmaps.rs:372-382), so asb_index_info reports GLOBAL statistics; n_items stays the local row count.  The
 * accessors (lambdas, assignments) return the local shard; centroids / sizes / Laplacian are the global ones. */
int asb_index_build_sharded(asb_ctx *ctx, asb_comm *comm, const double *rows_local, int64_t n_local, int64_t f,
                            int64_t shard_offset, int64_t n_global, const asb_build_params *params, asb_index **out);
int64_t asb_index_shard_offset(const asb_index *index);

/* asb_index_search over the shards: local top-k with GLOBAL indices, ncclAllGather of the lists (Q x k x 16 B per rank),
 * k-way merge by (score desc, index asc) = the reference's stable sort (src/core.rs:785).  Merged result on every rank;
 * an error on any rank (non-finite query, lambda_q == 0, NaN score) is returned by all of them. */
int asb_index_search_sharded(asb_ctx *ctx, asb_comm *comm, const asb_index *index, const double *queries, int64_t nq,
                             int64_t k, double alpha, int64_t *idx, double *score, int64_t *count, double *lambda_q_out);

#ifdef __cplusplus
}
#endif
#endif /* ARROWSPACE_B200_H */
