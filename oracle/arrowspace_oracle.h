/*
 * arrowspace_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, f64, left-to-right sums, no FMA contraction) of the
 * arrowspace-rs v0.18.1 hot path: incremental leader clustering, feature-graph
 * Laplacian, taumode lambda synthesis and lambda-aware search.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libarrowspace_b200.so) never links or calls it.
 *
 * PARITY PINNING: the Rust reference cannot be compiled in this environment (no
 * cargo/rustc, crates un-vendored), so this restatement is pinned against the
 * reference's own known-answer tests and fixtures (see tests/test_oracle_kat.py):
 *   - select_tau table            src/tests/test_taumode.rs:14-159
 *   - nearest_centroid            src/tests/test_clustering.rs:22-59
 *   - synthetic lambda closed form / scale invariance  src/tests/test_taumode.rs:499-528
 *   - alpha=1 top-3 {3,6,0} on the 64x24 table          paper.md:123-133
 *   - Laplacian structural invariants                    src/tests/test_laplacian.rs:51-152
 * The kNN inside the Laplacian lives in the un-vendored crate smartcore 0.4.5
 * (CosinePair::query_row_top_k): neighbour selection at exact distance ties and
 * the self-inclusion question are "parity unpinned"; they are explicit switches
 * here (aso_lap_params.self_included / rectified).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the arrowspace-rs repository root).
 */
#ifndef ARROWSPACE_ORACLE_H
#define ARROWSPACE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes mirror include/arrowspace_b200.h */
#define ASO_OK 0
#define ASO_ERR_INVALID 1
#define ASO_ERR_NONFINITE_QUERY 4
#define ASO_ERR_ZERO_LAMBDA 5
#define ASO_ERR_SHAPE 6
#define ASO_ERR_TOO_SPARSE 7
#define ASO_ERR_NO_CLUSTERS 8
#define ASO_ERR_NAN_SCORE 9
#define ASO_ERR_ZERO_NORM 10

#define ASO_TAU_FIXED 0
#define ASO_TAU_MEDIAN 1
#define ASO_TAU_MEAN 2
#define ASO_TAU_PERCENTILE 3

#define ASO_TAU_FLOOR 1e-10 /* src/taumode.rs:84 */

/* src/taumode.rs:87-127 */
double aso_select_tau(const double *x, size_t n, int mode, double value);

/* src/taumode.rs:552-660 (== :381-519) */
double aso_synthetic_lambda_g(const double *x, int64_t f_item, int64_t f_graph, const int64_t *indptr,
                              const int64_t *indices, const double *data, double tau);
double aso_synthetic_lambda(const double *x, int64_t f, const int64_t *indptr,
                            const int64_t *indices, const double *data, double tau);

/* src/taumode.rs:174-312 (per-item tau from the item's own values, :233-234).
 * threads<=0 -> omp default. */
int aso_compute_taumode(const double *items, int64_t n, int64_t f, const int64_t *indptr,
                        const int64_t *indices, const double *data, int mode, double value,
                        double *lambdas, int threads);

/* src/core.rs:533-549 ; returns ASO_ERR_NONFINITE_QUERY like the reference's assert */
int aso_prepare_query_item(const double *q, int64_t f, const int64_t *indptr,
                           const int64_t *indices, const double *data, int mode, double value,
                           double *lambda_out);

/* src/clustering.rs:913-928 */
int64_t aso_nearest_centroid(const double *row, const double *centroids, int64_t k, int64_t f,
                             double *d2_out);

/* src/clustering.rs:547-910, deterministic branch (:842-843), sampling None.
 * centroids: max_clusters*f doubles (row-major), assignments: n (-1 = None),
 * sizes: max_clusters. */
int aso_cluster_incremental(const double *rows, int64_t n, int64_t f, int64_t max_clusters,
                            double radius, double *centroids, int64_t *assignments,
                            uint64_t *sizes, int64_t *x_out);

/* src/clustering.rs:118-145 : two smallest Euclidean distances from each sampled
 * row to every other row.  d1/d2: s doubles each. */
int aso_twonn_distances(const double *rows, int64_t n, int64_t f, const int64_t *sample_idx,
                        int64_t s, double *d1, double *d2, int threads);
/* src/clustering.rs:108-110,136-163 : n<10 rule, ratios -> mean ->
 * clamp(round(1/ln(mean)),1,F) */
int64_t aso_intrinsic_dim(int64_t n, int64_t f, const double *d1, const double *d2, int64_t s);
int64_t aso_intrinsic_dim_from_distances(const double *d1, const double *d2, int64_t s,
                                         int64_t f);
/* src/clustering.rs:75-98 */
void aso_step1_bounds(int64_t n, int64_t f, int64_t id_est, int64_t *k_min, int64_t *k_max);

typedef struct {
    double eps;
    int64_t k;
    int64_t topk;
    double p;
    int has_sigma;
    double sigma;
    int normalise;      /* 0 off; 1: StandardScaler, population variance; 2: sample variance (third-party: unpinned) */
    int sparsity_check; /* src/graph.rs:185-193 */
    int self_included;  /* 0 (default): kNN result never contains the query row */
    int rectified;      /* 0 (default): dist = 1 - cos ; 1: 1 - max(0,cos) */
} aso_lap_params;

/* src/graph.rs:149-204 + src/laplacian.rs:122-417.
 * centroids: X x F row-major.  Output CSR is F x F.  indices/data capacity must be
 * >= F * (1 + 2*(topk+1)).  nnz_out receives the number of stored entries. */
int aso_feature_laplacian(const double *centroids, int64_t x, int64_t f,
                          const aso_lap_params *params, int64_t *indptr, int64_t *indices,
                          double *data, int64_t *nnz_out);

/* GraphFactory::build_spectral_laplacian, src/graph.rs:211-231 -- signals =
 * build_laplacian_matrix(dense(L)^T, params): the F rows of L are the "items", its F columns the nodes.
 * L in CSR (F x F); output CSR with the same capacity rule as aso_feature_laplacian. */
int aso_spectral_signals(const int64_t *l_indptr, const int64_t *l_indices, const double *l_data, int64_t f,
                         const aso_lap_params *params, int64_t *indptr, int64_t *indices, double *data,
                         int64_t *nnz_out);

/* src/core.rs:135-239,760-798.  Returns count = min(k,n) in *count_out. */
int aso_search_lambda_aware(const double *items, const double *lambdas, int64_t n, int64_t f,
                            const double *q, double lambda_q, int64_t k, double alpha,
                            int64_t *idx_out, double *score_out, int64_t *count_out);

/* the reference's "batched" search is a loop over queries
 * (benches/index_compute_bench.rs:250-262); here parallel over queries. */
int aso_search_lambda_aware_batch(const double *items, const double *lambdas, int64_t n,
                                  int64_t f, const double *queries, const double *lambda_q,
                                  int64_t nq, int64_t k, double alpha, int64_t *idx_out,
                                  double *score_out, int64_t *count_out, int threads);

/* src/core.rs:802-928 and :944-976 */
int aso_search_lambda_aware_hybrid(const double *items, const double *lambdas, int64_t n, int64_t f,
                                   const double *q, double lambda_q, int64_t k, double alpha,
                                   int64_t *idx_out, double *score_out, int64_t *count_out);
/* src/energymaps.rs:368-407 + :838-895 (no projection, no signals): (index, -energy) best first */
int aso_search_energy(const double *items, const double *lambdas, int64_t n, int64_t f, const double *q,
                      double lambda_q, int64_t k, double w_lambda, double w_dirichlet, int64_t *idx_out,
                      double *score_out, int64_t *count_out);
/* every branch of ProjectedEnergy::score (src/energymaps.rs:856-895): projection f x r (or NULL), signals CSR sig_n x
 * sig_n (or NULL) */
int aso_search_energy_ex(const double *items, const double *lambdas, int64_t n, int64_t f, const double *q,
                         double lambda_q, int64_t k, double w_lambda, double w_dirichlet, const double *projection,
                         int64_t r, const int64_t *sig_indptr, const int64_t *sig_indices, const double *sig_data,
                         int64_t sig_n, int64_t *idx_out, double *score_out, int64_t *count_out);
/* src/reduction.rs:127-141 and :143-199 with the Gaussian matrix materialised by the caller
 * (projection[j * r + k] = the sample drawn for feature j, output k) */
int64_t aso_jl_dimension(int64_t n_points, double epsilon);
int aso_project_matrix(const double *rows, int64_t n, int64_t f, const double *projection, int64_t r, double *out);
int aso_range_search(const double *lambdas, int64_t n, double lambda_q, double eps, int64_t *idx_out,
                     double *dist_out, int64_t *count_out);

int aso_num_threads(void);
int aso_set_num_threads(int threads);

#ifdef __cplusplus
}
#endif
#endif
