/*
 * arrowspace_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see the header).
 *
 * Plain-C restatement of the arrowspace-rs v0.18.1 CPU hot path.  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC
 * (-ffp-contract=off because Rust never fuses a*b+c; -O2 without -ffast-math keeps
 * every sum in source order).
 */
#include "arrowspace_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* src/graph.rs:211-231: aspace.signals = build_laplacian_matrix(sparse_to_dense(&gl.matrix).transpose(), params,
 * Some(nitems)).matrix.  aso_feature_laplacian(M, X, F) is build_laplacian_matrix(M^T), so M = dense(L). */
int aso_spectral_signals(const int64_t *l_indptr, const int64_t *l_indices, const double *l_data, int64_t f,
                         const aso_lap_params *params, int64_t *indptr, int64_t *indices, double *data,
                         int64_t *nnz_out) {
    if (!l_indptr || !l_indices || !l_data || f <= 0) return ASO_ERR_INVALID;
    double *dense = (double *)calloc((size_t)f * (size_t)f, sizeof(double));
    if (!dense) return ASO_ERR_INVALID;
    for (int64_t r = 0; r < f; ++r)
        for (int64_t e = l_indptr[r]; e < l_indptr[r + 1]; ++e) dense[(size_t)r * f + l_indices[e]] = l_data[e];
    const int rc = aso_feature_laplacian(dense, f, f, params, indptr, indices, data, nnz_out);
    free(dense);
    return rc;
}

int aso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* the OpenMP team size of every later call with threads <= 0 (a launcher may have exported OMP_NUM_THREADS=1:
 * torch.distributed.run does); returns the size now in effect */
int aso_set_num_threads(int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    return omp_get_max_threads();
#else
    (void)threads;
    return 1;
#endif
}

/* ------------------------------------------------------------------ taumode */

static int cmp_f64_asc(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* src/taumode.rs:87-127 */
double aso_select_tau(const double *x, size_t n, int mode, double value) {
    if (mode == ASO_TAU_FIXED) { /* :89-95 */
        return (isfinite(value) && value > 0.0) ? value : ASO_TAU_FLOOR;
    }
    if (mode == ASO_TAU_MEAN) { /* :96-107 */
        double sum = 0.0;
        size_t cnt = 0;
        for (size_t i = 0; i < n; ++i)
            if (isfinite(x[i])) {
                sum += x[i];
                cnt++;
            }
        double m = cnt > 0 ? sum / (double)cnt : 0.0;
        return fmax(m, ASO_TAU_FLOOR);
    }
    /* Median | Percentile :108-125 */
    double *v = (double *)malloc((n ? n : 1) * sizeof(double));
    size_t len = 0;
    for (size_t i = 0; i < n; ++i)
        if (isfinite(x[i])) v[len++] = x[i];
    if (len == 0) {
        free(v);
        return ASO_TAU_FLOOR;
    }
    qsort(v, len, sizeof(double), cmp_f64_asc);
    double r;
    if (mode == ASO_TAU_PERCENTILE) {
        double pp = value; /* p.clamp(0,1); NaN stays NaN and `as usize` gives 0 */
        if (pp < 0.0) pp = 0.0;
        if (pp > 1.0) pp = 1.0;
        double fi = round((double)(len - 1) * pp); /* half away from zero, like f64::round */
        size_t idx = (fi != fi || fi < 0.0) ? 0 : (size_t)fi;
        if (idx >= len) idx = len - 1;
        r = fmax(v[idx], ASO_TAU_FLOOR);
    } else if (len % 2 == 1) {
        r = fmax(v[len / 2], ASO_TAU_FLOOR);
    } else {
        r = fmax(0.5 * (v[len / 2 - 1] + v[len / 2]), ASO_TAU_FLOOR);
    }
    free(v);
    return r;
}

double aso_synthetic_lambda_g(const double *x, int64_t f_item, int64_t f, const int64_t *indptr,
                              const int64_t *indices, const double *data, double tau);

/* src/taumode.rs:552-660.  Rows are visited in order and their partial sums are
 * added in row order (the reference reduces them in a nondeterministic rayon
 * order, :565-588 -- any order is "the reference"). */
double aso_synthetic_lambda(const double *x, int64_t f, const int64_t *indptr,
                            const int64_t *indices, const double *data, double tau) {
    return aso_synthetic_lambda_g(x, f, f, indptr, indices, data, tau);
}

/* The same with a graph of rg <= f nodes: graph.outer_iterator() runs over the graph's rows (:565,:611) and indexes
 * item_vector[i], item_vector[j] with i, j < rg, while the denominator (:596) runs over the WHOLE item -- the shape a
 * JL-projected build produces (r x r graph, F-long items: src/eigenmaps.rs:248-269). */
double aso_synthetic_lambda_g(const double *x, int64_t f_item, int64_t f, const int64_t *indptr,
                              const int64_t *indices, const double *data, double tau) {
    double numerator = 0.0, edge_energy_sum = 0.0;
    for (int64_t i = 0; i < f; ++i) {
        double xi = x[i];
        double local_num = 0.0, local_edge = 0.0;
        for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
            int64_t j = indices[e];
            double lij = data[e];
            local_num += xi * lij * x[j]; /* (xi*lij)*x[j]  :575 */
            if (i != j) {
                double w = fmax(-lij, 0.0);
                if (w > 0.0) {
                    double d = xi - x[j];
                    local_edge += w * d * d; /* (w*d)*d  :581 */
                }
            }
        }
        numerator += local_num;
        edge_energy_sum += local_edge;
    }
    double denominator = 0.0; /* :596 */
    for (int64_t i = 0; i < f_item; ++i) denominator += x[i] * x[i];
    double e_raw = denominator > 1e-12 ? numerator / denominator : 0.0;

    double g_sq_sum = 0.0; /* :611-639 */
    if (edge_energy_sum > 0.0) {
        for (int64_t i = 0; i < f; ++i) {
            double xi = x[i];
            double local_g = 0.0;
            for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
                int64_t j = indices[e];
                if (i != j) {
                    double w = fmax(-data[e], 0.0);
                    if (w > 0.0) {
                        double d = xi - x[j];
                        double contrib = w * d * d;
                        double share = contrib / edge_energy_sum;
                        local_g += share * share;
                    }
                }
            }
            g_sq_sum += local_g;
        }
    }
    double g_raw = g_sq_sum; /* clamp(0,1) :641 */
    if (g_raw < 0.0) g_raw = 0.0;
    if (g_raw > 1.0) g_raw = 1.0;
    double e_bounded = e_raw / (e_raw + tau);  /* :642 */
    return tau * e_bounded + (1.0 - tau) * g_raw; /* :647 */
}

/* src/taumode.rs:230-259 */
int aso_compute_taumode(const double *items, int64_t n, int64_t f, const int64_t *indptr,
                        const int64_t *indices, const double *data, int mode, double value,
                        double *lambdas, int threads) {
    if (n < 0 || f <= 0) return ASO_ERR_INVALID;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t i = 0; i < n; ++i) {
        const double *x = items + i * f;
        double tau = aso_select_tau(x, (size_t)f, mode, value); /* :234 item values */
        lambdas[i] = aso_synthetic_lambda(x, f, indptr, indices, data, tau);
    }
    (void)threads;
    return ASO_OK;
}

/* src/core.rs:533-549 */
int aso_prepare_query_item(const double *q, int64_t f, const int64_t *indptr,
                           const int64_t *indices, const double *data, int mode, double value,
                           double *lambda_out) {
    for (int64_t j = 0; j < f; ++j)
        if (!isfinite(q[j])) return ASO_ERR_NONFINITE_QUERY; /* :534-537 */
    double tau = aso_select_tau(q, (size_t)f, mode, value);
    *lambda_out = aso_synthetic_lambda(q, f, indptr, indices, data, tau);
    return ASO_OK;
}

/* --------------------------------------------------------------- clustering */

/* src/clustering.rs:913-928 */
int64_t aso_nearest_centroid(const double *row, const double *centroids, int64_t k, int64_t f,
                             double *d2_out) {
    int64_t best_idx = 0;
    double best = INFINITY;
    for (int64_t i = 0; i < k; ++i) {
        const double *c = centroids + i * f;
        double d2 = 0.0;
        for (int64_t j = 0; j < f; ++j) {
            double diff = row[j] - c[j];
            d2 += diff * diff;
        }
        if (d2 < best) { /* strict: first minimum wins */
            best = d2;
            best_idx = i;
        }
    }
    if (d2_out) *d2_out = best;
    return best_idx;
}

/* src/clustering.rs:547-910 with deterministic_clustering (:842-843) and no sampler.
 * Sequential execution makes the snapshot (:574-577) equal to the live state, so
 * the "recompute with current centroids" calls (:721, :764) return the snapshot
 * result. */
int aso_cluster_incremental(const double *rows, int64_t n, int64_t f, int64_t max_clusters,
                            double radius, double *centroids, int64_t *assignments,
                            uint64_t *sizes, int64_t *x_out) {
    if (n <= 0 || f <= 0 || max_clusters <= 0) return ASO_ERR_INVALID;
    int64_t kc = 0;
    for (int64_t r = 0; r < n; ++r) assignments[r] = -1;
    for (int64_t r = 0; r < n; ++r) {
        const double *row = rows + r * f;
        if (kc == 0) { /* :637-658 */
            memcpy(centroids, row, (size_t)f * sizeof(double));
            sizes[0] = 1;
            assignments[r] = 0;
            kc = 1;
            continue;
        }
        double d2;
        int64_t b = aso_nearest_centroid(row, centroids, kc, f, &d2);
        if (kc < max_clusters && d2 > radius * 0.5) { /* :672-710 */
            memcpy(centroids + kc * f, row, (size_t)f * sizeof(double));
            sizes[kc] = 1;
            assignments[r] = kc;
            kc++;
        } else if (d2 <= radius) { /* :711-758 */
            double k_new = (double)sizes[b] + 1.0;
            double *c = centroids + b * f;
            for (int64_t j = 0; j < f; ++j) c[j] += (row[j] - c[j]) / k_new;
            sizes[b] += 1;
            assignments[r] = b;
        } else { /* :759-815 */
            if (d2 <= radius * 1.5) {
                sizes[b] += 1;
                assignments[r] = b;
            }
        }
    }
    *x_out = kc;
    return kc == 0 ? ASO_ERR_NO_CLUSTERS : ASO_OK; /* :869-874 */
}

/* src/clustering.rs:118-145 */
int aso_twonn_distances(const double *rows, int64_t n, int64_t f, const int64_t *sample_idx,
                        int64_t s, double *d1, double *d2, int threads) {
    if (n <= 0 || f <= 0 || s < 0) return ASO_ERR_INVALID;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int64_t t = 0; t < s; ++t) {
        int64_t i = sample_idx[t];
        const double *ri = rows + i * f;
        double m1 = INFINITY, m2 = INFINITY;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const double *rj = rows + j * f;
            double acc = 0.0;
            for (int64_t c = 0; c < f; ++c) {
                double df = ri[c] - rj[c];
                acc += df * df;
            }
            double d = sqrt(acc);
            if (d < m1) {
                m2 = m1;
                m1 = d;
            } else if (d < m2) {
                m2 = d;
            }
        }
        d1[t] = m1;
        d2[t] = m2;
    }
    (void)threads;
    return ASO_OK;
}

/* src/clustering.rs:108-110,136-163 */
int64_t aso_intrinsic_dim(int64_t n, int64_t f, const double *d1, const double *d2, int64_t s) {
    if (n < 10) return f < 2 ? f : 2;
    double sum = 0.0;
    int64_t cnt = 0;
    for (int64_t t = 0; t < s; ++t) {
        if (n - 1 >= 2 && d1[t] > 1e-12) { /* dists.len() >= 2 && d1 > 1e-12 */
            sum += d2[t] / d1[t];
            cnt++;
        }
    }
    if (cnt == 0) return f < 3 ? f : 3;
    double mean_ratio = sum / (double)cnt;
    double id = mean_ratio > 1.001 ? 1.0 / log(mean_ratio) : (double)f;
    double rd = round(id);
    int64_t idc = (rd != rd || rd < 0.0) ? 0 : (rd > 9e18 ? INT64_MAX : (int64_t)rd);
    if (idc < 1) idc = 1;
    if (idc > f) idc = f;
    return idc;
}

int64_t aso_intrinsic_dim_from_distances(const double *d1, const double *d2, int64_t s,
                                         int64_t f) {
    return aso_intrinsic_dim(INT64_MAX, f, d1, d2, s);
}

/* src/clustering.rs:85-97 */
void aso_step1_bounds(int64_t n, int64_t f, int64_t id_est, int64_t *k_min, int64_t *k_max) {
    int64_t kmin = (int64_t)ceil(sqrt((double)n / 10.0));
    if (kmin < 2) kmin = 2;
    int64_t cands[4] = {f, n / 10, 5 * id_est, (int64_t)pow((double)n, 0.5)};
    int64_t m = cands[0];
    for (int i = 1; i < 4; ++i)
        if (cands[i] < m) m = cands[i];
    if (m < kmin + 1) m = kmin + 1;
    if (m > n / 2) m = n / 2;
    *k_min = kmin;
    *k_max = m;
}

/* ---------------------------------------------------------------- Laplacian */

typedef struct {
    double dist;
    int64_t j;
} nb_t;

static int cmp_nb(const void *a, const void *b) {
    const nb_t *x = (const nb_t *)a, *y = (const nb_t *)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    return (x->j > y->j) - (x->j < y->j);
}

typedef struct {
    int64_t j;
    double w, score;
    int64_t ord;
} vn_t;

static int cmp_vn_score_desc(const void *a, const void *b) {
    const vn_t *x = (const vn_t *)a, *y = (const vn_t *)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->ord > y->ord) - (x->ord < y->ord); /* reference: unspecified (sort_unstable) */
}

/* src/graph.rs:149-204 -> src/laplacian.rs:122-178 -> :203-417 */
int aso_feature_laplacian(const double *centroids, int64_t x, int64_t f,
                          const aso_lap_params *P, int64_t *indptr, int64_t *indices,
                          double *data, int64_t *nnz_out) {
    if (x < 2 || f < 2) return ASO_ERR_SHAPE; /* laplacian.rs:129-134 */
    int64_t t1 = P->topk + 1;                   /* :211 */
    double sigma = P->has_sigma ? P->sigma : 1.0; /* :254 */
    int rc = ASO_OK;
    /* normalise (laplacian.rs:146-151): StandardScaler::fit(&transposed).transform -- per COLUMN of the F x X matrix =
     * per centroid over its F values.  smartcore 0.4.5 is absent from the mount (PARITY UNPINNED for this branch):
     * restated as mean = sum / F, var = sum(x^2) / F - mean^2 (normalise == 1) or x F / (F - 1) (normalise == 2),
     * zero deviation -> divisor 1. */
    double *scaled = NULL;
    if (P->normalise) {
        scaled = (double *)malloc((size_t)x * (size_t)f * sizeof(double));
        if (!scaled) return ASO_ERR_INVALID;
        for (int64_t c = 0; c < x; ++c) {
            const double *row = centroids + c * f;
            double s = 0.0, s2 = 0.0;
            for (int64_t j = 0; j < f; ++j) {
                s += row[j];
                s2 += row[j] * row[j];
            }
            const double n = (double)f;
            const double mean = s / n;
            double var = s2 / n - mean * mean;
            if (P->normalise == 2) var = var * (n / (n - 1.0));
            double sd = sqrt(var);
            if (!(sd > 0.0)) sd = 1.0;
            for (int64_t j = 0; j < f; ++j) scaled[c * f + j] = (row[j] - mean) / sd;
        }
        centroids = scaled;
    }

    /* node i = feature column i of the centroid matrix (graph.rs:172 transpose) */
    double *mag = (double *)malloc((size_t)f * sizeof(double));
    for (int64_t i = 0; i < f; ++i) {
        double s = 0.0;
        for (int64_t c = 0; c < x; ++c) {
            double v = centroids[c * f + i];
            s += v * v;
        }
        mag[i] = sqrt(s);
    }
    nb_t *cand = (nb_t *)malloc((size_t)f * sizeof(nb_t));
    nb_t *knn = (nb_t *)malloc((size_t)(f * t1) * sizeof(nb_t));
    int64_t *knn_len = (int64_t *)calloc((size_t)f, sizeof(int64_t));
    int64_t *deg = (int64_t *)calloc((size_t)f, sizeof(int64_t));
    double *W = (double *)calloc((size_t)(f * f), sizeof(double));
    unsigned char *A = (unsigned char *)calloc((size_t)(f * f), 1);
    vn_t *vn = (vn_t *)malloc((size_t)(t1 > 0 ? t1 : 1) * sizeof(vn_t));

    for (int64_t i = 0; i < f && rc == ASO_OK; ++i) {
        int64_t nc = 0;
        for (int64_t j = 0; j < f; ++j) {
            if (j == i && !P->self_included) continue;
            double dot = 0.0;
            for (int64_t c = 0; c < x; ++c) dot += centroids[c * f + i] * centroids[c * f + j];
            double cs = dot / (mag[i] * mag[j]);
            if (cs != cs) {
                rc = ASO_ERR_ZERO_NORM;
                break;
            }
            if (P->rectified && cs < 0.0) cs = 0.0;
            cand[nc].dist = 1.0 - cs;
            cand[nc].j = j;
            nc++;
        }
        if (rc != ASO_OK) break;
        qsort(cand, (size_t)nc, sizeof(nb_t), cmp_nb);
        int64_t take = nc < t1 ? nc : t1;
        memcpy(knn + i * t1, cand, (size_t)take * sizeof(nb_t));
        knn_len[i] = take;
        int64_t d = 0; /* :217-227 */
        for (int64_t q = 0; q < take; ++q)
            if (cand[q].j != i && cand[q].dist <= P->eps) d++;
        deg[i] = d;
    }
    if (rc == ASO_OK) {
        int64_t degsum = 0;
        for (int64_t i = 0; i < f; ++i) degsum += deg[i];
        double avg_degree = (double)degsum / (double)f; /* :229 */
        int sparsify = avg_degree > 10.0;               /* :230 */
        for (int64_t i = 0; i < f; ++i) {
            int64_t nv = 0;
            for (int64_t q = 0; q < knn_len[i]; ++q) { /* :249-271 */
                nb_t nb = knn[i * t1 + q];
                if (nb.j != i && nb.dist <= P->eps) {
                    double w = 1.0 / (1.0 + pow(nb.dist / sigma, P->p));
                    if (w > 1e-12) {
                        vn[nv].j = nb.j;
                        vn[nv].w = w;
                        vn[nv].score =
                            sparsify ? w * sqrt((double)(deg[i] * deg[nb.j])) : w;
                        vn[nv].ord = nv;
                        nv++;
                    }
                }
            }
            if (sparsify && nv > 2) { /* :274-280 */
                qsort(vn, (size_t)nv, sizeof(vn_t), cmp_vn_score_desc);
                int64_t keep = nv / 2;
                if (keep < 1) keep = 1;
                nv = keep;
            }
            for (int64_t q = 0; q < nv; ++q) { /* :317-320 union symmetrisation */
                int64_t j = vn[q].j;
                W[i * f + j] = vn[q].w;
                W[j * f + i] = vn[q].w;
                A[i * f + j] = 1;
                A[j * f + i] = 1;
            }
        }
        /* :349-417 : (i,i)=sum_j w_ij in ascending j, stored even when 0; (i,j)=-w */
        int64_t nnz = 0;
        for (int64_t i = 0; i < f; ++i) {
            indptr[i] = nnz;
            double degree = 0.0;
            for (int64_t j = 0; j < f; ++j)
                if (A[i * f + j] && j != i) degree += W[i * f + j];
            int diag_done = 0;
            for (int64_t j = 0; j < f; ++j) {
                if (j == i) {
                    indices[nnz] = i;
                    data[nnz] = degree;
                    nnz++;
                    diag_done = 1;
                } else if (A[i * f + j]) {
                    indices[nnz] = j;
                    data[nnz] = -W[i * f + j];
                    nnz++;
                }
            }
            (void)diag_done;
        }
        indptr[f] = nnz;
        *nnz_out = nnz;
        if (P->sparsity_check) { /* graph.rs:185-193 */
            double sparsity = 1.0 - (double)nnz / (double)(f * f);
            if (sparsity > 0.95) rc = ASO_ERR_TOO_SPARSE;
        }
    }
    free(mag);
    free(cand);
    free(knn);
    free(knn_len);
    free(deg);
    free(W);
    free(A);
    free(vn);
    free(scaled);
    return rc;
}

/* ------------------------------------------------------------------- search */

typedef struct {
    double s;
    int64_t i;
} sc_t;

/* stable bottom-up merge sort, descending by score (core.rs:785 sort_by is stable) */
static int merge_sort_desc(sc_t *a, sc_t *tmp, int64_t n) {
    for (int64_t width = 1; width < n; width *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * width) {
            int64_t mid = lo + width < n ? lo + width : n;
            int64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
            int64_t p = lo, q = mid, o = lo;
            while (p < mid && q < hi) {
                /* take right only if strictly greater -> stable */
                if (a[q].s > a[p].s) tmp[o++] = a[q++];
                else tmp[o++] = a[p++];
            }
            while (p < mid) tmp[o++] = a[p++];
            while (q < hi) tmp[o++] = a[q++];
        }
        memcpy(a, tmp, (size_t)n * sizeof(sc_t));
    }
    return 0;
}

/* src/core.rs:760-798 with :135-239 */
int aso_search_lambda_aware(const double *items, const double *lambdas, int64_t n, int64_t f,
                            const double *q, double lambda_q, int64_t k, double alpha,
                            int64_t *idx_out, double *score_out, int64_t *count_out) {
    if (n <= 0 || f <= 0 || k < 0) return ASO_ERR_INVALID;
    if (lambda_q == 0.0) return ASO_ERR_ZERO_LAMBDA; /* :773-776 */
    sc_t *res = (sc_t *)malloc((size_t)n * sizeof(sc_t));
    sc_t *tmp = (sc_t *)malloc((size_t)n * sizeof(sc_t));
    int has_nan = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double *xr = items + i * f;
        double nq2 = 0.0, nx2 = 0.0; /* norms recomputed per pair, :230 */
        for (int64_t j = 0; j < f; ++j) nq2 += q[j] * q[j];
        for (int64_t j = 0; j < f; ++j) nx2 += xr[j] * xr[j];
        double denom = sqrt(nq2) * sqrt(nx2);
        double cosine = 0.0;
        if (denom > 0.0) {
            double dot = 0.0;
            for (int64_t j = 0; j < f; ++j) dot += q[j] * xr[j];
            cosine = dot / denom;
        }
        double ld = fabs(lambda_q - lambdas[i]); /* :136-137 */
        double lam = 1.0 - fmin(ld, 1.0);        /* f64::min ignores NaN */
        double s = alpha * cosine + (1.0 - alpha) * lam; /* :165 */
        if (s != s) has_nan = 1;
        res[i].s = s;
        res[i].i = i;
    }
    int rc = ASO_OK;
    if (has_nan && n > 1) {
        rc = ASO_ERR_NAN_SCORE; /* partial_cmp().unwrap() panics, :785 */
    } else {
        merge_sort_desc(res, tmp, n);
        int64_t cnt = k < n ? k : n; /* truncate :786 */
        for (int64_t r = 0; r < cnt; ++r) {
            idx_out[r] = res[r].i;
            score_out[r] = res[r].s;
        }
        *count_out = cnt;
    }
    free(res);
    free(tmp);
    return rc;
}

int aso_search_lambda_aware_batch(const double *items, const double *lambdas, int64_t n,
                                  int64_t f, const double *queries, const double *lambda_q,
                                  int64_t nq, int64_t k, double alpha, int64_t *idx_out,
                                  double *score_out, int64_t *count_out, int threads) {
    int rc_all = ASO_OK;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int64_t t = 0; t < nq; ++t) {
        int64_t cnt = 0;
        int rc = aso_search_lambda_aware(items, lambdas, n, f, queries + t * f, lambda_q[t], k,
                                         alpha, idx_out + t * k, score_out + t * k, &cnt);
        count_out[t] = cnt;
        if (rc != ASO_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = rc;
        }
    }
    (void)threads;
    return rc_all;
}

/* ------------------------------------------------ "next" rows */

typedef struct {
    double s;
    int64_t i;
} hs_t;

static int cmp_hs_desc(const void *a, const void *b) {
    const hs_t *x = (const hs_t *)a, *y = (const hs_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i); /* reference: unspecified (rayon fold + sort_unstable) */
}

/* src/core.rs:802-928.  Union of {cos > 0.9999} (scored by cosine), the lambda-aware top-k
 * (scored alpha*cos + (1-alpha)*lam unless already present) and the semantic top-1 (cosine),
 * sorted by that score, truncated to k.  Ties (unspecified in the reference) -> lower index. */
int aso_search_lambda_aware_hybrid(const double *items, const double *lambdas, int64_t n, int64_t f,
                                   const double *q, double lambda_q, int64_t k, double alpha,
                                   int64_t *idx_out, double *score_out, int64_t *count_out) {
    if (n <= 0 || f <= 0 || k < 0) return ASO_ERR_INVALID;
    *count_out = 0;
    if (k == 0) return ASO_OK; /* :810-812 */
    double beta = 1.0 - alpha;
    double *cosv = (double *)malloc((size_t)n * sizeof(double));
    hs_t *ls = (hs_t *)malloc((size_t)n * sizeof(hs_t));
    double nq2 = 0.0;
    for (int64_t j = 0; j < f; ++j) nq2 += q[j] * q[j];
    int64_t sem_i = 0;
    double sem_s = -INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        const double *xr = items + i * f;
        double nx2 = 0.0, dot = 0.0;
        for (int64_t j = 0; j < f; ++j) nx2 += xr[j] * xr[j];
        double denom = sqrt(nq2) * sqrt(nx2);
        if (denom > 0.0) {
            for (int64_t j = 0; j < f; ++j) dot += q[j] * xr[j];
            cosv[i] = dot / denom;
        } else {
            cosv[i] = 0.0;
        }
        double lam = 1.0 - fmin(fabs(lambda_q - lambdas[i]), 1.0);
        ls[i].s = alpha * cosv[i] + beta * lam; /* :834 */
        ls[i].i = i;
        if (cosv[i] > sem_s) { /* :837-839 */
            sem_s = cosv[i];
            sem_i = i;
        }
    }
    qsort(ls, (size_t)n, sizeof(hs_t), cmp_hs_desc);
    int64_t kk = k < n ? k : n;
    /* union */
    int64_t cap = kk + 1, cnt = 0;
    for (int64_t i = 0; i < n; ++i)
        if (cosv[i] > 0.9999) cap++;
    hs_t *u = (hs_t *)malloc((size_t)cap * sizeof(hs_t));
    unsigned char *in = (unsigned char *)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i)
        if (cosv[i] > 0.9999) { /* :842-844, :902-905 */
            u[cnt].i = i;
            u[cnt].s = cosv[i];
            cnt++;
            in[i] = 1;
        }
    for (int64_t r = 0; r < kk; ++r) /* :908-911 or_insert */
        if (!in[ls[r].i]) {
            u[cnt] = ls[r];
            cnt++;
            in[ls[r].i] = 1;
        }
    if (!in[sem_i]) { /* :914-915 */
        u[cnt].i = sem_i;
        u[cnt].s = sem_s;
        cnt++;
    }
    qsort(u, (size_t)cnt, sizeof(hs_t), cmp_hs_desc);
    int64_t outn = cnt < k ? cnt : k;
    for (int64_t r = 0; r < outn; ++r) {
        idx_out[r] = u[r].i;
        score_out[r] = u[r].s;
    }
    *count_out = outn;
    free(cosv);
    free(ls);
    free(u);
    free(in);
    return ASO_OK;
}

/* src/core.rs:944-976 (after the lambda==0 re-preparation, which is host logic): every item with
 * lambda_q - lambda_i <= eps (signed difference, as written), in index order. */
int aso_range_search(const double *lambdas, int64_t n, double lambda_q, double eps, int64_t *idx_out,
                     double *dist_out, int64_t *count_out) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) {
        double d = lambda_q - lambdas[i];
        if (d <= eps) {
            idx_out[c] = i;
            dist_out[c] = d;
            c++;
        }
    }
    *count_out = c;
    return ASO_OK;
}

/* EnergyMaps::search_energy (src/energymaps.rs:368-407) scoring every item with ProjectedEnergy::score
 * (:884-894) for an ArrowSpace without projection_matrix (project_vec is the identity, :858-864) and without
 * signals (projected_dirichlet falls back to bounded_l2_energy, :866-882,:846-850).  The reference recomputes
 * lambda_q per item (:885); it is the same number every time and is an input here.  Scores are -energy, sorted
 * descending with a stable sort (:397; partial_cmp().unwrap_or(Equal): a NaN score has no defined place -- reported
 * as ASO_ERR_NAN_SCORE), truncated to k (:398). */
int aso_search_energy(const double *items, const double *lambdas, int64_t n, int64_t f, const double *q,
                      double lambda_q, int64_t k, double w_lambda, double w_dirichlet, int64_t *idx_out,
                      double *score_out, int64_t *count_out) {
    if (n <= 0 || f <= 0 || k < 0) return ASO_ERR_INVALID;
    sc_t *res = (sc_t *)malloc((size_t)n * sizeof(sc_t));
    sc_t *tmp = (sc_t *)malloc((size_t)n * sizeof(sc_t));
    int has_nan = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double *xr = items + i * f;
        const double d_lambda = fabs(lambda_q - lambdas[i]); /* :887 */
        double ss = 0.0;                                     /* l2_norm(vec_diff(q, x)), :841-843,:892 */
        for (int64_t j = 0; j < f; ++j) {
            const double df = q[j] - xr[j];
            ss += df * df;
        }
        const double num = sqrt(ss);
        const double d_dir = fmin(num / (1.0 + num), 1.0);   /* bounded_l2_energy, :847-850 */
        const double e = w_lambda * d_lambda + w_dirichlet * d_dir; /* :894 */
        if (e != e) has_nan = 1;
        res[i].s = -e; /* :393 */
        res[i].i = i;
    }
    int rc = ASO_OK;
    if (has_nan) {
        rc = ASO_ERR_NAN_SCORE;
    } else {
        merge_sort_desc(res, tmp, n);
        int64_t cnt = k < n ? k : n;
        for (int64_t r = 0; r < cnt; ++r) {
            idx_out[r] = res[r].i;
            score_out[r] = res[r].s;
        }
        *count_out = cnt;
    }
    free(res);
    free(tmp);
    return rc;
}

int aso_project_matrix(const double *rows, int64_t n, int64_t f, const double *projection, int64_t r, double *out);

/* ProjectedEnergy::score, every branch (src/energymaps.rs:856-895): project_vec through a materialised projection (f x r,
 * or NULL), projected_dirichlet through the signals CSR (d x d, or NULL; used when its column count equals the
 * difference's length, :868) else bounded L2.  lambda_q is an input (prepare_query_item of the PROJECTED query,
 * src/core.rs:540-548, computed by the caller with aso_compute_taumode on the projected query).  Same ordering rules as
 * aso_search_energy. */
int aso_search_energy_ex(const double *items, const double *lambdas, int64_t n, int64_t f, const double *q,
                         double lambda_q, int64_t k, double w_lambda, double w_dirichlet, const double *projection,
                         int64_t r, const int64_t *sig_indptr, const int64_t *sig_indices, const double *sig_data,
                         int64_t sig_n, int64_t *idx_out, double *score_out, int64_t *count_out) {
    if (n <= 0 || f <= 0 || k < 0) return ASO_ERR_INVALID;
    const int64_t d = projection ? r : f;
    sc_t *res = (sc_t *)malloc((size_t)n * sizeof(sc_t));
    sc_t *tmp = (sc_t *)malloc((size_t)n * sizeof(sc_t));
    double *qp = (double *)malloc((size_t)d * sizeof(double));
    double *xp = (double *)malloc((size_t)d * sizeof(double));
    double *diff = (double *)malloc((size_t)d * sizeof(double));
    double *y = (double *)malloc((size_t)(sig_n > 0 ? sig_n : 1) * sizeof(double));
    if (projection) aso_project_matrix(q, 1, f, projection, r, qp);      /* project_vec(query), :889 */
    else memcpy(qp, q, (size_t)f * sizeof(double));
    int has_nan = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double d_lambda = fabs(lambda_q - lambdas[i]);               /* :887 */
        if (projection) aso_project_matrix(items + i * f, 1, f, projection, r, xp);   /* project_vec(item), :891 */
        else memcpy(xp, items + i * f, (size_t)f * sizeof(double));
        for (int64_t j = 0; j < d; ++j) diff[j] = qp[j] - xp[j];            /* vec_diff, :499-501,:892 */
        double ss = 0.0;
        if (sig_indptr && sig_n > 0 && sig_n == d) {                        /* :868: signals.cols() == diff.len() */
            for (int64_t row = 0; row < sig_n; ++row) {                     /* :869-876 */
                double sum = 0.0;
                for (int64_t e = sig_indptr[row]; e < sig_indptr[row + 1]; ++e) sum += sig_data[e] * diff[sig_indices[e]];
                y[row] = sum;
            }
            for (int64_t row = 0; row < sig_n; ++row) ss += y[row] * y[row];
        } else {
            for (int64_t j = 0; j < d; ++j) ss += diff[j] * diff[j];
        }
        const double num = sqrt(ss);
        const double d_dir = fmin(num / (1.0 + num), 1.0);                  /* :878 / :847-850 */
        const double e = w_lambda * d_lambda + w_dirichlet * d_dir;         /* :894 */
        if (e != e) has_nan = 1;
        res[i].s = -e;
        res[i].i = i;
    }
    int rc = ASO_OK;
    if (has_nan) {
        rc = ASO_ERR_NAN_SCORE;
    } else {
        merge_sort_desc(res, tmp, n);
        int64_t cnt = k < n ? k : n;
        for (int64_t t = 0; t < cnt; ++t) {
            idx_out[t] = res[t].i;
            score_out[t] = res[t].s;
        }
        *count_out = cnt;
    }
    free(res);
    free(tmp);
    free(qp);
    free(xp);
    free(diff);
    free(y);
    return rc;
}

/* compute_jl_dimension, src/reduction.rs:127-141 */
int64_t aso_jl_dimension(int64_t n_points, double epsilon) {
    const double log_n = log((double)n_points);
    const double eps_sq = pow(epsilon, 2.0);
    const double v = ceil(8.0 * log_n / eps_sq);
    int64_t jl = (v != v || v < 0.0) ? 0 : (v > 9.0e18 ? INT64_MAX : (int64_t)v); /* `as usize` saturates */
    return jl > 32 ? jl : 32;
}

/* project_matrix -> ImplicitProjection::project per row, src/reduction.rs:143-166,180-199: for each feature (outer
 * loop) and each output (inner loop) `*reduced += original * sample * scale`, samples drawn in that order. */
int aso_project_matrix(const double *rows, int64_t n, int64_t f, const double *projection, int64_t r, double *out) {
    if (!rows || !projection || !out || n <= 0 || f <= 0 || r <= 0) return ASO_ERR_INVALID;
    const double scale = 1.0 / sqrt((double)r); /* :184 */
    for (int64_t i = 0; i < n; ++i) {
        double *y = out + i * r;
        for (int64_t k = 0; k < r; ++k) y[k] = 0.0;
        for (int64_t j = 0; j < f; ++j) {
            const double original = rows[i * f + j];
            for (int64_t k = 0; k < r; ++k) y[k] += original * projection[j * r + k] * scale; /* :195 */
        }
    }
    return ASO_OK;
}
